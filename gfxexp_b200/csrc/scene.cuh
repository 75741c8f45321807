// scene.cuh — device-side scene tables, BVH views and frame buffers (all plain pointers into HBM).
//
// Layout decisions (B200-first, see DESIGN.md "Data layout in HBM"):
//  * vertices: AoS of three float4 per vertex (pos+u | normal+v | tangent) = 48 B, so a gather of one
//    vertex is three aligned 16-byte loads from two 32-byte sectors (the reference's shared::Vertex
//    is 44 B and straddles sectors; common_shared.h:1109-1114);
//  * triangles: uint4 {i0,i1,i2,-} per mesh triangle (aligned 16-byte load);
//  * BVH: the reference's own formats (CompressedInternalNode_T<8> 80 B = five 16-byte loads,
//    TriangleStorage 48 B = three 16-byte loads, PrimitiveReference 4 B), world-space triangles of
//    ALL instances flattened into one single-level BVH (180 GB HBM makes the IAS/GAS split of the
//    reference unnecessary and saves the per-instance ray transform);
//  * per-pixel frame state: SoA planes of float4/uint4 (one 16-byte load per thread, fully coalesced
//    along x), reservoirs as three float4 planes.
#pragma once
#include "vec.cuh"
#include "../../include/gfxb200.h"

namespace gfx {

constexpr uint32_t kInstGuideSize = 2048;
constexpr uint32_t kPrimGuideSize = 128;

struct DevMesh {
    uint32_t vertexBase;   // into vertices (units of vertices)
    uint32_t triBase;      // into triangles / primWeights / primCdf
    uint32_t numTriangles;
    uint32_t materialSlot;
    float primIntegral;    // emitterPrimDist.integral()
    uint32_t pad[3];
};

// GFX_WIDE_TABLE_LOADS (compile-time, default 0): the light-sampling chain reads an instance's four sampling fields with one
// 128-bit load and its normal matrix with 2 x 128 + 1 x 32 bits instead of 4 + 9 scalar loads.  ncu source view of the RIS
// candidate kernel (profiles/r01_summary.md, r01f): those 13 divergent scalar loads are 502 M of its 1 210 M L1 tag requests
// and the kernel runs at 80 % of the L1 data-pipe rate.  Written at the end of round 1 without GPU time left to validate
// it, therefore off; the values loaded are identical, so parity is not at stake once it is switched on and measured.
#ifndef GFX_WIDE_TABLE_LOADS
#define GFX_WIDE_TABLE_LOADS 0
#endif

// GFX_LIGHT_CULL_SPHERES (compile-time, default 0): every light-triangle record gets a seventh float4, the bounding sphere of
// the world-space triangle (centre, radius; radius < 0 = never cull), read first by sampleLightUnlessDark: on config 2 65 % of
// the RIS candidates lie below the shading horizon (tools/ris_candidate_stats.py) and most of them can be rejected from the
// sphere alone - one 16-byte fetch instead of three.  Like GFX_WIDE_TABLE_LOADS: written after round 1's GPU minutes were
// spent, unvalidated, therefore off; the default build's SASS is unchanged.
#ifndef GFX_LIGHT_CULL_SPHERES
#define GFX_LIGHT_CULL_SPHERES 0
#endif
constexpr uint32_t kLightTriStride = GFX_LIGHT_CULL_SPHERES ? 7u : 6u; // float4 per light triangle

struct DevInstance {
    float transform[12];
    float curToPrevTransform[12];
    float normalMatrix[9];
    float uniformScale;
#if GFX_WIDE_TABLE_LOADS
    uint32_t pad[2];       // the four light-sampling fields below then start a 16-byte line (offset 144): one 128-bit load
#endif
    uint32_t firstMeshSlot;
    uint32_t numMeshSlots;
    float geomIntegral;    // lightGeomInstDist.integral()
    uint32_t geomBase;     // index of this instance's first flattened geometry (instance order)
#if !GFX_WIDE_TABLE_LOADS
    uint32_t pad[2];
#endif
};
static_assert(sizeof(DevInstance) % 16 == 0, "DevInstance must stay 16-byte aligned");

struct DevBvh {
    const uint4* nodes;     // 5 x uint4 per node
    const uint32_t* primRefs;
    const float4* tris;     // 3 x float4 per triangle (reference layout, indexed by TriangleStorage index)
    const float4* leafTris; // 3 x float4 per primitive reference, leaf order; .w of the third = primRef word
    uint32_t numNodes;
    uint32_t* overflowFlag; // set to 1 if a traversal stack overflowed (checked by the host)
};

struct DevScene {
    const float4* vertices;          // 3 per vertex
    const uint4* triangles;
    const DevMesh* meshes;
    const GfxMaterialDesc* materials;
    const DevInstance* instances;
    const uint32_t* instanceMeshSlots;
    const uint2* geomToInstMesh;     // flattened geometry -> (instance, mesh slot)
    const float* primWeights;        // per mesh triangle
    const float* primCdf;
    const float* geomWeights;        // per (instance, mesh slot)
    const float* geomCdf;
    const float* instWeights;        // per instance
    const float* instCdf;
    const float* instIntegral;       // device scalar (rebuilt every frame on the GPU)
    // pre-divided selection probabilities weights[i] / integral (same IEEE division the sampler did)
    const float* primProb;
    const float* geomProb;
    const float* instProb;
    // world-space table of the triangles of every emissive geometry, 6 float4 each:
    // (pA, recArea) (pB, nA.x) (pC, nA.y) (nA.z, nB) (nC, -) (emittance, -); lightTriBase[g] = first entry
    // of flattened geometry g or 0xFFFFFFFF.  Values are produced by the very expressions of
    // sampleLight (restir_di_shared.h:417-425,485-511), so reading them is bit-identical to recomputing.
    const float4* lightTris;
    const uint32_t* lightTriBase;
    // guide tables for the CDF searches: guide[b] = search(cdf, fl(b / G * integral)), b = 0..G, so the
    // answer for u = fl(ul * integral) with ul in [b/G, (b+1)/G) lies in [guide[b], guide[b+1]]
    // (rounding is monotone) and a short scan finishes the exact search.
    const uint32_t* instGuide;       // kInstGuideSize + 1 entries
    const uint32_t* primGuide;       // per mesh: kPrimGuideSize + 1 entries
    uint32_t numInstances;
    unsigned long long* rayCounter;  // frame statistics: rays traced (primary + visibility)
    DevBvh bvh;
};

struct DevFrame {
    uint32_t W, H;
    uint4* gb0[2];
    float2* gb1[2];
    float4* gb2[2];
    uint4* gb3[2];
    unsigned long long* rng;
    float4* reservoir[2];       // 3 planes of W*H float4
    float2* reservoirInfo[2];
    float4* beauty;
    float4* albedo;
    float4* normal;
    const float2* neighborDeltas; // 1024 entries
    unsigned long long* stats;    // [0] rays traced (primary + visibility)
    float4* rayQueue;             // wavefront visibility queue, see context.h
    uint32_t* rayPixel;
    uint32_t* rayCounters;
    uint8_t* visibility;
};

struct DevCamera {
    float aspect, fovY;
    f3 position;
    float orientation[9];
    float invOrientation[9];
    float vh, vw;
};

struct DevFrameParams {
    DevCamera camera, prevCamera;
    uint32_t numAccumFrames, frameIndex, bufferIndex;
    float spatialNeighborRadius;
    uint32_t log2NumCandidateSamples, numSpatialNeighbors;
    uint32_t useLowDiscrepancyNeighbors, reuseVisibility, enableTemporalReuse, enableSpatialReuse;
    uint32_t useUnbiasedEstimator, resetFlowBuffer, enableJittering;
    uint32_t currentReservoirIndex, spatialNeighborBaseIndex;
    uint32_t y0, y1; // rows owned by this rank
    uint32_t maxPathLength;
    f3 sceneAabbMin, sceneAabbMax;
    float radianceScale;
    uint32_t reuseVisibilityForTemporal, reuseVisibilityForSpatiotemporal;
    float radiusThresholdForSpatialVisReuse;
};

} // namespace gfx
