// lights.cu — light importance, CDFs, selection probabilities and the emissive-triangle table.
//
// Replaces common/gpu_kernels/compute_light_probs.cu (computeTriangleProbBuffer :68-82,
// computeGeomInstProbBuffer :115-129, computeInstProbBuffer :162-174, finalizeDiscreteDistribution1D
// :206-212) and the ext/cubd ExclusiveSum calls in Scene::setupLightGeomDistributions /
// setupLightInstDistribution (common/common_host.h:1102-1359).
//
// CUB's scan order is unspecified, which makes CDF entries irreproducible at the ULP level; here
// every CDF is a *sequential* fp32 exclusive scan (the order the oracle uses), so CDFs are
// bit-identical.  The triangle- and geometry-level distributions are static and built once; the
// instance-level one is rebuilt per frame by ONE single-block kernel (importance in parallel ->
// sequential scan out of shared memory -> probabilities in parallel), as the reference rebuilds it
// every frame (restir_di_main.cpp:2303-2309).
//
// Two derived tables make the 32-candidate RIS loop cheap without changing a single bit:
//  * prob[i] = weights[i] / integral — the very division DiscreteDistribution1D::sample performs
//    (common_shared.h:243), hoisted out of the per-candidate path;
//  * lightTris — for every triangle of every emissive geometry the world-space vertices
//    (inst.transform * v.position), recArea = 2 / |cross|, the object-space vertex normals and the
//    material emittance, i.e. the operands sampleLight (restir_di_shared.h:417-425,485-511) would
//    recompute for each of the 66 M candidates per frame.
#include "scene.cuh"
#include "context.h"

namespace gfx {

// DiscreteDistribution1DTemplate::sample's stepping search (common_shared.h:226-232): last index with cdf <= u
GFX_D uint32_t steppingSearch(const float* cdf, uint32_t numValues, float u) {
    int idx = 0;
    uint32_t p2 = numValues <= 1 ? numValues : 1u << (32 - __clz(numValues - 1));
    for (int d = (int)(p2 >> 1); d >= 1; d >>= 1) {
        if (idx + d >= (int)numValues)
            continue;
        if (cdf[idx + d] <= u)
            idx += d;
    }
    return (uint32_t)idx;
}

// per-mesh guide table for the primitive-level search
__global__ void k_primGuide(DevScene scene, uint32_t numMeshes) {
    const uint32_t mesh = blockIdx.x;
    if (mesh >= numMeshes)
        return;
    const DevMesh m = scene.meshes[mesh];
    uint32_t* guide = const_cast<uint32_t*>(scene.primGuide) + (size_t)mesh * (kPrimGuideSize + 1);
    for (uint32_t b = threadIdx.x; b <= kPrimGuideSize; b += blockDim.x) {
        const float u = ((float)b / (float)kPrimGuideSize) * m.primIntegral;
        guide[b] = m.numTriangles ? steppingSearch(scene.primCdf + m.triBase, m.numTriangles, u) : 0u;
    }
}

// computeTriangleImportance (compute_light_probs.cu:22-46)
__global__ void k_triangleImportance(DevScene scene, uint32_t numMeshes) {
    const uint32_t mesh = blockIdx.y;
    if (mesh >= numMeshes)
        return;
    const DevMesh m = scene.meshes[mesh];
    const GfxMaterialDesc mat = scene.materials[m.materialSlot];
    float* w = const_cast<float*>(scene.primWeights);
    for (uint32_t t = blockIdx.x * blockDim.x + threadIdx.x; t < m.numTriangles; t += gridDim.x * blockDim.x) {
        const uint4 tri = scene.triangles[m.triBase + t];
        const float4 a = scene.vertices[3 * (size_t)(m.vertexBase + tri.x)];
        const float4 b = scene.vertices[3 * (size_t)(m.vertexBase + tri.y)];
        const float4 c = scene.vertices[3 * (size_t)(m.vertexBase + tri.z)];
        const f3 p0(a.x, a.y, a.z), p1(b.x, b.y, b.z), p2(c.x, c.y, c.z);
        const f3 normal = cross(p1 - p0, p2 - p0);
        const float area = 0.5f * length(normal);
        const f3 e = mat.hasEmittance ? f3(mat.emittance[0], mat.emittance[1], mat.emittance[2]) : f3(0.0f);
        f3 emittanceEstimate(0.0f);
        emittanceEstimate += e;
        emittanceEstimate += e;
        emittanceEstimate += e;
        emittanceEstimate /= 3;
        w[m.triBase + t] = sRGB_calcLuminance(emittanceEstimate) * area;
    }
}

// one thread per distribution: sequential exclusive scan + finalize (common_shared.h:268-271)
__global__ void k_scanMeshes(DevScene scene, uint32_t numMeshes) {
    const uint32_t mesh = blockIdx.x * blockDim.x + threadIdx.x;
    if (mesh >= numMeshes)
        return;
    DevMesh* m = const_cast<DevMesh*>(scene.meshes) + mesh;
    const float* w = scene.primWeights + m->triBase;
    float* cdf = const_cast<float*>(scene.primCdf) + m->triBase;
    float sum = 0.0f;
    float last = 0.0f, lastCdf = 0.0f;
    for (uint32_t t = 0; t < m->numTriangles; ++t) {
        cdf[t] = sum;
        lastCdf = sum;
        last = w[t];
        sum = sum + last;
    }
    m->primIntegral = m->numTriangles ? lastCdf + last : 0.0f;
}

__global__ void k_primProb(DevScene scene, uint32_t numMeshes) {
    const uint32_t mesh = blockIdx.y;
    if (mesh >= numMeshes)
        return;
    const DevMesh m = scene.meshes[mesh];
    float* prob = const_cast<float*>(scene.primProb);
    for (uint32_t t = blockIdx.x * blockDim.x + threadIdx.x; t < m.numTriangles; t += gridDim.x * blockDim.x)
        prob[m.triBase + t] = scene.primWeights[m.triBase + t] / m.primIntegral;
}

__global__ void k_scanInstanceGeoms(DevScene scene, uint32_t numInstances) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= numInstances)
        return;
    DevInstance* inst = const_cast<DevInstance*>(scene.instances) + i;
    float* w = const_cast<float*>(scene.geomWeights) + inst->firstMeshSlot;
    float* cdf = const_cast<float*>(scene.geomCdf) + inst->firstMeshSlot;
    float* prob = const_cast<float*>(scene.geomProb) + inst->firstMeshSlot;
    float sum = 0.0f, last = 0.0f, lastCdf = 0.0f;
    for (uint32_t k = 0; k < inst->numMeshSlots; ++k) {
        // computeGeomInstImportance (compute_light_probs.cu:68-82)
        const float imp = scene.meshes[scene.instanceMeshSlots[inst->firstMeshSlot + k]].primIntegral;
        w[k] = imp;
        cdf[k] = sum;
        lastCdf = sum;
        last = imp;
        sum = sum + imp;
    }
    const float integral = inst->numMeshSlots ? lastCdf + last : 0.0f;
    inst->geomIntegral = integral;
    for (uint32_t k = 0; k < inst->numMeshSlots; ++k)
        prob[k] = w[k] / integral;
}

// emissive-triangle table: one block per emissive flattened geometry
__global__ void k_lightTris(DevScene scene, const uint32_t* __restrict__ emissiveGeoms, uint32_t numEmissiveGeoms) {
    if (blockIdx.x >= numEmissiveGeoms)
        return;
    const uint32_t g = emissiveGeoms[blockIdx.x];
    const uint2 im = scene.geomToInstMesh[g];
    const DevInstance* inst = scene.instances + im.x;
    const DevMesh mesh = scene.meshes[im.y];
    const GfxMaterialDesc* mat = scene.materials + mesh.materialSlot;
    float4* out = const_cast<float4*>(scene.lightTris) + kLightTriStride * (size_t)scene.lightTriBase[g];
    for (uint32_t prim = threadIdx.x; prim < mesh.numTriangles; prim += blockDim.x) {
        const uint4 tri = scene.triangles[mesh.triBase + prim];
        const float4* vA = scene.vertices + 3 * (size_t)(mesh.vertexBase + tri.x);
        const float4* vB = scene.vertices + 3 * (size_t)(mesh.vertexBase + tri.y);
        const float4* vC = scene.vertices + 3 * (size_t)(mesh.vertexBase + tri.z);
        const float4 a0 = vA[0], a1 = vA[1], b0 = vB[0], b1 = vB[1], c0 = vC[0], c1 = vC[1];
        // restir_di_shared.h:417-425
        const f3 pA = xfmPoint(inst->transform, f3(a0.x, a0.y, a0.z));
        const f3 pB = xfmPoint(inst->transform, f3(b0.x, b0.y, b0.z));
        const f3 pC = xfmPoint(inst->transform, f3(c0.x, c0.y, c0.z));
        const f3 geomNormal = cross(pB - pA, pC - pA);
        const float recArea = 2.0f / length(geomNormal); // :496
        f3 emittance(0.0f);
        if (mat->hasEmittance) { // :505-511 with a 1x1 emittance texture
            emittance = f3(1.0f, 1.0f, 1.0f);
            emittance *= f3(mat->emittance[0], mat->emittance[1], mat->emittance[2]);
        }
        float4* o = out + kLightTriStride * (size_t)prim;
        o[0] = make_float4(pA.x, pA.y, pA.z, recArea);
        o[1] = make_float4(pB.x, pB.y, pB.z, a1.x);
        o[2] = make_float4(pC.x, pC.y, pC.z, a1.y);
        o[3] = make_float4(a1.z, b1.x, b1.y, b1.z);
        o[4] = make_float4(c1.x, c1.y, c1.z, 0.0f);
        o[5] = make_float4(emittance.x, emittance.y, emittance.z, 0.0f);
#if GFX_LIGHT_CULL_SPHERES
        {   // bounding sphere of the world-space triangle, slightly inflated; negative radius = never cull (degenerate area)
            const f3 c = (pA + pB + pC) * (1.0f / 3.0f);
            const float r2 = fmaxf(fmaxf(sqLength(pA - c), sqLength(pB - c)), sqLength(pC - c));
            const bool usable = recArea > 0.0f && isfinite(recArea) && isfinite(r2);
            o[6] = make_float4(c.x, c.y, c.z, usable ? sqrtf(r2) * 1.0001f + 1e-30f : -1.0f);
        }
#endif
    }
}

// per frame: computeInstImportance (compute_light_probs.cu:115-129) + exclusive scan + finalize +
// selection probabilities, one block, weights staged in shared memory
__global__ void __launch_bounds__(1024) k_instanceDist(DevScene scene, uint32_t numInstances) {
    extern __shared__ float smem[]; // weights[n], cdf[n]
    float* sw = smem;
    float* sc = smem + numInstances;
    __shared__ float sIntegral;
    for (uint32_t i = threadIdx.x; i < numInstances; i += blockDim.x) {
        const DevInstance* inst = scene.instances + i;
        sw[i] = pow2f(inst->uniformScale) * inst->geomIntegral;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        float sum = 0.0f, last = 0.0f, lastCdf = 0.0f;
        for (uint32_t i = 0; i < numInstances; ++i) {
            sc[i] = sum;
            lastCdf = sum;
            last = sw[i];
            sum = sum + last;
        }
        sIntegral = numInstances ? lastCdf + last : 0.0f;
        *const_cast<float*>(scene.instIntegral) = sIntegral;
    }
    __syncthreads();
    const float integral = sIntegral;
    for (uint32_t i = threadIdx.x; i < numInstances; i += blockDim.x) {
        const_cast<float*>(scene.instWeights)[i] = sw[i];
        const_cast<float*>(scene.instCdf)[i] = sc[i];
        const_cast<float*>(scene.instProb)[i] = sw[i] / integral;
    }
    for (uint32_t b = threadIdx.x; b <= kInstGuideSize; b += blockDim.x) {
        const float u = ((float)b / (float)kInstGuideSize) * integral;
        const_cast<uint32_t*>(scene.instGuide)[b] = steppingSearch(sc, numInstances, u);
    }
}

int buildLightDistributions(gfx_ctx* ctx, cudaStream_t stream, uint32_t /*bufferIndex*/) {
    SceneState &S = ctx->scene;
    const DevScene dev = ctx->devScene();
    if (!S.staticLightDistsBuilt) {
        if (S.numMeshes) {
            k_triangleImportance<<<dim3(64, S.numMeshes), 128, 0, stream>>>(dev, S.numMeshes); ctx->launches++;
            k_scanMeshes<<<(S.numMeshes + 31) / 32, 32, 0, stream>>>(dev, S.numMeshes); ctx->launches++;
            k_primProb<<<dim3(64, S.numMeshes), 128, 0, stream>>>(dev, S.numMeshes); ctx->launches++;
            k_primGuide<<<S.numMeshes, 128, 0, stream>>>(dev, S.numMeshes); ctx->launches++;
        }
        if (S.numInstances) {
            k_scanInstanceGeoms<<<(S.numInstances + 63) / 64, 64, 0, stream>>>(dev, S.numInstances); ctx->launches++;
        }
        S.staticLightDistsBuilt = true;
        S.lightTrisDirty = true;
    }
    if (S.lightTrisDirty && S.numEmissiveGeoms) {
        // world-space light triangles follow the instance transforms (gfx_scene_update_instances marks them dirty)
        k_lightTris<<<S.numEmissiveGeoms, 128, 0, stream>>>(dev, S.emissiveGeoms, S.numEmissiveGeoms); ctx->launches++;
    }
    S.lightTrisDirty = false;
    if (S.numInstances) {
        if (S.numInstances > 28000) {
            ctx->setError("gfx_light_dist_build: more than 28 000 instances (shared-memory scan limit)");
            return GFX_ERR_UNSUPPORTED;
        }
        const size_t smem = (size_t)S.numInstances * 8;
        if (smem > 48 * 1024)
            GFX_CUDA(ctx, cudaFuncSetAttribute(k_instanceDist, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        k_instanceDist<<<1, 1024, smem, stream>>>(dev, S.numInstances); ctx->launches++;
    }
    GFX_CUDA(ctx, cudaGetLastError());
    return GFX_OK;
}

} // namespace gfx
