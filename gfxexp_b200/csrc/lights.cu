// lights.cu — light importance + CDFs on the device.
//
// Replaces common/gpu_kernels/compute_light_probs.cu (computeTriangleProbBuffer :68-82,
// computeGeomInstProbBuffer :115-129, computeInstProbBuffer :162-174, finalizeDiscreteDistribution1D
// :206-212) and the ext/cubd ExclusiveSum calls in Scene::setupLightGeomDistributions /
// setupLightInstDistribution (common/common_host.h:1102-1359).
//
// CUB's scan order is unspecified, which makes CDF entries irreproducible at the ULP level; here
// every CDF is a *sequential* fp32 exclusive scan (one thread per distribution), which is the order
// the oracle uses, so CDFs are bit-identical.  The triangle- and geometry-level distributions are
// static and built once; only the instance-level one (<= 16 k entries) is rebuilt per frame, as in
// the reference (restir_di_main.cpp:2303-2309).
#include "scene.cuh"
#include "context.h"

namespace gfx {

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

__global__ void k_scanInstanceGeoms(DevScene scene, uint32_t numInstances) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= numInstances)
        return;
    DevInstance* inst = const_cast<DevInstance*>(scene.instances) + i;
    float* w = const_cast<float*>(scene.geomWeights) + inst->firstMeshSlot;
    float* cdf = const_cast<float*>(scene.geomCdf) + inst->firstMeshSlot;
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
    inst->geomIntegral = inst->numMeshSlots ? lastCdf + last : 0.0f;
}

// computeInstImportance (compute_light_probs.cu:115-129) + scan + finalize, per frame
__global__ void k_instanceImportance(DevScene scene, uint32_t numInstances) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= numInstances)
        return;
    const DevInstance* inst = scene.instances + i;
    const_cast<float*>(scene.instWeights)[i] = pow2f(inst->uniformScale) * inst->geomIntegral;
}
__global__ void k_scanInstances(DevScene scene, uint32_t numInstances) {
    if (blockIdx.x != 0 || threadIdx.x != 0)
        return;
    const float* w = scene.instWeights;
    float* cdf = const_cast<float*>(scene.instCdf);
    float sum = 0.0f, last = 0.0f, lastCdf = 0.0f;
    for (uint32_t i = 0; i < numInstances; ++i) {
        cdf[i] = sum;
        lastCdf = sum;
        last = w[i];
        sum = sum + last;
    }
    *const_cast<float*>(scene.instIntegral) = numInstances ? lastCdf + last : 0.0f;
}

int buildLightDistributions(gfx_ctx* ctx, cudaStream_t stream, uint32_t /*bufferIndex*/) {
    SceneState &S = ctx->scene;
    const DevScene dev = ctx->devScene();
    if (!S.staticLightDistsBuilt) {
        if (S.numMeshes) {
            k_triangleImportance<<<dim3(64, S.numMeshes), 128, 0, stream>>>(dev, S.numMeshes); ctx->launches++;
            k_scanMeshes<<<(S.numMeshes + 31) / 32, 32, 0, stream>>>(dev, S.numMeshes); ctx->launches++;
        }
        if (S.numInstances) {
            k_scanInstanceGeoms<<<(S.numInstances + 63) / 64, 64, 0, stream>>>(dev, S.numInstances); ctx->launches++;
        }
        S.staticLightDistsBuilt = true;
    }
    if (S.numInstances) {
        k_instanceImportance<<<(S.numInstances + 127) / 128, 128, 0, stream>>>(dev, S.numInstances); ctx->launches++;
        k_scanInstances<<<1, 32, 0, stream>>>(dev, S.numInstances); ctx->launches++;
    }
    GFX_CUDA(ctx, cudaGetLastError());
    return GFX_OK;
}

} // namespace gfx
