// oracle/ref_tcnn/tcnn_ref.cu — TEST INFRASTRUCTURE: thin extern "C" launchers around the reference's own (vendored)
// tiny-cuda-nn device code, compiled from the sources where they lie under /root/reference/ext/tiny-cuda-nn into
// oracle/_ref/libtcnn_ref.so (recipe: oracle/ref_tcnn/Makefile; nothing of tiny-cuda-nn is copied into this repo).
//
// Purpose: pin oracle/nrc.cpp's restatement of the input encoding against the real third-party kernels on the GPU box -
//   kernel_grid<__half, 3, 2>      tiny-cuda-nn/encodings/grid.h:132-255   (hash grid, incl. grid_index / fast_hash :76-111)
//   kernel_one_blob_soa<__half>    tiny-cuda-nn/encodings/oneblob.h:110-139
// with exactly the arguments GridEncodingTemplated::forward_impl / OneBlobEncoding pass (grid.h:960-990, oneblob.h:205-216):
// linear interpolation, GridType::Hash, max_level 1000, no quantisation.  Called only by tests/tcnn_ref_check.py.
#include <tiny-cuda-nn/encodings/grid.h>
#include <tiny-cuda-nn/encodings/oneblob.h>

#include <cstdint>
#include <vector>

using namespace tcnn;

#define REF_CUDA(call) do { const cudaError_t e_ = (call); if (e_ != cudaSuccess) { fprintf(stderr, "tcnn_ref: %s: %s\n", #call, cudaGetErrorString(e_)); return 1; } } while (0)

extern "C" int tcnn_ref_grid_forward(uint32_t numElements, uint32_t numLevels, const uint32_t* hostOffsets /* numLevels + 1 */,
                                     uint32_t baseResolution, float log2PerLevelScale, const uint16_t* hostGridHalf,
                                     size_t numGridParams, const float* hostPositions /* [numElements][3] */,
                                     uint16_t* hostOutHalf /* [numElements][numLevels * 2] */) {
    uint32_t* offsets = nullptr;
    __half *grid = nullptr, *out = nullptr;
    float* positions = nullptr;
    const uint32_t numFeatures = numLevels * 2;
    REF_CUDA(cudaMalloc(&offsets, (numLevels + 1) * 4));
    REF_CUDA(cudaMalloc(&grid, numGridParams * 2));
    REF_CUDA(cudaMalloc(&positions, (size_t)numElements * 3 * 4));
    REF_CUDA(cudaMalloc(&out, (size_t)numElements * numFeatures * 2));
    REF_CUDA(cudaMemcpy(offsets, hostOffsets, (numLevels + 1) * 4, cudaMemcpyHostToDevice));
    REF_CUDA(cudaMemcpy(grid, hostGridHalf, numGridParams * 2, cudaMemcpyHostToDevice));
    REF_CUDA(cudaMemcpy(positions, hostPositions, (size_t)numElements * 3 * 4, cudaMemcpyHostToDevice));
    // positions_in(dim, i) = data[dim * stride_i + i * stride_j]: row-major [numElements][3]
    const MatrixView<const float> positionsIn(positions, 1u, 3u);
    const uint32_t threads = 512;
    const dim3 blocks((numElements + threads - 1) / threads, numLevels, 1);
    kernel_grid<__half, 3, 2><<<blocks, threads>>>(numElements, numFeatures, offsets, baseResolution, log2PerLevelScale,
                                                   0.0f /* quantize_threshold */, 1000.0f /* max_level */, nullptr,
                                                   InterpolationType::Linear, GridType::Hash, grid, positionsIn, out, nullptr);
    REF_CUDA(cudaGetLastError());
    REF_CUDA(cudaDeviceSynchronize());
    std::vector<uint16_t> soa((size_t)numElements * numFeatures);
    REF_CUDA(cudaMemcpy(soa.data(), out, soa.size() * 2, cudaMemcpyDeviceToHost));
    for (uint32_t i = 0; i < numElements; ++i)          // encoded_positions[i + feature * num_elements] -> [i][feature]
        for (uint32_t f = 0; f < numFeatures; ++f)
            hostOutHalf[(size_t)i * numFeatures + f] = soa[(size_t)f * numElements + i];
    cudaFree(offsets); cudaFree(grid); cudaFree(positions); cudaFree(out);
    return 0;
}

extern "C" int tcnn_ref_oneblob_forward(uint32_t numElements, uint32_t numBinsLog2, uint32_t numToEncode,
                                        const float* hostIn /* [numElements][numToEncode] */,
                                        uint16_t* hostOutHalf /* [numElements][numToEncode << numBinsLog2] */) {
    const uint32_t numBins = 1u << numBinsLog2, fanOut = numToEncode * numBins;
    float* in = nullptr;
    __half* out = nullptr;
    REF_CUDA(cudaMalloc(&in, (size_t)numElements * numToEncode * 4));
    REF_CUDA(cudaMalloc(&out, (size_t)numElements * fanOut * 2));
    REF_CUDA(cudaMemcpy(in, hostIn, (size_t)numElements * numToEncode * 4, cudaMemcpyHostToDevice));
    // data_in(j, i) = data[j * stride_i + i * stride_j]: row-major [numElements][numToEncode]
    const MatrixView<const float> dataIn(in, 1u, numToEncode);
    // OneBlobEncoding::forward_impl, SoA branch (oneblob.h): threads = { numToEncode, ceil(128 / numToEncode) }
    const uint32_t minThreads = 128;
    const dim3 threads(numToEncode, (minThreads + numToEncode - 1) / numToEncode, 1);
    const uint32_t blocks = (numElements + threads.y - 1) / threads.y;
    kernel_one_blob_soa<__half><<<blocks, threads>>>(numElements, numBinsLog2, numToEncode, dataIn, out);
    REF_CUDA(cudaGetLastError());
    REF_CUDA(cudaDeviceSynchronize());
    std::vector<uint16_t> soa((size_t)numElements * fanOut);
    REF_CUDA(cudaMemcpy(soa.data(), out, soa.size() * 2, cudaMemcpyDeviceToHost));
    for (uint32_t i = 0; i < numElements; ++i)          // data_out[i + (j * n_bins + k) * num_elements] -> [i][j * n_bins + k]
        for (uint32_t f = 0; f < fanOut; ++f)
            hostOutHalf[(size_t)i * fanOut + f] = soa[(size_t)f * numElements + i];
    cudaFree(in); cudaFree(out);
    return 0;
}
