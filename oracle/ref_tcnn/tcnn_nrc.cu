// oracle/ref_tcnn/tcnn_nrc.cu — TEST INFRASTRUCTURE: the reference's NRC network, built from the reference's own (vendored)
// tiny-cuda-nn sources where they lie under /root/reference/ext/tiny-cuda-nn (src/*.cu + headers; recipe:
// oracle/ref_tcnn/Makefile -> oracle/_ref/libtcnn_nrc.so; nothing of tiny-cuda-nn is copied into this repo).
//
// What it pins: NeuralRadianceCache::{initialize,infer,train} (neural_radiance_caching/network_interface.cu:48-157) is a thin
// wrapper around tcnn::NetworkWithInputEncoding + tcnn::Trainer built from a JSON config.  network_interface.cu itself cannot
// be compiled here (it includes common/common_shared.h -> basic_types.h -> OptiX SDK headers, absent), so this file hands the
// SAME config to the SAME tcnn classes; everything below the constructor calls - kernel_grid / kernel_one_blob_soa /
// kernel_mlp_fused / kernel_mlp_fused_backward / kernel_grid_backward / the split-k CUTLASS weight-gradient GEMMs /
// relative_l2_luminance_loss / Adam / EMA / Trainer's pcg32{1337} initialisation - is the reference's code, unmodified.
// It is also the one "reference GPU kernel on the same box" of this target: tcnn_nrc_time_* time tcnn's own launches.
//
// C ABI (host pointers unless stated; return 0 on success, message via tcnn_nrc_last_error()).  Called only by tests/ and
// tools/nrc_bench.py / bench.py's reference-side NRC timing.
#include <tiny-cuda-nn/config.h>
#include <tiny-cuda-nn/trainer.h>
#include <tiny-cuda-nn/network_with_input_encoding.h>

#include <cstdint>
#include <memory>
#include <string>
#include <vector>

using namespace tcnn;
using precision_t = network_precision_t;

namespace {

constexpr uint32_t kNumInputDims = 14;  // network_interface.cu:21
constexpr uint32_t kNumOutputDims = 3;  // :23

std::string g_lastError;

struct RefNrc {
    std::shared_ptr<Loss<precision_t>> loss;
    std::shared_ptr<Optimizer<precision_t>> optimizer;
    std::shared_ptr<NetworkWithInputEncoding<precision_t>> network;
    std::shared_ptr<Trainer<float, precision_t, precision_t>> trainer;
    size_t numParams = 0;
    // device staging for the host-pointer entry points and the timing loops
    GPUMemory<float> dIn, dTarget, dOut;
};

// the config of NeuralRadianceCache::initialize (network_interface.cu:48-123), key for key
json nrcConfig(uint32_t posEnc, uint32_t numHiddenLayers, float learningRate) {
    json config = {
        {"loss", {{"otype", "RelativeL2Luminance"}}},
        {"optimizer", {{"otype", "EMA"}, {"decay", 0.99f},
                       {"nested", {{"otype", "Adam"}, {"learning_rate", learningRate}, {"beta1", 0.9f}, {"beta2", 0.99f},
                                   {"l2_reg", 1e-6f}}}}},
        {"network", {{"otype", "FullyFusedMLP"}, {"n_neurons", 64}, {"n_hidden_layers", numHiddenLayers},
                     {"activation", "ReLU"}, {"output_activation", "None"}}},
    };
    json oneBlob = {{"n_dims_to_encode", 5}, {"otype", "OneBlob"}, {"n_bins", 4}};
    json identity = {{"n_dims_to_encode", 6}, {"otype", "Identity"}};
    if (posEnc == 0) { // PositionEncoding::TriangleWave
        json pos = {{"n_dims_to_encode", 3}, {"otype", "TriangleWave"}, {"n_frequencies", 12}};
        config["encoding"] = {{"otype", "Composite"}, {"nested", json::array({pos, oneBlob, identity})}};
        config["optimizer"]["nested"]["epsilon"] = 1e-8f;
    }
    else {             // PositionEncoding::HashGrid
        json pos = {{"n_dims_to_encode", 3}, {"otype", "HashGrid"}, {"per_level_scale", 2.0f}, {"log2_hashmap_size", 15},
                    {"base_resolution", 16}, {"n_levels", 16}, {"n_features_per_level", 2}};
        config["encoding"] = {{"otype", "Composite"}, {"nested", json::array({pos, oneBlob, identity})}};
        config["optimizer"]["nested"]["epsilon"] = 1e-15f;
    }
    return config;
}

// Trainer keeps [fp32 master | params | params_backward | gradients] in one allocation (trainer.h:78-84) and only exposes the
// first pointer (params()); the others follow at fixed offsets.
precision_t* trainingParams(RefNrc* n) { return (precision_t*)((char*)n->trainer->params() + sizeof(float) * n->numParams); }
precision_t* paramGradients(RefNrc* n) {
    return (precision_t*)((char*)n->trainer->params() + sizeof(float) * n->numParams + sizeof(precision_t) * n->numParams * 2);
}

template <typename F>
int guarded(F&& f) {
    try {
        f();
        return 0;
    }
    catch (const std::exception& e) {
        g_lastError = e.what();
        fprintf(stderr, "tcnn_nrc: %s\n", e.what());
        return 1;
    }
}

} // namespace

extern "C" {

const char* tcnn_nrc_last_error() { return g_lastError.c_str(); }

int tcnn_nrc_create(uint32_t posEnc, uint32_t numHiddenLayers, float learningRate, void** handle) {
    static_assert(std::is_same<precision_t, __half>::value, "the reference runs tiny-cuda-nn in half precision");
    return guarded([&] {
        auto n = std::make_unique<RefNrc>();
        json config = nrcConfig(posEnc, numHiddenLayers, learningRate);
        n->loss.reset(create_loss<precision_t>(config.value("loss", json::object())));
        n->optimizer.reset(create_optimizer<precision_t>(config.value("optimizer", json::object())));
        n->network = std::make_shared<NetworkWithInputEncoding<precision_t>>(
            kNumInputDims, kNumOutputDims, config.value("encoding", json::object()), config.value("network", json::object()));
        n->trainer = std::make_shared<Trainer<float, precision_t, precision_t>>(n->network, n->optimizer, n->loss);
        n->numParams = n->network->n_params();
        *handle = n.release();
    });
}

void tcnn_nrc_destroy(void* handle) { delete (RefNrc*)handle; }

uint32_t tcnn_nrc_num_params(void* handle) { return (uint32_t)((RefNrc*)handle)->numParams; }

// which: 0 = fp32 master weights (as float), 1 = training weights (half bits), 2 = inference / EMA weights (half bits),
//        3 = gradients of the last backward pass (half bits, loss-scaled by 128)
int tcnn_nrc_read(void* handle, int which, void* hostOut) {
    RefNrc* n = (RefNrc*)handle;
    return guarded([&] {
        CUDA_CHECK_THROW(cudaDeviceSynchronize());
        const void* src = nullptr;
        size_t bytes = n->numParams * sizeof(precision_t);
        switch (which) {
        case 0: src = n->trainer->params(); bytes = n->numParams * sizeof(float); break;
        case 1: src = trainingParams(n); break;
        case 2: src = n->optimizer->custom_weights() ? n->optimizer->custom_weights() : trainingParams(n); break;
        case 3: src = paramGradients(n); break;
        default: throw std::runtime_error{"tcnn_nrc_read: bad selector"};
        }
        CUDA_CHECK_THROW(cudaMemcpy(hostOut, src, bytes, cudaMemcpyDeviceToHost));
    });
}

// Trainer::set_params_full_precision: master <- values, training and inference weights <- half(values); the optimizer state is
// NOT reset (as in tiny-cuda-nn) - call it on a fresh network.
int tcnn_nrc_set_params(void* handle, const float* hostParams) {
    RefNrc* n = (RefNrc*)handle;
    return guarded([&] { n->trainer->set_params_full_precision(hostParams, n->numParams); });
}

// NeuralRadianceCache::infer (network_interface.cu:141-147): in [numData][14], out [numData][3] (column-major GPUMatrix)
int tcnn_nrc_infer(void* handle, const float* hostIn, uint32_t numData, float* hostOut) {
    RefNrc* n = (RefNrc*)handle;
    return guarded([&] {
        if (numData & 0x7F)
            throw std::runtime_error{"numData must be a multiple of 128."};
        n->dIn.resize((size_t)numData * kNumInputDims);
        n->dOut.resize((size_t)numData * kNumOutputDims);
        CUDA_CHECK_THROW(cudaMemcpy(n->dIn.data(), hostIn, n->dIn.get_bytes(), cudaMemcpyHostToDevice));
        GPUMatrix<float> inputs(n->dIn.data(), kNumInputDims, numData);
        GPUMatrix<float> predictions(n->dOut.data(), kNumOutputDims, numData);
        n->network->inference(nullptr, inputs, predictions);
        CUDA_CHECK_THROW(cudaDeviceSynchronize());
        CUDA_CHECK_THROW(cudaMemcpy(hostOut, n->dOut.data(), n->dOut.get_bytes(), cudaMemcpyDeviceToHost));
    });
}

// NeuralRadianceCache::train (network_interface.cu:149-157): one Trainer::training_step (forward, loss, backward, optimizer)
int tcnn_nrc_train(void* handle, const float* hostIn, const float* hostTarget, uint32_t numData, float* lossOut) {
    RefNrc* n = (RefNrc*)handle;
    return guarded([&] {
        if (numData & 0x7F)
            throw std::runtime_error{"numData must be a multiple of 128."};
        n->dIn.resize((size_t)numData * kNumInputDims);
        n->dTarget.resize((size_t)numData * kNumOutputDims);
        CUDA_CHECK_THROW(cudaMemcpy(n->dIn.data(), hostIn, n->dIn.get_bytes(), cudaMemcpyHostToDevice));
        CUDA_CHECK_THROW(cudaMemcpy(n->dTarget.data(), hostTarget, n->dTarget.get_bytes(), cudaMemcpyHostToDevice));
        GPUMatrix<float> inputs(n->dIn.data(), kNumInputDims, numData);
        GPUMatrix<float> targets(n->dTarget.data(), kNumOutputDims, numData);
        auto context = n->trainer->training_step(nullptr, inputs, targets);
        if (lossOut)
            *lossOut = n->trainer->loss(nullptr, *context);
        CUDA_CHECK_THROW(cudaDeviceSynchronize());
    });
}

// The first half of a training step only (Trainer::forward + Trainer::backward with training_step's loss scale of 128,
// trainer.h:178-197), no optimizer step: leaves the loss-scaled half gradients for tcnn_nrc_read(handle, 3, ...).
int tcnn_nrc_forward_backward(void* handle, const float* hostIn, const float* hostTarget, uint32_t numData, float* lossOut,
                              uint16_t* hostOutputHalf /* optional: [numData][16] network output */) {
    RefNrc* n = (RefNrc*)handle;
    return guarded([&] {
        n->dIn.resize((size_t)numData * kNumInputDims);
        n->dTarget.resize((size_t)numData * kNumOutputDims);
        CUDA_CHECK_THROW(cudaMemcpy(n->dIn.data(), hostIn, n->dIn.get_bytes(), cudaMemcpyHostToDevice));
        CUDA_CHECK_THROW(cudaMemcpy(n->dTarget.data(), hostTarget, n->dTarget.get_bytes(), cudaMemcpyHostToDevice));
        GPUMatrix<float> inputs(n->dIn.data(), kNumInputDims, numData);
        GPUMatrix<float> targets(n->dTarget.data(), kNumOutputDims, numData);
        auto context = n->trainer->forward(nullptr, 128.0f, inputs, targets);
        n->trainer->backward(nullptr, *context, inputs);
        CUDA_CHECK_THROW(cudaDeviceSynchronize());
        if (lossOut)
            *lossOut = n->trainer->loss(nullptr, *context);
        if (hostOutputHalf)
            CUDA_CHECK_THROW(cudaMemcpy(hostOutputHalf, context->output.data(), (size_t)numData * 16 * sizeof(precision_t),
                                        cudaMemcpyDeviceToHost));
    });
}

// Timing of the reference's own launches on this GPU: `iters` back-to-back calls between two CUDA events on the default stream,
// after `warmup` untimed calls; inputs (and targets) are uploaded once and stay resident.  msPerCall = elapsed / iters.
int tcnn_nrc_time_infer(void* handle, const float* hostIn, uint32_t numData, uint32_t warmup, uint32_t iters, float* msPerCall) {
    RefNrc* n = (RefNrc*)handle;
    return guarded([&] {
        n->dIn.resize((size_t)numData * kNumInputDims);
        n->dOut.resize((size_t)numData * kNumOutputDims);
        CUDA_CHECK_THROW(cudaMemcpy(n->dIn.data(), hostIn, n->dIn.get_bytes(), cudaMemcpyHostToDevice));
        GPUMatrix<float> inputs(n->dIn.data(), kNumInputDims, numData);
        GPUMatrix<float> predictions(n->dOut.data(), kNumOutputDims, numData);
        for (uint32_t i = 0; i < warmup; ++i)
            n->network->inference(nullptr, inputs, predictions);
        cudaEvent_t e0, e1;
        CUDA_CHECK_THROW(cudaEventCreate(&e0));
        CUDA_CHECK_THROW(cudaEventCreate(&e1));
        CUDA_CHECK_THROW(cudaDeviceSynchronize());
        CUDA_CHECK_THROW(cudaEventRecord(e0, nullptr));
        for (uint32_t i = 0; i < iters; ++i)
            n->network->inference(nullptr, inputs, predictions);
        CUDA_CHECK_THROW(cudaEventRecord(e1, nullptr));
        CUDA_CHECK_THROW(cudaEventSynchronize(e1));
        float ms = 0.0f;
        CUDA_CHECK_THROW(cudaEventElapsedTime(&ms, e0, e1));
        *msPerCall = ms / (float)iters;
        cudaEventDestroy(e0);
        cudaEventDestroy(e1);
    });
}

int tcnn_nrc_time_train(void* handle, const float* hostIn, const float* hostTarget, uint32_t numData, uint32_t warmup,
                        uint32_t iters, float* msPerCall) {
    RefNrc* n = (RefNrc*)handle;
    return guarded([&] {
        n->dIn.resize((size_t)numData * kNumInputDims);
        n->dTarget.resize((size_t)numData * kNumOutputDims);
        CUDA_CHECK_THROW(cudaMemcpy(n->dIn.data(), hostIn, n->dIn.get_bytes(), cudaMemcpyHostToDevice));
        CUDA_CHECK_THROW(cudaMemcpy(n->dTarget.data(), hostTarget, n->dTarget.get_bytes(), cudaMemcpyHostToDevice));
        GPUMatrix<float> inputs(n->dIn.data(), kNumInputDims, numData);
        GPUMatrix<float> targets(n->dTarget.data(), kNumOutputDims, numData);
        for (uint32_t i = 0; i < warmup; ++i)
            n->trainer->training_step(nullptr, inputs, targets);
        cudaEvent_t e0, e1;
        CUDA_CHECK_THROW(cudaEventCreate(&e0));
        CUDA_CHECK_THROW(cudaEventCreate(&e1));
        CUDA_CHECK_THROW(cudaDeviceSynchronize());
        CUDA_CHECK_THROW(cudaEventRecord(e0, nullptr));
        for (uint32_t i = 0; i < iters; ++i)
            n->trainer->training_step(nullptr, inputs, targets);
        CUDA_CHECK_THROW(cudaEventRecord(e1, nullptr));
        CUDA_CHECK_THROW(cudaEventSynchronize(e1));
        float ms = 0.0f;
        CUDA_CHECK_THROW(cudaEventElapsedTime(&ms, e0, e1));
        *msPerCall = ms / (float)iters;
        cudaEventDestroy(e0);
        cudaEventDestroy(e1);
    });
}

} // extern "C"
