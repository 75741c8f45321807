// oracle/nrc.cpp — TEST INFRASTRUCTURE (CPU oracle).  Not part of the product path.
//
// Scalar restatement of the NRC network the reference builds through tiny-cuda-nn
// (neural_radiance_caching/network_interface.cu:48-157; ext/tiny-cuda-nn @ be460660, vendored source):
//   Composite encoding       HashGrid(3 dims, 16 levels, F=2, T=2^15, base 16, scale 2)
//                            + OneBlob(5 dims x 4 bins) + Identity(6 dims), padded 58 -> 64 with ones
//     kernel_grid<half,3,2>  encodings/grid.h:132-304 ; grid_index/fast_hash :76-111 ; level sizing :885-922
//     pos_fract              common_device.h:425-431 ; quartic_cdf :478-483 ; kernel_one_blob_soa oneblob.h:110-139
//   FullyFusedMLP            64 neurons, n hidden layers, ReLU, linear output 3 (padded 16)
//                            src/fully_fused_mlp.cu:47-129 (forward), :150-259,:499-557 (backward)
//   RelativeL2Luminance      losses/relative_l2_luminance.h:41-88, loss scale 128 (trainer.h:187)
//   Adam + EMA(0.99)         optimizers/adam.h:49-115, optimizers/ema.h:46-121 (half-precision EMA)
// Arithmetic model: fp16 storage of weights/activations/gradients exactly where tcnn stores halves.  Two accumulation modes
// for the fully-connected layers (orc_nrc_set_accumulate_half):
//   0 (default)  fp32 accumulation of each 64-term dot product, one rounding to half per layer - what this repo's tcgen05
//                kernels compute (fp16 operands, fp32 TMEM accumulators);
//   1            tiny-cuda-nn's own model: kernel_mlp_fused keeps its wmma accumulator fragments in __half
//                (fully_fused_mlp.cu:68, backward :198), i.e. the running sum is rounded to half after every 16-wide k step
//                of the 16x16x16 HMMA (the order INSIDE one HMMA is hardware-defined: compared at 1e-3 relative L2).
// Pinned against the reference's own code on the GPU box: the encoding bit for bit (oracle/ref_tcnn/tcnn_ref.cu,
// tests/test_gpu_tcnn_ref.py), the network forward / one training step's gradients / the weights after four steps / the
// pcg32{1337} initial parameters against tiny-cuda-nn's NetworkWithInputEncoding + Trainer built from network_interface.cu's
// config (oracle/ref_tcnn/tcnn_nrc.cu, tests/tcnn_nrc_check.py, tests/test_gpu_tcnn_nrc.py).  CPU-side it is pinned by
// finite-difference gradient checks and loss descent in tests/test_oracle_nrc.py.
// Parameter layout (this repo's own, shared with the CUDA library): MLP matrices [out][in] in layer
// order (first 64x64, hidden 64x64 ..., last 16x64), then the hash-grid table level by level.
#include <cstdint>
#include <cstring>
#include <cmath>
#include <vector>
#include <algorithm>
#include <omp.h>

namespace {

typedef _Float16 half;
inline float h2f(half h) { return (float)h; }
inline half f2h(float f) { return (half)f; }

constexpr uint32_t kInputDims = 14, kOutputDims = 3, kWidth = 64, kPaddedOutput = 16;
constexpr uint32_t kLevels = 16, kFeaturesPerLevel = 2, kLog2HashmapSize = 15, kBaseResolution = 16;
constexpr float kPerLevelScale = 2.0f;
constexpr float kLossScale = 128.0f;

struct Level {
    uint32_t offset;       // in grid entries (not features)
    uint32_t hashmapSize;
    float scale;
    uint32_t resolution;
};

inline float quarticCdf(float x, float invRadius) { // common_device.h:478-483
    const float u = x * invRadius;
    const float u2 = u * u;
    const float u4 = u2 * u2;
    return std::fmax(0.0f, std::fmin(1.0f, (15.0f / 16.0f) * u * (1 - (2.0f / 3.0f) * u2 + (1.0f / 5.0f) * u4) + 0.5f));
}

} // namespace

struct orc_nrc {
    uint32_t numHiddenLayers;
    float learningRate;
    uint32_t numMatrixWeights, numParams;
    Level levels[kLevels];
    std::vector<half> params;       // current fp16 weights
    std::vector<half> paramsEma;    // inference weights (EMA)
    std::vector<float> master;      // fp32 master weights
    std::vector<float> m1, m2;      // Adam moments
    std::vector<uint32_t> steps;    // per-parameter step counters (adam.h:108)
    uint32_t globalStep = 0;
    int accumulateHalf = 0;         // see the header: 1 = tiny-cuda-nn's half accumulator fragments
    std::vector<float> lastGradients; // loss-scaled gradients of the last training step, rounded to half like tcnn's buffer
};

// one output of a fully-connected layer: sum_i w[i * stride] * x[i]
static inline float dotLayer(const orc_nrc* n, const half* w, size_t stride, const half* x, uint32_t count) {
    if (!n->accumulateHalf) {
        float acc = 0.0f;
        for (uint32_t i = 0; i < count; ++i)
            acc += h2f(w[i * stride]) * h2f(x[i]);
        return acc;
    }
    half acc = f2h(0.0f);
    for (uint32_t c = 0; c < count; c += 16) {
        float partial = 0.0f;
        for (uint32_t i = c; i < c + 16 && i < count; ++i)
            partial += h2f(w[i * stride]) * h2f(x[i]);
        acc = f2h(h2f(acc) + partial);
    }
    return h2f(acc);
}

static void setupLevels(orc_nrc* n) { // grid.h:885-922
    uint32_t offset = 0;
    for (uint32_t l = 0; l < kLevels; ++l) {
        const float scale = std::exp2(l * std::log2(kPerLevelScale)) * kBaseResolution - 1.0f;
        const uint32_t resolution = (uint32_t)std::ceil(scale) + 1;
        const double dense = std::pow((double)resolution, 3.0);
        uint32_t paramsInLevel = dense > (double)(0xFFFFFFFFu / 2) ? 0xFFFFFFFFu / 2 : resolution * resolution * resolution;
        paramsInLevel = (paramsInLevel + 7u) / 8u * 8u;
        paramsInLevel = std::min(paramsInLevel, 1u << kLog2HashmapSize);
        n->levels[l] = Level{ offset, paramsInLevel, scale, resolution };
        offset += paramsInLevel;
    }
    n->numMatrixWeights = kWidth * kWidth * n->numHiddenLayers + kPaddedOutput * kWidth;
    n->numParams = n->numMatrixWeights + offset * kFeaturesPerLevel;
}

static inline uint32_t gridIndex(const Level &lv, const uint32_t pos[3]) { // grid.h:95-111 (feature 0, GridType::Hash)
    uint32_t stride = 1;
    uint32_t index = 0;
    for (uint32_t dim = 0; dim < 3 && stride <= lv.hashmapSize; ++dim) {
        index += pos[dim] * stride;
        stride *= lv.resolution;
    }
    if (lv.hashmapSize < stride) // fast_hash :76-91
        index = (pos[0] * 1u) ^ (pos[1] * 2654435761u) ^ (pos[2] * 805459861u);
    return (index % lv.hashmapSize) * kFeaturesPerLevel;
}

// encoding of one query into 64 halves; optionally records the trilinear stencil for the backward pass
struct GridStencil { uint32_t index[kLevels][8]; float weight[kLevels][8]; };

static void encode(const orc_nrc* n, const half* table, const float* in, half* out, GridStencil* st) {
    // HashGrid (kernel_grid<half,3,2>, grid.h:132-255)
    for (uint32_t l = 0; l < kLevels; ++l) {
        const Level &lv = n->levels[l];
        const half* grid = table + (size_t)lv.offset * kFeaturesPerLevel;
        float pos[3];
        uint32_t posGrid[3];
        for (uint32_t d = 0; d < 3; ++d) { // pos_fract, common_device.h:425-431
            pos[d] = in[d] * lv.scale + 0.5f;
            const int tmp = (int)std::floor(pos[d]);
            posGrid[d] = (uint32_t)tmp;
            pos[d] -= (float)tmp;
        }
        half result[2] = { f2h(0.0f), f2h(0.0f) };
        for (uint32_t idx = 0; idx < 8; ++idx) {
            float weight = 1;
            uint32_t local[3];
            for (uint32_t d = 0; d < 3; ++d) {
                if ((idx & (1u << d)) == 0) {
                    weight *= 1 - pos[d];
                    local[d] = posGrid[d];
                }
                else {
                    weight *= pos[d];
                    local[d] = posGrid[d] + 1;
                }
            }
            const uint32_t gi = gridIndex(lv, local);
            if (st) {
                st->index[l][idx] = lv.offset * kFeaturesPerLevel + gi;
                st->weight[l][idx] = weight;
            }
            for (uint32_t f = 0; f < 2; ++f) {
                const float data = h2f(grid[gi + f]);
                result[f] = f2h(h2f(result[f]) + h2f(f2h(weight * data)));
            }
        }
        out[l * 2 + 0] = result[0];
        out[l * 2 + 1] = result[1];
    }
    // OneBlob, 5 dims x 4 bins.  The Composite encoding hands every nested encoding an SoA slice (all three prefer SoA:
    // composite.h:253-262, grid.h:1257, oneblob.h:299, identity.h:173), so OneBlobEncoding::forward_impl takes its SoA branch
    // (oneblob.h:219-232) and the kernel is kernel_one_blob_soa (oneblob.h:110-139): the CDF at each of the five bin
    // boundaries k / 4 is summed over the three periodic images, and a bin is the difference of consecutive boundaries.
    // (The AoS kernel, :83-108 with one_blob_subwarp_aligned :47-69, closes the last bin with "first boundary + 1" instead,
    // which differs in the last bit of ~0.3 % of the last-bin features.)
    for (uint32_t d = 0; d < 5; ++d) {
        const float x = in[3 + d];
        float leftCdf = quarticCdf(-x, 4.0f) + quarticCdf(-x - 1.0f, 4.0f) + quarticCdf(-x + 1.0f, 4.0f);
        for (uint32_t k = 0; k < 4; ++k) {
            const float rightBoundary = std::scalbn((float)(k + 1), -2);
            const float rightCdf = quarticCdf(rightBoundary - x, 4.0f) + quarticCdf(rightBoundary - x - 1.0f, 4.0f) +
                                   quarticCdf(rightBoundary - x + 1.0f, 4.0f);
            out[32 + d * 4 + k] = f2h(rightCdf - leftCdf);
            leftCdf = rightCdf;
        }
    }
    // Identity, 6 dims (identity.h), then padding with ones (oneblob.h:102-104 style "bias-like" pad)
    for (uint32_t d = 0; d < 6; ++d)
        out[52 + d] = f2h(in[8 + d]);
    for (uint32_t k = 58; k < 64; ++k)
        out[k] = f2h(1.0f);
}

// forward through the MLP; acts[layer] receives the post-activation halves of every hidden layer
static void mlpForward(const orc_nrc* n, const half* w, const half* x, half* acts /*numHidden*64*/, half* out /*16*/) {
    const half* in = x;
    for (uint32_t layer = 0; layer < n->numHiddenLayers; ++layer) {
        const half* W = w + (size_t)layer * kWidth * kWidth;
        half* o = acts + (size_t)layer * kWidth;
        for (uint32_t j = 0; j < kWidth; ++j) {
            const float acc = dotLayer(n, W + j * kWidth, 1, in, kWidth);
            o[j] = f2h(std::fmax(acc, 0.0f)); // ReLU, stored as half (fully_fused_mlp.cu:93-112)
        }
        in = o;
    }
    const half* W = w + (size_t)n->numHiddenLayers * kWidth * kWidth;
    for (uint32_t j = 0; j < kPaddedOutput; ++j)
        out[j] = f2h(dotLayer(n, W + j * kWidth, 1, in, kWidth));
}

extern "C" {

orc_nrc* orc_nrc_create(uint32_t numHiddenLayers, float learningRate) {
    orc_nrc* n = new orc_nrc();
    n->numHiddenLayers = numHiddenLayers;
    n->learningRate = learningRate;
    setupLevels(n);
    n->params.assign(n->numParams, f2h(0.0f));
    n->paramsEma.assign(n->numParams, f2h(0.0f));
    n->master.assign(n->numParams, 0.0f);
    n->m1.assign(n->numParams, 0.0f);
    n->m2.assign(n->numParams, 0.0f);
    n->steps.assign(n->numParams, 0u);
    return n;
}
void orc_nrc_destroy(orc_nrc* n) { delete n; }
uint32_t orc_nrc_num_params(orc_nrc* n) { return n->numParams; }
uint32_t orc_nrc_num_matrix_weights(orc_nrc* n) { return n->numMatrixWeights; }

// installs the same fp16 values as current weights, EMA (inference) weights and fp32 master weights;
// resets the optimizer state
void orc_nrc_set_params(orc_nrc* n, const uint16_t* halfBits) {
    std::memcpy(n->params.data(), halfBits, (size_t)n->numParams * 2);
    n->paramsEma = n->params;
    for (uint32_t i = 0; i < n->numParams; ++i)
        n->master[i] = h2f(n->params[i]);
    std::fill(n->m1.begin(), n->m1.end(), 0.0f);
    std::fill(n->m2.begin(), n->m2.end(), 0.0f);
    std::fill(n->steps.begin(), n->steps.end(), 0u);
    n->globalStep = 0;
}
void orc_nrc_set_accumulate_half(orc_nrc* n, int on) { n->accumulateHalf = on; }
// the state of a freshly constructed tcnn::Trainer (trainer.h:72-109, ema.h:88-100): fp32 master = the given values, training
// weights = their halves, inference (EMA) weights and optimizer state zero
void orc_nrc_init_master(orc_nrc* n, const float* master) {
    for (uint32_t i = 0; i < n->numParams; ++i) {
        n->master[i] = master[i];
        n->params[i] = f2h(master[i]);
        n->paramsEma[i] = f2h(0.0f);
    }
    std::fill(n->m1.begin(), n->m1.end(), 0.0f);
    std::fill(n->m2.begin(), n->m2.end(), 0.0f);
    std::fill(n->steps.begin(), n->steps.end(), 0u);
    n->globalStep = 0;
}
void orc_nrc_get_master(orc_nrc* n, float* out) { std::memcpy(out, n->master.data(), (size_t)n->numParams * 4); }
// loss-scaled (x128) gradients of the last orc_nrc_train, as the halves tcnn's gradient buffer would hold, widened to float
int orc_nrc_get_gradients(orc_nrc* n, float* out) {
    if (n->lastGradients.size() != n->numParams)
        return 1;
    std::memcpy(out, n->lastGradients.data(), (size_t)n->numParams * 4);
    return 0;
}
void orc_nrc_get_params(orc_nrc* n, uint16_t* out, int ema) {
    std::memcpy(out, ema ? n->paramsEma.data() : n->params.data(), (size_t)n->numParams * 2);
}

void orc_nrc_encode(orc_nrc* n, const float* in, uint32_t num, uint16_t* outHalfBits, int ema) {
    const half* table = (ema ? n->paramsEma.data() : n->params.data()) + n->numMatrixWeights;
#pragma omp parallel for schedule(static)
    for (int64_t q = 0; q < (int64_t)num; ++q)
        encode(n, table, in + (size_t)q * kInputDims, reinterpret_cast<half*>(outHalfBits) + (size_t)q * 64, nullptr);
}

// NeuralRadianceCache::infer (network_interface.cu:141-147): inference runs on the EMA weights (trainer.h:89-93)
void orc_nrc_infer(orc_nrc* n, const float* in, float* out, uint32_t num) {
    const half* w = n->paramsEma.data();
    const half* table = w + n->numMatrixWeights;
#pragma omp parallel for schedule(static)
    for (int64_t q = 0; q < (int64_t)num; ++q) {
        half x[64], acts[64 * 8], o[16];
        encode(n, table, in + (size_t)q * kInputDims, x, nullptr);
        mlpForward(n, w, x, acts, o);
        for (uint32_t k = 0; k < kOutputDims; ++k)
            out[(size_t)q * kOutputDims + k] = h2f(o[k]); // trim_and_cast_from
    }
}

// Trainer::training_step (trainer.h:178-197): forward, RelativeL2Luminance, backward, Adam, EMA. Returns the loss.
float orc_nrc_train(orc_nrc* n, const float* in, const float* target, uint32_t num) {
    const uint32_t H = n->numHiddenLayers;
    std::vector<float> grad(n->numParams, 0.0f); // fp32 accumulation of the loss-scaled gradients
    double lossSum = 0.0;
    const half* w = n->params.data();
    const half* table = w + n->numMatrixWeights;
    const int nt = omp_get_max_threads();
    std::vector<std::vector<float>> tgrad(nt);
#pragma omp parallel
    {
        std::vector<float> &g = tgrad[omp_get_thread_num()];
        g.assign(n->numParams, 0.0f);
        double localLoss = 0.0;
#pragma omp for schedule(static)
        for (int64_t q = 0; q < (int64_t)num; ++q) {
            half x[64], acts[64 * 8], o[16];
            GridStencil st;
            encode(n, table, in + (size_t)q * kInputDims, x, &st);
            mlpForward(n, w, x, acts, o);
            // relative_l2_luminance_loss (relative_l2_luminance.h:41-88), stride 16, dims 3
            const uint32_t nTotal = num * kOutputDims;
            const float r = h2f(o[0]), gch = h2f(o[1]), b = h2f(o[2]);
            const float luminance = 0.299f * r + 0.587f * gch + 0.114f * b;
            const float denom = luminance * luminance + 0.01f;
            half dOut[16];
            for (uint32_t k = 0; k < kPaddedOutput; ++k) {
                if (k >= kOutputDims) {
                    dOut[k] = f2h(0.0f);
                    continue;
                }
                const float difference = h2f(o[k]) - target[(size_t)q * kOutputDims + k];
                localLoss += difference * difference / denom / nTotal;
                const float gradient = 2 * difference / denom;
                dOut[k] = f2h(kLossScale * gradient / nTotal);
            }
            // backward through the layers (fully_fused_mlp.cu:150-259, 773-900): dW = dY^T X, dX = W^T dY
            half dCur[64];
            {
                const half* W = w + (size_t)H * kWidth * kWidth;
                const half* hin = H > 0 ? acts + (size_t)(H - 1) * kWidth : x;
                float* gW = g.data() + (size_t)H * kWidth * kWidth;
                for (uint32_t j = 0; j < kPaddedOutput; ++j) {
                    const float dj = h2f(dOut[j]);
                    if (dj != 0.0f)
                        for (uint32_t i = 0; i < kWidth; ++i)
                            gW[j * kWidth + i] += dj * h2f(hin[i]);
                }
                for (uint32_t i = 0; i < kWidth; ++i) {
                    const float acc = dotLayer(n, W + i, kWidth, dOut, kPaddedOutput);
                    // ReLU backward on the stored forward activation (fully_fused_mlp.cu:163-176)
                    dCur[i] = (H > 0 && !(h2f(hin[i]) > 0.0f)) ? f2h(0.0f) : f2h(acc);
                }
            }
            for (int layer = (int)H - 1; layer >= 0; --layer) {
                const half* W = w + (size_t)layer * kWidth * kWidth;
                const half* hin = layer > 0 ? acts + (size_t)(layer - 1) * kWidth : x;
                float* gW = g.data() + (size_t)layer * kWidth * kWidth;
                for (uint32_t j = 0; j < kWidth; ++j) {
                    const float dj = h2f(dCur[j]);
                    if (dj != 0.0f)
                        for (uint32_t i = 0; i < kWidth; ++i)
                            gW[j * kWidth + i] += dj * h2f(hin[i]);
                }
                half dNext[64];
                for (uint32_t i = 0; i < kWidth; ++i) {
                    const float acc = dotLayer(n, W + i, kWidth, dCur, kWidth);
                    dNext[i] = (layer > 0 && !(h2f(hin[i]) > 0.0f)) ? f2h(0.0f) : f2h(acc);
                }
                std::memcpy(dCur, dNext, sizeof(dCur));
            }
            // kernel_grid_backward (grid.h:306-429): scatter dL/dy of the 32 grid features
            float* gGrid = g.data() + n->numMatrixWeights;
            for (uint32_t l = 0; l < kLevels; ++l)
                for (uint32_t idx = 0; idx < 8; ++idx)
                    for (uint32_t f = 0; f < 2; ++f)
                        gGrid[st.index[l][idx] + f] += st.weight[l][idx] * h2f(dCur[l * 2 + f]);
        }
#pragma omp critical
        lossSum += localLoss;
    }
    for (int t = 0; t < nt; ++t)
        for (uint32_t i = 0; i < n->numParams; ++i)
            grad[i] += tgrad[t][i];

    n->lastGradients.resize(n->numParams);
    for (uint32_t i = 0; i < n->numParams; ++i)
        n->lastGradients[i] = h2f(f2h(grad[i]));
    // Adam (adam.h:49-115) on half gradients, then EMA (ema.h:61-77,103-121)
    ++n->globalStep;
    const float beta1 = 0.9f, beta2 = 0.99f, epsilon = 1e-15f, l2Reg = 1e-6f;
    const float emaDecay = 0.99f;
    const float emaDebiasOld = 1 - (float)std::pow(emaDecay, (double)(n->globalStep - 1));
    const float emaDebiasNew = 1.0f / (1 - (float)std::pow(emaDecay, (double)n->globalStep));
#pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < (int64_t)n->numParams; ++i) {
        float gradient = h2f(f2h(grad[i])) / kLossScale; // gradients are stored as half
        const bool matrix = (uint32_t)i < n->numMatrixWeights;
        bool update = true;
        if (!matrix && gradient == 0)
            update = false;
        if (update) {
            const float weightFp = n->master[i];
            if (matrix)
                gradient += l2Reg * weightFp;
            const float gradientSq = gradient * gradient;
            const float firstMoment = n->m1[i] = beta1 * n->m1[i] + (1 - beta1) * gradient;
            const float secondMoment = n->m2[i] = beta2 * n->m2[i] + (1 - beta2) * gradientSq;
            float lr = n->learningRate;
            const uint32_t currentStep = ++n->steps[i];
            lr *= std::sqrt(1 - std::pow(beta2, (float)currentStep)) / (1 - std::pow(beta1, (float)currentStep));
            const float effectiveLr = std::fmin(std::fmax(lr / (std::sqrt(secondMoment) + epsilon), 0.0f), 3.402823466e+38f);
            const float newWeight = weightFp - effectiveLr * firstMoment;
            n->master[i] = newWeight;
            n->params[i] = f2h(newWeight);
        }
        const float filtered = (h2f(n->paramsEma[i]) * emaDecay * emaDebiasOld + h2f(n->params[i]) * (1 - emaDecay)) * emaDebiasNew;
        n->paramsEma[i] = f2h(filtered);
    }
    return (float)lossSum;
}

} // extern "C"
