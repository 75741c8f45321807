// gfx_headless.cpp — a headless C++ host on top of the C ABI of include/gfxb200.h.
//
// The reference hosts are C++ programs (restir_di/restir_di_main.cpp, path_tracing/path_tracing_main.cpp,
// regir/regir_main.cpp, neural_radiance_caching/neural_radiance_caching_main.cpp) whose frame loops end in OptiX /
// CUDA launches.  This file is those frame loops with the window, UI and asset import removed and every launch line
// replaced by the gfx_* call that INTEGRATION.md names for it; it exists to show (and to test, see
// tests/test_gpu_host_cpp.py) that the boundary is a plain C ABI: nothing here knows about Python, torch or CUDA
// types.  Scene and default parameters come from a flat binary file written by gfxexp_b200/scenes.py::save_scene_bin.
//
//   gfx_headless <scene.bin> <renderer> <width> <height> <frames> <out.raw> [nrc_params.f16]
//   renderer: restir | restir_unbiased | rearch | rearch_unbiased | pathtrace | regir | nrc
// Output: the accumulated beauty buffer (float4 per pixel, row-major) after <frames> frames, plus one line of JSON
// with the mean milliseconds per frame (wall clock around a synchronised loop).
#include "gfxb200.h"
#include "png_writer.h"

#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <random>
#include <stdexcept>
#include <string>
#include <vector>

namespace {

void check(gfx_ctx* ctx, int rc, const char* what) {
    if (rc != GFX_OK)
        throw std::runtime_error(std::string(what) + " failed (" + std::to_string(rc) + "): " +
                                 (ctx ? gfx_last_error_string(ctx) : "no context"));
}

struct SceneFile {
    std::vector<std::vector<float>> positions, normals, tangents, texcoords;
    std::vector<std::vector<uint32_t>> triangles;
    std::vector<GfxMeshDesc> meshes;
    std::vector<GfxMaterialDesc> materials;
    std::vector<GfxInstanceDesc> instances;
    std::vector<uint32_t> slots;
    GfxFrameParams defaults;
    GfxSceneDesc desc = {};
};

template <typename T>
void readInto(FILE* f, T* dst, size_t count) {
    if (count && fread(dst, sizeof(T), count, f) != count)
        throw std::runtime_error("scene file truncated");
}

void loadScene(const char* path, uint32_t width, uint32_t height, SceneFile* s) {
    FILE* f = fopen(path, "rb");
    if (!f)
        throw std::runtime_error(std::string("cannot open ") + path);
    char magic[4];
    uint32_t hdr[5];
    readInto(f, magic, 4);
    readInto(f, hdr, 5);
    if (memcmp(magic, "GFXS", 4) != 0 || hdr[4] != sizeof(GfxFrameParams))
        throw std::runtime_error("not a GFXS scene file of this ABI version");
    const uint32_t numMeshes = hdr[0], numMaterials = hdr[1], numInstances = hdr[2], numSlots = hdr[3];
    s->positions.resize(numMeshes); s->normals.resize(numMeshes); s->tangents.resize(numMeshes);
    s->texcoords.resize(numMeshes); s->triangles.resize(numMeshes); s->meshes.resize(numMeshes);
    for (uint32_t m = 0; m < numMeshes; ++m) {
        uint32_t mh[3];
        readInto(f, mh, 3);
        s->positions[m].resize(3 * (size_t)mh[0]); s->normals[m].resize(3 * (size_t)mh[0]);
        s->tangents[m].resize(3 * (size_t)mh[0]); s->texcoords[m].resize(2 * (size_t)mh[0]);
        s->triangles[m].resize(3 * (size_t)mh[1]);
        readInto(f, s->positions[m].data(), s->positions[m].size());
        readInto(f, s->normals[m].data(), s->normals[m].size());
        readInto(f, s->tangents[m].data(), s->tangents[m].size());
        readInto(f, s->texcoords[m].data(), s->texcoords[m].size());
        readInto(f, s->triangles[m].data(), s->triangles[m].size());
        GfxMeshDesc &d = s->meshes[m];
        d.positions = s->positions[m].data(); d.normals = s->normals[m].data(); d.tangents = s->tangents[m].data();
        d.texcoords = s->texcoords[m].data(); d.triangles = s->triangles[m].data();
        d.numVertices = mh[0]; d.numTriangles = mh[1]; d.materialSlot = mh[2];
    }
    s->materials.resize(numMaterials); s->instances.resize(numInstances); s->slots.resize(numSlots);
    readInto(f, s->materials.data(), numMaterials);
    readInto(f, s->instances.data(), numInstances);
    readInto(f, s->slots.data(), numSlots);
    readInto(f, &s->defaults, 1);
    fclose(f);
    // the file carries the parameter defaults for a square image; aspect follows the requested size
    s->defaults.camera.aspect = (float)width / (float)height;
    s->defaults.prevCamera.aspect = s->defaults.camera.aspect;
    s->desc.meshes = s->meshes.data(); s->desc.materials = s->materials.data(); s->desc.instances = s->instances.data();
    s->desc.instanceMeshSlots = s->slots.data();
    s->desc.numMeshes = numMeshes; s->desc.numMaterials = numMaterials; s->desc.numInstances = numInstances;
    s->desc.numInstanceMeshSlots = numSlots;
}

} // namespace

int main(int argc, char** argv) {
    if (argc < 7) {
        fprintf(stderr, "usage: %s <scene.bin> <renderer> <width> <height> <frames> <out.raw | out.png> [nrc_params.f16]\n", argv[0]);
        return 2;
    }
    const std::string renderer = argv[2];
    const uint32_t width = (uint32_t)atoi(argv[3]), height = (uint32_t)atoi(argv[4]);
    const int frames = atoi(argv[5]);
    gfx_ctx* gfx = nullptr;
    gfx_nrc* nrc = nullptr;
    try {
        SceneFile scene;
        loadScene(argv[1], width, height, &scene);
        check(nullptr, gfx_ctx_create(0, &gfx), "gfx_ctx_create");                        // GPUEnvironment::initialize
        check(gfx, gfx_scene_upload(gfx, &scene.desc), "gfx_scene_upload");                // Scene::initialize + create*
        check(gfx, gfx_bvh_build(gfx, nullptr, 0), "gfx_bvh_build");                       // Scene::updateASs
        check(gfx, gfx_frame_create(gfx, width, height), "gfx_frame_create");
        check(gfx, gfx_rng_seed(gfx, 591842031321323413ull), "gfx_rng_seed");              // restir_di_main.cpp:1309-1321
        check(gfx, gfx_restir_setup_neighbor_table(gfx), "gfx_restir_setup_neighbor_table");

        const bool unbiased = renderer == "restir_unbiased" || renderer == "rearch_unbiased";
        if (renderer == "nrc") {
            check(gfx, gfx_nrc_create(gfx, 2, 1e-2f, &nrc), "gfx_nrc_create");             // neuralRadianceCache.initialize
            if (argc >= 8) {
                FILE* pf = fopen(argv[7], "rb");
                if (!pf)
                    throw std::runtime_error("cannot open the NRC parameter file");
                std::vector<uint16_t> halfParams(gfx_nrc_num_params(nrc));
                readInto(pf, halfParams.data(), halfParams.size());
                fclose(pf);
                check(gfx, gfx_nrc_set_params(nrc, halfParams.data(), halfParams.size() * 2), "gfx_nrc_set_params");
            }
        }
        std::mt19937 perFrameRng(72139121);                                                // neural_radiance_caching_main.cpp:1602

        GfxFrameParams fp = scene.defaults;
        uint32_t lastSpatialNeighborBaseIndex = 0, lastReservoirIndex = 1;                 // restir_di_main.cpp:1685-1686
        const uint32_t numSpatialReusePasses = 1;
        double totalMs = 0.0;
        for (int frameIndex = 0; frameIndex < frames; ++frameIndex) {
            check(gfx, gfx_synchronize(gfx, nullptr), "gfx_synchronize");
            const auto t0 = std::chrono::steady_clock::now();
            const bool newSequence = frameIndex == 0;
            const uint32_t bufferIndex = frameIndex % 2;
            fp.numAccumFrames = (uint32_t)frameIndex;
            fp.frameIndex = (uint32_t)frameIndex;
            fp.bufferIndex = bufferIndex;
            fp.resetFlowBuffer = newSequence;
            fp.useUnbiasedEstimator = unbiased;
            check(gfx, gfx_light_dist_build(gfx, nullptr, bufferIndex), "gfx_light_dist_build");

            if (renderer == "restir" || renderer == "restir_unbiased") {                   // restir_di_main.cpp:2352-2421
                uint32_t currentReservoirIndex = (lastReservoirIndex + 1) % 2;
                fp.currentReservoirIndex = currentReservoirIndex;
                fp.spatialNeighborBaseIndex = lastSpatialNeighborBaseIndex;
                check(gfx, gfx_gbuffer_launch(gfx, nullptr, &fp), "gfx_gbuffer_launch");
                const int entry = (fp.enableTemporalReuse && !newSequence)
                    ? (unbiased ? GFX_RESTIR_INITIAL_AND_TEMPORAL_UNBIASED : GFX_RESTIR_INITIAL_AND_TEMPORAL_BIASED)
                    : GFX_RESTIR_INITIAL_RIS;
                check(gfx, gfx_restir_launch(gfx, nullptr, &fp, entry), "performInitialAndTemporalRIS");
                if (fp.enableSpatialReuse) {
                    for (uint32_t i = 0; i < numSpatialReusePasses; ++i) {
                        fp.spatialNeighborBaseIndex = lastSpatialNeighborBaseIndex + fp.numSpatialNeighbors * i;
                        check(gfx, gfx_restir_launch(gfx, nullptr, &fp, unbiased ? GFX_RESTIR_SPATIAL_UNBIASED : GFX_RESTIR_SPATIAL_BIASED),
                              "performSpatialRIS");
                        currentReservoirIndex = (currentReservoirIndex + 1) % 2;
                        fp.currentReservoirIndex = currentReservoirIndex;
                    }
                    lastSpatialNeighborBaseIndex += fp.numSpatialNeighbors * numSpatialReusePasses;
                }
                check(gfx, gfx_restir_launch(gfx, nullptr, &fp, GFX_RESTIR_SHADING), "shading");
                lastReservoirIndex = currentReservoirIndex;
            }
            else if (renderer == "rearch" || renderer == "rearch_unbiased") {              // restir_di_main.cpp:2423-2493
                const uint32_t currentReservoirIndex = (lastReservoirIndex + 1) % 2;
                fp.currentReservoirIndex = currentReservoirIndex;
                fp.spatialNeighborBaseIndex = lastSpatialNeighborBaseIndex;
                check(gfx, gfx_gbuffer_launch(gfx, nullptr, &fp), "gfx_gbuffer_launch");
                check(gfx, gfx_restir_launch(gfx, nullptr, &fp, GFX_RESTIR_PRESAMPLE_LIGHTS), "performLightPreSampling");
                check(gfx, gfx_restir_launch(gfx, nullptr, &fp, GFX_RESTIR_PER_PIXEL_RIS), "performPerPixelRIS");
                GfxFrameParams launch = fp;
                if (newSequence)
                    launch.enableTemporalReuse = launch.enableSpatialReuse = 0;            // plain entry points on a new sequence
                check(gfx, gfx_restir_launch(gfx, nullptr, &launch, GFX_RESTIR_TRACE_SHADOW_RAYS), "traceShadowRays");
                check(gfx, gfx_restir_launch(gfx, nullptr, &launch, GFX_RESTIR_SHADE_AND_RESAMPLE), "shadeAndResample");
                ++lastSpatialNeighborBaseIndex;
                lastReservoirIndex = currentReservoirIndex;
            }
            else if (renderer == "pathtrace") {                                            // path_tracing_main.cpp:1771-1789
                check(gfx, gfx_gbuffer_launch(gfx, nullptr, &fp), "gfx_gbuffer_launch");
                check(gfx, gfx_pathtrace_launch(gfx, nullptr, &fp, GFX_PT_BASELINE), "pathTraceBaseline");
            }
            else if (renderer == "regir") {                                                // regir_main.cpp:2022-2068
                check(gfx, gfx_gbuffer_launch(gfx, nullptr, &fp), "gfx_gbuffer_launch");
                check(gfx, gfx_regir_build_cells(gfx, nullptr, &fp, (uint32_t)frameIndex, !newSequence), "buildCellReservoirs");
                check(gfx, gfx_pathtrace_launch(gfx, nullptr, &fp, GFX_PT_REGIR), "pathTraceReGIR");
                check(gfx, gfx_regir_update_access(gfx, nullptr, &fp, (uint32_t)frameIndex), "updateLastAccessFrameIndices");
            }
            else if (renderer == "nrc") {                                                  // neural_radiance_caching_main.cpp:2256-2368
                check(gfx, gfx_gbuffer_launch(gfx, nullptr, &fp), "gfx_gbuffer_launch");
                const uint32_t offsetToSelectUnbiasedTile = perFrameRng();
                const uint32_t offsetToSelectTrainingPath = perFrameRng();
                check(gfx, gfx_nrc_preprocess(gfx, nullptr, &fp, offsetToSelectUnbiasedTile, offsetToSelectTrainingPath, newSequence),
                      "preprocessNRC");
                check(gfx, gfx_pathtrace_launch(gfx, nullptr, &fp, GFX_PT_NRC), "pathTraceNRC");
                check(gfx, gfx_nrc_frame_infer(gfx, nrc, nullptr), "neuralRadianceCache.infer");
                check(gfx, gfx_nrc_accumulate(gfx, nullptr, &fp), "accumulateInferredRadianceValues");
                check(gfx, gfx_nrc_propagate(gfx, nullptr, &fp), "propagateRadianceValues");
                check(gfx, gfx_nrc_shuffle(gfx, nullptr, &fp), "shuffleTrainingData");
                check(gfx, gfx_nrc_frame_train(gfx, nrc, nullptr, nullptr), "neuralRadianceCache.train");
            }
            else {
                throw std::runtime_error("unknown renderer " + renderer);
            }
            check(gfx, gfx_synchronize(gfx, nullptr), "gfx_synchronize");
            totalMs += std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
        }

        const std::string outPath = argv[6];
        if (outPath.size() > 4 && outPath.compare(outPath.size() - 4, 4, ".png") == 0) {
            // the reference's screenshot: tone map + sRGB gamma on the accumulated beauty (saveImage with
            // SDRImageSaverConfig, common_host.cpp:2859-2897; brightness 0 -> scale 10^0), written as PNG
            GfxPresentParams pp = { GFX_BUF_BEAUTY_ACCUM, 0, GFX_PRESENT_COLOR, GFX_PRESENT_TONE_MAP | GFX_PRESENT_SRGB_GAMMA, 1.0f, 1.0f };
            check(gfx, gfx_present_launch(gfx, nullptr, &pp), "gfx_present_launch");
            std::vector<uint32_t> image((size_t)width * height);
            check(gfx, gfx_buffer_download(gfx, nullptr, GFX_BUF_PRESENT_RGBA8, 0, image.data(), image.size() * 4), "gfx_buffer_download");
            if (!gfxhost::writePng(outPath.c_str(), width, height, image.data()))
                throw std::runtime_error("cannot write the output image");
        }
        else {
            std::vector<float> beauty((size_t)width * height * 4);
            check(gfx, gfx_buffer_download(gfx, nullptr, GFX_BUF_BEAUTY_ACCUM, 0, beauty.data(), beauty.size() * 4), "gfx_buffer_download");
            FILE* of = fopen(outPath.c_str(), "wb");
            if (!of || fwrite(beauty.data(), 4, beauty.size(), of) != beauty.size())
                throw std::runtime_error("cannot write the output image");
            fclose(of);
        }
        printf("{\"renderer\": \"%s\", \"width\": %u, \"height\": %u, \"frames\": %d, \"ms_per_frame\": %.4f, \"kernel_launches\": %llu}\n",
               renderer.c_str(), width, height, frames, totalMs / frames, (unsigned long long)gfx_kernel_launch_count(gfx));
    }
    catch (const std::exception &e) {
        fprintf(stderr, "gfx_headless: %s\n", e.what());
        if (nrc) gfx_nrc_destroy(nrc);
        if (gfx) gfx_ctx_destroy(gfx);
        return 1;
    }
    if (nrc) gfx_nrc_destroy(nrc);
    gfx_ctx_destroy(gfx);
    return 0;
}
