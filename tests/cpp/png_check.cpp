// TEST INFRASTRUCTURE (CPU): writes a deterministic RGBA8 test pattern with host/png_writer.h; tests/test_host_png.py
// decodes the file with zlib and compares.
#include "../../host/png_writer.h"
#include <cstdlib>
int main(int argc, char** argv) {
    if (argc < 4) return 2;
    const uint32_t w = (uint32_t)atoi(argv[2]), h = (uint32_t)atoi(argv[3]);
    std::vector<uint32_t> img((size_t)w * h);
    uint32_t s = 12345u;
    for (auto &p : img) { s = s * 1664525u + 1013904223u; p = s; }
    return gfxhost::writePng(argv[1], w, h, img.data()) ? 0 : 1;
}
