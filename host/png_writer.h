// png_writer.h — dependency-free PNG writer for the headless host: 8-bit RGBA, filter 0, stored (uncompressed) deflate
// blocks.  The reference writes its screenshots with stb_image_write (saveImage, common/common_host.cpp:2715-2723) from
// the packed image R | G << 8 | B << 16 | A << 24 that gfx_present_launch produces; this stands in for it where stb is
// not vendored.  tests/test_host_png.py builds a small CPU program around it and decodes the file with zlib.
#pragma once
#include <cstdint>
#include <cstdio>
#include <vector>

namespace gfxhost {

inline uint32_t crc32(const uint8_t* data, size_t n, uint32_t crc = 0) {
    static uint32_t table[256];
    static bool init = false;
    if (!init) {
        for (uint32_t i = 0; i < 256; ++i) {
            uint32_t c = i;
            for (int k = 0; k < 8; ++k)
                c = (c & 1u) ? 0xEDB88320u ^ (c >> 1) : c >> 1;
            table[i] = c;
        }
        init = true;
    }
    crc = ~crc;
    for (size_t i = 0; i < n; ++i)
        crc = table[(crc ^ data[i]) & 0xFFu] ^ (crc >> 8);
    return ~crc;
}

inline void put32(std::vector<uint8_t> &v, uint32_t x) {
    v.push_back((uint8_t)(x >> 24)); v.push_back((uint8_t)(x >> 16)); v.push_back((uint8_t)(x >> 8)); v.push_back((uint8_t)x);
}

inline void chunk(std::vector<uint8_t> &png, const char tag[4], const std::vector<uint8_t> &body) {
    put32(png, (uint32_t)body.size());
    const size_t at = png.size();
    png.insert(png.end(), tag, tag + 4);
    png.insert(png.end(), body.begin(), body.end());
    put32(png, crc32(png.data() + at, png.size() - at));
}

// rgba8: width * height packed pixels, R in the low byte (little-endian memory order R, G, B, A)
inline bool writePng(const char* path, uint32_t width, uint32_t height, const uint32_t* rgba8) {
    std::vector<uint8_t> raw;
    raw.reserve((size_t)height * (1 + 4 * (size_t)width));
    for (uint32_t y = 0; y < height; ++y) {
        raw.push_back(0); // filter type 0
        for (uint32_t x = 0; x < width; ++x) {
            const uint32_t p = rgba8[(size_t)y * width + x];
            raw.push_back((uint8_t)p); raw.push_back((uint8_t)(p >> 8)); raw.push_back((uint8_t)(p >> 16)); raw.push_back((uint8_t)(p >> 24));
        }
    }
    // zlib stream of stored blocks
    std::vector<uint8_t> z;
    z.push_back(0x78); z.push_back(0x01);
    uint32_t a = 1, b = 0; // Adler-32
    size_t pos = 0;
    do {
        const size_t n = raw.size() - pos < 65535 ? raw.size() - pos : 65535;
        const bool last = pos + n == raw.size();
        z.push_back(last ? 1 : 0);
        z.push_back((uint8_t)n); z.push_back((uint8_t)(n >> 8));
        z.push_back((uint8_t)~n); z.push_back((uint8_t)(~n >> 8));
        for (size_t i = 0; i < n; ++i) {
            const uint8_t c = raw[pos + i];
            z.push_back(c);
            a = (a + c) % 65521u;
            b = (b + a) % 65521u;
        }
        pos += n;
    } while (pos < raw.size());
    put32(z, (b << 16) | a);

    std::vector<uint8_t> png = { 0x89, 'P', 'N', 'G', '\r', '\n', 0x1A, '\n' };
    std::vector<uint8_t> ihdr;
    put32(ihdr, width); put32(ihdr, height);
    ihdr.push_back(8); ihdr.push_back(6); ihdr.push_back(0); ihdr.push_back(0); ihdr.push_back(0);
    chunk(png, "IHDR", ihdr);
    chunk(png, "IDAT", z);
    chunk(png, "IEND", {});
    FILE* f = fopen(path, "wb");
    if (!f)
        return false;
    const bool ok = fwrite(png.data(), 1, png.size(), f) == png.size();
    fclose(f);
    return ok;
}

} // namespace gfxhost
