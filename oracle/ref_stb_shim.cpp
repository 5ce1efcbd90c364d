// ref_stb_shim.cpp -- the reference's own image decoder (Libraries/stb/stb_image.h, vendored in /root/reference) called the way
// Texture::initFromFile / Texture3D::initFromFile call it (Texture.cpp:212-246, 502-538: stbi_load(path, &w, &h, &channels,
// STBI_rgb_alpha)).  TEST INFRASTRUCTURE (oracle/_ref/libref_stb.so): pins the committed asset fixtures to the decoder the engine uses.
#define STB_IMAGE_IMPLEMENTATION
#include "stb_image.h"
#include <cstring>

extern "C" int ref_stbi_load_rgba(const char *path, int *w, int *h, unsigned char *out, size_t out_bytes) {
    int channels = 0;
    stbi_uc *pixels = stbi_load(path, w, h, &channels, STBI_rgb_alpha);
    if (!pixels) return -1;
    size_t n = (size_t)(*w) * (size_t)(*h) * 4;
    int rc = 0;
    if (out && n <= out_bytes) memcpy(out, pixels, n); else rc = -2;
    stbi_image_free(pixels);
    return rc;
}
