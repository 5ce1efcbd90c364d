// curl_table.h -- host side of K2: the gradient index of every lattice point the curl-noise Perlin can touch.  Host-only (the one step of the
// generator that must run with the C library's sinf, see the header of curl_noise.cu); shared by capi.cu and the CPU suite's host build of K2.
#pragma once
#include <math.h>

namespace mm {

// Gradient index of every lattice point the curl-noise Perlin can touch (coordinates -1..24, stored at
// +1), following hashNoise / hashVec (ImageUtils.cpp:25-34) with the host C library's sinf -- see the
// header of curl_noise.cu for why this one step stays on the host.
inline void build_curl_gradient_table(unsigned char *table) {
    const float kx = 12.9898f, ky = 78.233f, kz = (float)47.387;
    for (int z = -1; z < 25; z++)
        for (int y = -1; y < 25; y++)
            for (int x = -1; x < 25; x++) {
                float d = (((float)x * kx) + ((float)y * ky)) + ((float)z * kz);
                float n = sinf(d) * 43758.5453f;
                n = n - floorf(n);
                table[((z + 1) * 26 + (y + 1)) * 26 + (x + 1)] = (unsigned char)(int)floorf(12.f * n);
            }
}

}  // namespace mm
