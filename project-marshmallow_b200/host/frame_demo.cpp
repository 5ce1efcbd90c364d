// frame_demo.cpp -- the reference's per-frame sequence (VulkanApplication.cpp:351-386 updateUniformBuffer, :164-177
// drawFrame's compute submit) for the cloud pass alone, headless: build the uniform blocks with SkyManager / Camera,
// update, dispatch, read back.  Textures come from raw RGBA8 files (w*h*d*4 bytes each).
//   frame_demo <placement512.raw> <curl128.raw> <low128.raw> <hi32.raw> <W> <H> <out.f32> [elevation] [filter]
#include <cstdio>
#include <cstdlib>
#include <fstream>
#include <iostream>
#include <iterator>

#include "Camera.h"
#include "ComputeShader.h"
#include "SkyManager.h"

using namespace marshmallow;

static std::vector<uint8_t> readFile(const char *path) {
    std::ifstream f(path, std::ios::binary);
    if (!f) throw std::runtime_error(std::string("failed to open file ") + path);
    return std::vector<uint8_t>(std::istreambuf_iterator<char>(f), {});
}

int main(int argc, char **argv) {
    if (argc < 8) { std::cerr << "usage: frame_demo placement.raw curl.raw low.raw hi.raw W H out.f32 [elevation] [filter]\n"; return 2; }
    try {
        auto placement = readFile(argv[1]), curl = readFile(argv[2]), low = readFile(argv[3]), hi = readFile(argv[4]);
        int W = std::atoi(argv[5]), H = std::atoi(argv[6]);
        float elevation = argc > 8 ? (float)std::atof(argv[8]) : 0.25f;
        int filter = argc > 9 ? std::atoi(argv[9]) : MM_FILTER_HW;
        TextureData tp{placement.data(), 512, 512, 1}, tc{curl.data(), 128, 128, 1}, tl{low.data(), 128, 128, 128}, th{hi.data(), 32, 32, 32};
        ComputeShader computeShader(0, Extent2D{W, H}, tp, nullptr, tc, tl, th);
        computeShader.setFilterMode(filter);

        const float pos[3] = {0.0f, 1.0f, 1.0f};
        Camera mainCamera(pos, (float)(-3.141592653589793 / 2.0), (float)(-20 * 0.01745));   // config C1 (tests/scenes.py)
        SkyManager skySystem;
        UniformCameraObject uco, ucoPrev;
        mainCamera.fillUniform(uco);
        ucoPrev = uco;
        skySystem.rebuildSkyFromNewSun(elevation, 0.25f);
        skySystem.setTime(0.0f);
        UniformSkyObject sky = skySystem.getSky();
        UniformSunObject &sun = skySystem.getSun();
        computeShader.updateUniformBuffers(uco, ucoPrev, sky, sun);
        computeShader.dispatch(MM_FULL);
        computeShader.waitIdle();
        std::vector<float> img = computeShader.readImage();
        std::ofstream out(argv[7], std::ios::binary);
        out.write(reinterpret_cast<const char *>(img.data()), (std::streamsize)(img.size() * sizeof(float)));
        std::cout << "frame " << W << "x" << H << " kernel " << computeShader.lastKernelMs() << " ms\n";
    } catch (const std::runtime_error &e) {
        std::cerr << e.what() << std::endl;      // main.cpp:8-14
        return EXIT_FAILURE;
    }
    return EXIT_SUCCESS;
}
