// ComputeShader.h -- C++ host mirror of the reference's cloud-pass wrapper over the C-ABI.
//
//   reference (Shader.h:286-377, Shader.cpp:633-992)          here
//   ComputeShader(device, physicalDevice, commandPool, queue,   ComputeShader(cudaDevice, extent, placement, nightSky,
//                 extent, renderPass, spv, out, outPrev,                      curl, lowRes, hiRes)
//                 placement, nightSky, curl, lowRes, hiRes)
//   updateUniformBuffers(cam, camPrev, sky, sun)                updateUniformBuffers(cam, camPrev, sky, sun)   (same order)
//   bindShader(cmdBuf) + vkCmdDispatch + vkQueueSubmit          dispatch(mode, stream)
//   cleanupUniforms() / ~Shader                                 ~ComputeShader
//   ReprojectShader (Shader.h:380-452)                          bindPrevious + dispatchReproject
//   PostProcessShader x3 (Shader.h:460-520): god-ray.frag,      godRay / radialBlur / tonemapPresent, or postChain (fused)
//       radialBlur.frag, tonemap.frag
//   the shadow march inside model.frag:240-283                  cloudShadow(positions)
// Errors throw std::runtime_error, as every Vulkan failure does in the reference (caught in main.cpp:8-14).
// Header-only; link against libmarshmallow_b200.so.
#pragma once
#include <cstdint>
#include <stdexcept>
#include <string>
#include <vector>

#include "../../include/marshmallow.h"
#include "uniform_blocks.h"

namespace marshmallow {

struct Extent2D { int width, height; };

// The decoded bytes of one texture, as Texture::initFromFile / Texture3D::initFromFile stage them
// (Texture.cpp:212-246, 502-538): RGBA8, x fastest, then y, then z (slice i -> z = i).
struct TextureData {
    const uint8_t *rgba8 = nullptr;
    int width = 0, height = 0, depth = 1;
};

class ComputeShader {
public:
    ComputeShader(int cudaDevice, Extent2D extent, const TextureData &placement, const TextureData *nightSky,
                  const TextureData &curl, const TextureData &lowRes, const TextureData &hiRes)
        : extent_(extent) {
        if (mm_create(cudaDevice, &ctx_) != MM_OK) throw std::runtime_error(std::string("failed to create cloud pass: ") + mm_last_error(nullptr));
        try {
            upload2d(MM_TEX_PLACEMENT, placement);
            if (nightSky && nightSky->rgba8) upload2d(MM_TEX_NIGHTSKY, *nightSky);
            upload2d(MM_TEX_CURL, curl);
            upload3d(MM_TEX_LOWRES, lowRes);
            upload3d(MM_TEX_HIRES, hiRes);
            check(mm_alloc_output(ctx_, extent.width, extent.height, &image_, &pitch_));
        } catch (...) {
            mm_destroy(ctx_);
            throw;
        }
    }
    ComputeShader(const ComputeShader &) = delete;
    ComputeShader &operator=(const ComputeShader &) = delete;
    ~ComputeShader() { if (ctx_) mm_destroy(ctx_); }

    // Shader.cpp:967-992 (the reference memcpy's the four structs into host-visible uniform buffers)
    void updateUniformBuffers(const UniformCameraObject &cam, const UniformCameraObject &camPrev, const UniformSkyObject &sky,
                              const UniformSunObject &sun) {
        check(mm_set_uniforms(ctx_, &cam, &camPrev, &sun, &sky));
    }

    // bind a different output image (e.g. the engine's backgroundTexture exported from Vulkan, or peer memory)
    void bindOutput(float *devicePtr, size_t pitchBytes) {
        check(mm_bind_output_linear(ctx_, devicePtr, pitchBytes, extent_.width, extent_.height));
        image_ = devicePtr; pitch_ = pitchBytes;
    }
    void bindOutputExternalFd(int opaqueFd, size_t allocBytes) { check(mm_bind_output_external_fd(ctx_, opaqueFd, allocBytes, extent_.width, extent_.height)); }

    void setFilterMode(int mode) { check(mm_set_filter_mode(ctx_, mode)); }
    void setLanesPerRay(int lanes) { check(mm_set_lanes_per_ray(ctx_, lanes)); }     // scheduling only; 0 = per dispatch
    void setArithmetic(int arith) { check(mm_set_arithmetic(ctx_, arith)); }          // MM_ARITH_IEEE (default) / MM_ARITH_FMA
    void setScheduler(int scheduler, int refillLanes = 0) { check(mm_set_scheduler(ctx_, scheduler, refillLanes)); }   // scheduling only

    // ReprojectShader: previous image (descriptor set 1) -> bound output, then dispatch(MM_PHASE16) re-marches 1/16 of it
    void bindPrevious(const float *devicePtr, size_t pitchBytes) { check(mm_bind_previous_linear(ctx_, devicePtr, pitchBytes)); }
    void dispatchReproject(void *cudaStream = nullptr) { check(mm_dispatch_reproject(ctx_, cudaStream)); }

    // PostProcessShader x3 (VulkanApplication.cpp:943-968, 1016); device images, source != destination
    void godRay(const UniformCameraObject &cam, const UniformSunObject &sun, const float *src, float *dst, void *cudaStream = nullptr) {
        check(mm_god_ray(ctx_, &cam, &sun, src, pitch_, dst, pitch_, extent_.width, extent_.height, cudaStream));
    }
    void radialBlur(const UniformCameraObject &cam, const UniformSunObject &sun, const float *src, float *dst, void *cudaStream = nullptr) {
        check(mm_radial_blur(ctx_, &cam, &sun, src, pitch_, dst, pitch_, extent_.width, extent_.height, cudaStream));
    }
    void tonemapPresent(const float *src, uint8_t *dst8888, bool bgra = true, void *cudaStream = nullptr) {
        check(mm_tonemap_present(ctx_, src, pitch_, dst8888, (size_t)extent_.width * 4, extent_.width, extent_.height, bgra ? 1 : 0, cudaStream));
    }
    // the whole chain on the bound cloud image: god rays -> radial blur -> tone map + vignette -> swapchain bytes
    void postChain(const UniformCameraObject &cam, const UniformSunObject &sun, uint8_t *dst8888, bool bgra = true, void *cudaStream = nullptr) {
        check(mm_post_chain(ctx_, &cam, &sun, image_, pitch_, dst8888, (size_t)extent_.width * 4, extent_.width, extent_.height, bgra ? 1 : 0, cudaStream));
    }

    // model.frag:240-283 for n world positions (host arrays): accumDensity per point; shade with 1 - 2*density (:281)
    std::vector<float> cloudShadow(const float *positionsXYZ, int n) {
        std::vector<float> out((size_t)n);
        check(mm_cloud_shadow(ctx_, positionsXYZ, n, 0, out.data(), nullptr, nullptr));
        return out;
    }

    // sharded frames: a page-locked host frame every dispatch also stores into (mm_host_register'ed or cudaHostAlloc'ed)
    void bindHostMirror(float *hostFrame) { check(mm_bind_host_mirror(ctx_, hostFrame)); }

    // VulkanApplication.cpp:1062-1071 + 168-177.  mode MM_PHASE16 reproduces one reference dispatch.
    void dispatch(int mode = MM_FULL, void *cudaStream = nullptr, int rowBegin = 0, int rowStride = 1, int rowBlock = 1) {
        check(mm_dispatch(ctx_, mode, rowBegin, rowStride, rowBlock, cudaStream));
    }
    void waitIdle() { check(mm_synchronize(ctx_)); }

    std::vector<float> readImage() {
        std::vector<float> out((size_t)extent_.width * extent_.height * 4);
        check(mm_read_output(ctx_, out.data()));
        return out;
    }
    std::vector<uint8_t> readTonemapped() {
        std::vector<uint8_t> out((size_t)extent_.width * extent_.height * 4);
        check(mm_tonemap_rgba8(ctx_, out.data(), 0, nullptr));
        return out;
    }
    float lastKernelMs() { float ms = 0; check(mm_last_kernel_ms(ctx_, &ms)); return ms; }

    float *image() const { return image_; }
    size_t pitch() const { return pitch_; }
    mm_ctx *handle() const { return ctx_; }

private:
    void check(int rc) { if (rc != MM_OK) throw std::runtime_error(mm_last_error(ctx_)); }
    void upload2d(int slot, const TextureData &t) { check(mm_upload_tex2d(ctx_, slot, t.rgba8, t.width, t.height)); }
    void upload3d(int slot, const TextureData &t) { check(mm_upload_tex3d(ctx_, slot, t.rgba8, t.width, t.height, t.depth)); }
    mm_ctx *ctx_ = nullptr;
    Extent2D extent_;
    float *image_ = nullptr;
    size_t pitch_ = 0;
};

}  // namespace marshmallow
