// Compiles include/trtx_preprocess_compat.h (the source-level drop-in for the reference's preprocess.h) against the
// OpenCV type shim and links it to libtrtx_hot.so.  Without arguments only the host surface is exercised; with --gpu a
// batch of frames goes through cuda_batch_preprocess and is compared with a direct call of the C ABI.
#define TRTX_PREPROCESS_COMPAT_IMPL
#include "trtx_preprocess_compat.h"

#include <cstdio>
#include <cstring>
#include <string>

#define CHECK(c)                                                           \
    do {                                                                   \
        if (!(c)) {                                                        \
            std::printf("CHECK failed %s:%d: %s\n", __FILE__, __LINE__, #c); \
            return 1;                                                      \
        }                                                                  \
    } while (0)

int main(int argc, char** argv) {
    const bool gpu = argc > 1 && std::string(argv[1]) == "--gpu";
    // the four entry points of preprocess.h exist with the reference's signatures
    void (*f0)(int) = &cuda_preprocess_init;
    void (*f1)() = &cuda_preprocess_destroy;
    void (*f2)(uint8_t*, int, int, float*, int, int, cudaStream_t) = &cuda_preprocess;
    void (*f3)(std::vector<cv::Mat>&, float*, int, int, cudaStream_t) = &cuda_batch_preprocess;
    CHECK(f0 && f1 && f2 && f3);
    if (gpu) {
        const int B = 3, W = 320, H = 200, D = 256;
        std::vector<std::vector<uint8_t>> frames(B, std::vector<uint8_t>((size_t)W * H * 3));
        std::vector<cv::Mat> batch;
        for (int b = 0; b < B; ++b) {
            for (size_t i = 0; i < frames[b].size(); ++i) frames[b][i] = (uint8_t)((i * 7 + b * 31) & 255);
            batch.emplace_back(H, W, CV_8UC3, frames[b].data());
        }
        float *dst = nullptr, *ref = nullptr;
        uint8_t* dev = nullptr;
        CHECK(cudaMalloc((void**)&dst, sizeof(float) * B * 3 * D * D) == cudaSuccess);
        CHECK(cudaMalloc((void**)&ref, sizeof(float) * B * 3 * D * D) == cudaSuccess);
        CHECK(cudaMalloc((void**)&dev, (size_t)B * W * H * 3) == cudaSuccess);
        cuda_preprocess_init(W * H);
        for (int rep = 0; rep < 3; ++rep) cuda_batch_preprocess(batch, dst, D, D, nullptr);  // ring slots wrap around
        std::vector<trtx_image_desc> d(B);
        for (int b = 0; b < B; ++b) {
            CHECK(cudaMemcpy(dev + (size_t)b * W * H * 3, frames[b].data(), (size_t)W * H * 3, cudaMemcpyHostToDevice) == cudaSuccess);
            d[b] = {dev + (size_t)b * W * H * 3, W, H, W * 3, 0};
        }
        CHECK(trtx_preprocess_batch_enqueue(d.data(), B, ref, D, D, TRTX_F32, nullptr) == TRTX_OK);
        CHECK(cudaDeviceSynchronize() == cudaSuccess);
        std::vector<float> a((size_t)B * 3 * D * D), c(a.size());
        CHECK(cudaMemcpy(a.data(), dst, a.size() * 4, cudaMemcpyDeviceToHost) == cudaSuccess);
        CHECK(cudaMemcpy(c.data(), ref, c.size() * 4, cudaMemcpyDeviceToHost) == cudaSuccess);
        CHECK(std::memcmp(a.data(), c.data(), a.size() * 4) == 0);
        cuda_preprocess(frames[0].data(), W, H, dst, D, D, nullptr);
        CHECK(cudaDeviceSynchronize() == cudaSuccess);
        CHECK(cudaMemcpy(a.data(), dst, (size_t)3 * D * D * 4, cudaMemcpyDeviceToHost) == cudaSuccess);
        CHECK(std::memcmp(a.data(), c.data(), (size_t)3 * D * D * 4) == 0);
        cuda_preprocess_destroy();
        std::printf("gpu preprocess compat ok\n");
    }
    std::printf("preprocess compat check ok\n");
    return 0;
}
