/* trtx_preprocess_compat.h -- source-level drop-in for the reference's include/preprocess.h.
 *
 * Declares (and, in the ONE translation unit that defines TRTX_PREPROCESS_COMPAT_IMPL, implements) the four functions of
 * yolov8/include/preprocess.h:8-16 (identical in yolov5/7/9/10/11/12/13) on top of the C ABI of trtx_hot.h:
 *     void cuda_preprocess_init(int max_image_size);
 *     void cuda_preprocess_destroy();
 *     void cuda_preprocess(uint8_t* src, int src_width, int src_height, float* dst, int dst_width, int dst_height, cudaStream_t);
 *     void cuda_batch_preprocess(std::vector<cv::Mat>& img_batch, float* dst, int dst_width, int dst_height, cudaStream_t);
 * so that a driver such as yolov8_det.cpp builds unchanged with `#include "trtx_preprocess_compat.h"` instead of
 * "preprocess.h" and without src/preprocess.cu.
 *
 * What changes underneath (yolov8/src/preprocess.cu:89-127): the reference copies ONE image to a pinned buffer, issues one
 * H2D copy and one kernel, and calls cudaStreamSynchronize() after EVERY image of the batch.  Here the whole batch is
 * packed into a pinned ring slot, goes to the device with ONE cudaMemcpyAsync and is processed by ONE launch; nothing
 * synchronises the stream (work that follows on the same stream -- the engine's enqueue -- is ordered by the stream).
 * The host only waits when it is about to overwrite a ring slot whose previous H2D copy has not finished.
 *
 * Requires OpenCV's cv::Mat (rows, cols, ptr(), isContinuous-style packed BGR u8 data, like the reference assumes).
 */
#ifndef TRTX_PREPROCESS_COMPAT_H
#define TRTX_PREPROCESS_COMPAT_H

#include <cuda_runtime_api.h>

#include <cstdint>
#include <opencv2/opencv.hpp>
#include <vector>

#include "trtx_hot.h"

void cuda_preprocess_init(int max_image_size);
void cuda_preprocess_destroy();
void cuda_preprocess(uint8_t* src, int src_width, int src_height, float* dst, int dst_width, int dst_height,
                     cudaStream_t stream);
void cuda_batch_preprocess(std::vector<cv::Mat>& img_batch, float* dst, int dst_width, int dst_height, cudaStream_t stream);

#ifdef TRTX_PREPROCESS_COMPAT_IMPL
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>

namespace trtx_compat {

constexpr int kRingSlots = 2;  // batch i+1 is packed on the host while batch i is still being copied

struct Slot {
    uint8_t* host = nullptr;  // pinned
    uint8_t* dev = nullptr;
    size_t bytes = 0;
    cudaEvent_t copied = nullptr;  // the slot's last H2D copy
    bool used = false;
};

struct State {
    std::mutex mu;
    size_t max_image_bytes = 0;
    Slot slot[kRingSlots];
    int next = 0;
};

inline State& state() {
    static State s;
    return s;
}

inline void die(const char* what, cudaError_t e) {
    std::fprintf(stderr, "trtx_preprocess_compat: %s: %s\n", what, cudaGetErrorString(e));
    std::abort();  // the reference's CUDA_CHECK asserts
}

inline size_t align_up(size_t v) { return (v + 255) & ~(size_t)255; }

// a slot that holds at least `bytes`; waits for the slot's previous copy before its pinned memory is reused
inline Slot& acquire(State& s, size_t bytes) {
    Slot& sl = s.slot[s.next];
    s.next = (s.next + 1) % kRingSlots;
    if (sl.used) {
        cudaError_t e = cudaEventSynchronize(sl.copied);
        if (e != cudaSuccess) die("cudaEventSynchronize", e);
    }
    if (sl.bytes < bytes) {
        if (sl.host) cudaFreeHost(sl.host);
        if (sl.dev) cudaFree(sl.dev);
        cudaError_t e = cudaMallocHost(reinterpret_cast<void**>(&sl.host), bytes);
        if (e != cudaSuccess) die("cudaMallocHost", e);
        e = cudaMalloc(reinterpret_cast<void**>(&sl.dev), bytes);
        if (e != cudaSuccess) die("cudaMalloc", e);
        sl.bytes = bytes;
    }
    if (!sl.copied) {
        cudaError_t e = cudaEventCreateWithFlags(&sl.copied, cudaEventDisableTiming);
        if (e != cudaSuccess) die("cudaEventCreate", e);
    }
    return sl;
}

inline void run(const uint8_t* const* src, const int* w, const int* h, int n, float* dst, int dst_width, int dst_height,
                cudaStream_t stream) {
    if (n <= 0) return;
    State& s = state();
    std::lock_guard<std::mutex> lock(s.mu);
    std::vector<size_t> off((size_t)n + 1, 0);
    for (int i = 0; i < n; ++i) off[i + 1] = off[i] + align_up((size_t)w[i] * h[i] * 3);
    Slot& sl = acquire(s, off[n]);
    std::vector<trtx_image_desc> d((size_t)n);
    for (int i = 0; i < n; ++i) {
        std::memcpy(sl.host + off[i], src[i], (size_t)w[i] * h[i] * 3);  // packed BGR rows, like preprocess.cu:93
        d[i].data_dev = sl.dev + off[i];
        d[i].width = w[i];
        d[i].height = h[i];
        d[i].pitch = w[i] * 3;
        d[i].reserved = 0;
    }
    cudaError_t e = cudaMemcpyAsync(sl.dev, sl.host, off[n], cudaMemcpyHostToDevice, stream);
    if (e != cudaSuccess) die("cudaMemcpyAsync", e);
    e = cudaEventRecord(sl.copied, stream);
    if (e != cudaSuccess) die("cudaEventRecord", e);
    sl.used = true;
    const int rc = trtx_preprocess_batch_enqueue(d.data(), n, dst, dst_width, dst_height, TRTX_F32, stream);
    if (rc != TRTX_OK) {
        std::fprintf(stderr, "trtx_preprocess_compat: trtx_preprocess_batch_enqueue failed (%d)\n", rc);
        std::abort();
    }
}

}  // namespace trtx_compat

void cuda_preprocess_init(int max_image_size) {
    trtx_compat::State& s = trtx_compat::state();
    std::lock_guard<std::mutex> lock(s.mu);
    s.max_image_bytes = (size_t)max_image_size * 3;  // preprocess.cu:121-124; slots grow on demand to the batch actually used
}

void cuda_preprocess_destroy() {
    trtx_compat::State& s = trtx_compat::state();
    std::lock_guard<std::mutex> lock(s.mu);
    for (trtx_compat::Slot& sl : s.slot) {
        if (sl.used) cudaEventSynchronize(sl.copied);
        if (sl.host) cudaFreeHost(sl.host);
        if (sl.dev) cudaFree(sl.dev);
        if (sl.copied) cudaEventDestroy(sl.copied);
        sl = trtx_compat::Slot();
    }
}

void cuda_preprocess(uint8_t* src, int src_width, int src_height, float* dst, int dst_width, int dst_height,
                     cudaStream_t stream) {
    const uint8_t* p = src;
    trtx_compat::run(&p, &src_width, &src_height, 1, dst, dst_width, dst_height, stream);
}

void cuda_batch_preprocess(std::vector<cv::Mat>& img_batch, float* dst, int dst_width, int dst_height, cudaStream_t stream) {
    const int n = (int)img_batch.size();
    std::vector<const uint8_t*> src((size_t)n);
    std::vector<int> w((size_t)n), h((size_t)n);
    for (int i = 0; i < n; ++i) {
        src[i] = img_batch[i].ptr();
        w[i] = img_batch[i].cols;
        h[i] = img_batch[i].rows;
    }
    trtx_compat::run(src.data(), w.data(), h.data(), n, dst, dst_width, dst_height, stream);
}
#endif /* TRTX_PREPROCESS_COMPAT_IMPL */

#endif /* TRTX_PREPROCESS_COMPAT_H */
