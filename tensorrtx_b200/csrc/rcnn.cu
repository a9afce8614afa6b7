// rcnn.cu -- placeholder until the batched RPN / predictor kernels land (next commit).
#include "common.cuh"
using namespace trtx;
extern "C" {
TRTX_API int64_t trtx_rpn_decode(int, const float*, const float*, float*, float*, int, int, int, int, float,
                                 const float*, int, int, void*, size_t, trtx_stream_t) { return -TRTX_ERR_UNSUPPORTED; }
TRTX_API int64_t trtx_rpn_nms(int, const float*, const float*, float*, int, int, float, void*, size_t, trtx_stream_t) {
    return -TRTX_ERR_UNSUPPORTED;
}
TRTX_API int64_t trtx_predictor_decode(int, const float*, const float*, const float*, float*, float*, float*, int, int,
                                       int, int, const float*, void*, size_t, trtx_stream_t) {
    return -TRTX_ERR_UNSUPPORTED;
}
TRTX_API int64_t trtx_batched_nms(int, int, const float*, const float*, const float*, float*, float*, float*, int, int,
                                  float, void*, size_t, trtx_stream_t) {
    return -TRTX_ERR_UNSUPPORTED;
}
}
