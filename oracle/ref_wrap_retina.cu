// ref_wrap_retina.cu -- TEST INFRASTRUCTURE.  extern "C" entry points around the REFERENCE's RetinaFace
// Decode_TRT plugin (retinaface/decode.cu; INPUT_H x INPUT_W = 480 x 640 is compiled in, decode.h:16-17)
// and host nms() (retinaface/common.hpp:110-130), compiled from /root/reference.
#include <cuda_runtime_api.h>

#include <cstring>
#include <vector>

#include "common.hpp"
#include "decode.h"

extern "C" {
#define REF_API __attribute__((visibility("default")))
REF_API int ref_retina_input_h() { return decodeplugin::INPUT_H; }
REF_API int ref_retina_input_w() { return decodeplugin::INPUT_W; }

REF_API int ref_retina_plugin_enqueue(int batch, const void* const* inputs_dev, float* output_dev, void* stream) {
    nvinfer1::DecodePlugin p;
    void* outs[1] = {output_dev};
    int rc = p.enqueue(batch, inputs_dev, outs, nullptr, static_cast<cudaStream_t>(stream));
    cudaError_t e = cudaStreamSynchronize(static_cast<cudaStream_t>(stream));
    return rc != 0 ? rc : (int)e;
}

REF_API int ref_retina_nms(float* output_host, float nms_thresh, float* res_out) {
    std::vector<decodeplugin::Detection> res;
    nms(res, output_host, nms_thresh);
    for (size_t i = 0; i < res.size(); ++i) memcpy(res_out + i * 15, &res[i], sizeof(decodeplugin::Detection));
    return (int)res.size();
}
}
