// ref_wrap_anticov.cu -- TEST INFRASTRUCTURE.  extern "C" entry point around the REFERENCE's retinafaceAntiCov Decode_TRT
// plugin (retinafaceAntiCov/decode.cu: batch 1, 640x640 compiled in, 16-float rows with a mask confidence), compiled
// from /root/reference.  The plugin launches on the DEFAULT stream (decode.cu:156-166): the wrapper synchronises the device.
#include <cuda_runtime_api.h>

#include "decode.h"

extern "C" {
#define REF_API __attribute__((visibility("default")))
REF_API int ref_anticov_input_h() { return decodeplugin::INPUT_H; }
REF_API int ref_anticov_input_w() { return decodeplugin::INPUT_W; }
REF_API int ref_anticov_det_floats() { return (int)(sizeof(decodeplugin::Detection) / sizeof(float)); }

REF_API int ref_anticov_plugin_enqueue(const void* const* inputs_dev, float* output_dev) {
    nvinfer1::DecodePlugin p;
    void* outs[1] = {output_dev};
    int rc = p.enqueue(1, inputs_dev, outs, nullptr, nullptr);
    cudaError_t e = cudaDeviceSynchronize();
    return rc != 0 ? rc : (int)e;
}
}
