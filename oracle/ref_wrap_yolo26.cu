// ref_wrap_yolo26.cu -- TEST INFRASTRUCTURE.  extern "C" entry point around the REFERENCE's yolo26 YoloLayer plugin
// (yolo26/plugin/yololayer.cu: NMS-free gatherKernel over AoS rows [anchor_count, 4 + nc (+1)], batch 1 only),
// compiled from /root/reference.
#include <cuda_runtime_api.h>

#include "types.h"
#include "yololayer.h"

extern "C" {
#define REF_API __attribute__((visibility("default")))
REF_API int ref_v26_det_floats() { return (int)(sizeof(Detection) / sizeof(float)); }

REF_API int ref_v26_plugin_enqueue(int nc, int nk, int max_det, int is_det, int is_obb, int anchor_count, float conf_thresh,
                                   const void* input_dev, float* output_dev, void* stream) {
    nvinfer1::setPluginDeviceParams(conf_thresh);
    nvinfer1::YoloLayerPlugin p(nc, nk, max_det, is_det != 0, false, false, is_obb != 0, anchor_count);
    const void* ins[1] = {input_dev};
    void* outs[1] = {output_dev};
    int rc = p.enqueue(1, ins, outs, nullptr, static_cast<cudaStream_t>(stream));
    cudaError_t e = cudaStreamSynchronize(static_cast<cudaStream_t>(stream));
    return rc != 0 ? rc : (int)e;
}
}
