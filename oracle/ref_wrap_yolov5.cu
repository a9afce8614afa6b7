// ref_wrap_yolov5.cu -- TEST INFRASTRUCTURE.  extern "C" entry points around the REFERENCE's yolov5 plugin
// (yolov5/plugin/yololayer.cu) and host nms() (yolov5/src/postprocess.cpp), compiled from /root/reference.
#include <cuda_runtime_api.h>

#include <cstring>
#include <vector>

#include "postprocess.h"
#include "types.h"
#include "yololayer.h"

extern "C" {
#define REF_API __attribute__((visibility("default")))
REF_API int ref_v5_det_floats() { return (int)(sizeof(Detection) / sizeof(float)); }

// kernels: nlevels x {int w, int h, float anchors[6]} = YoloKernel (types.h:5-9)
REF_API int ref_v5_plugin_enqueue(int nc, int netw, int neth, int max_out, int seg, const void* kernels, int nlevels,
                                  int batch, const void* const* inputs_dev, float* output_dev, void* stream) {
    std::vector<YoloKernel> ks(nlevels);
    memcpy(ks.data(), kernels, sizeof(YoloKernel) * nlevels);
    nvinfer1::YoloLayerPlugin p(nc, netw, neth, max_out, seg != 0, ks);
    void* outs[1] = {output_dev};
    int rc = p.enqueue(batch, inputs_dev, outs, nullptr, static_cast<cudaStream_t>(stream));
    cudaError_t e = cudaStreamSynchronize(static_cast<cudaStream_t>(stream));
    return rc != 0 ? rc : (int)e;
}

REF_API int ref_v5_nms(float* output_host, float conf_thresh, float nms_thresh, float* res_out) {
    std::vector<Detection> res;
    nms(res, output_host, conf_thresh, nms_thresh);
    for (size_t i = 0; i < res.size(); ++i) memcpy(res_out + i * (sizeof(Detection) / 4), &res[i], sizeof(Detection));
    return (int)res.size();
}
}
