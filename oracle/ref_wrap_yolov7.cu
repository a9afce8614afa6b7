// ref_wrap_yolov7.cu -- TEST INFRASTRUCTURE.  extern "C" entry point around the REFERENCE's yolov7 YoloLayer plugin
// (yolov7/plugin/yololayer.cu: 6-float Detection rows, yolov7/include/types.h:11-16), compiled from /root/reference.
#include <cuda_runtime_api.h>

#include <cstring>
#include <vector>

#include "types.h"
#include "yololayer.h"

extern "C" {
#define REF_API __attribute__((visibility("default")))
REF_API int ref_v7_det_floats() { return (int)(sizeof(Detection) / sizeof(float)); }

// kernels: nlevels x {int w, int h, float anchors[6]} = YoloKernel (types.h:5-9)
REF_API int ref_v7_plugin_enqueue(int nc, int netw, int neth, int max_out, const void* kernels, int nlevels, int batch,
                                  const void* const* inputs_dev, float* output_dev, void* stream) {
    std::vector<YoloKernel> ks(nlevels);
    memcpy(ks.data(), kernels, sizeof(YoloKernel) * nlevels);
    nvinfer1::YoloLayerPlugin p(nc, netw, neth, max_out, ks);
    void* outs[1] = {output_dev};
    int rc = p.enqueue(batch, inputs_dev, outs, nullptr, static_cast<cudaStream_t>(stream));
    cudaError_t e = cudaStreamSynchronize(static_cast<cudaStream_t>(stream));
    return rc != 0 ? rc : (int)e;
}
}
