// ref_wrap_yolov3.cu -- TEST INFRASTRUCTURE.  extern "C" entry point around the REFERENCE's yolov3-spp YoloLayer plugin
// (yolov3-spp/yololayer.cu: IPluginV2DynamicExt, 7-float Detection rows with separate det/class confidence,
// yololayer.h:47-53; compiled-in CLASS_NUM = 80, anchors and level order stride 32, 16, 8), compiled from /root/reference.
// The plugin launches on the DEFAULT stream (yololayer.cu:203) and memsets synchronously: the wrapper synchronises the device.
#include <cuda_runtime_api.h>

#include "yololayer.h"

extern "C" {
#define REF_API __attribute__((visibility("default")))
REF_API int ref_v3_det_floats() { return (int)(sizeof(Yolo::Detection) / sizeof(float)); }
REF_API int ref_v3_num_classes() { return Yolo::CLASS_NUM; }

// inputs_dev[i]: level i tensor [batch, 3*(5+80), gh[i], gw[i]], levels in the plugin's order (stride 32, 16, 8)
REF_API int ref_v3_plugin_enqueue(int batch, const int* gh, const int* gw, const void* const* inputs_dev, float* output_dev) {
    nvinfer1::YoloLayerPlugin p;
    nvinfer1::PluginTensorDesc in[3], out[1];
    for (int i = 0; i < 3; ++i) {
        in[i].dims.nbDims = 4;
        in[i].dims.d[0] = batch;
        in[i].dims.d[1] = 3 * (5 + Yolo::CLASS_NUM);
        in[i].dims.d[2] = gh[i];
        in[i].dims.d[3] = gw[i];
        in[i].type = nvinfer1::DataType::kFLOAT;
        in[i].format = nvinfer1::TensorFormat::kLINEAR;
    }
    out[0] = in[0];
    void* outs[1] = {output_dev};
    int rc = p.enqueue(in, out, inputs_dev, outs, nullptr, nullptr);
    cudaError_t e = cudaDeviceSynchronize();
    return rc != 0 ? rc : (int)e;
}
}
