// ref_wrap_yolov8.cu -- TEST INFRASTRUCTURE.  extern "C" entry points around the REFERENCE's own yolov8
// sources (compiled where they lie under /root/reference by oracle/Makefile target `ref`, against
// tests/mock_trt/NvInfer.h and oracle/shim/opencv2): the YoloLayer plugin (yolov8/plugin/yololayer.cu),
// the host nms() (yolov8/src/postprocess.cpp), cuda_decode/cuda_nms (yolov8/src/postprocess.cu) and
// cuda_preprocess (yolov8/src/preprocess.cu).  Nothing here re-implements the reference.
#include <cuda_runtime_api.h>

#include <vector>

#include "postprocess.h"
#include "preprocess.h"
#include "types.h"
#include "yololayer.h"

extern "C" {
#define REF_API __attribute__((visibility("default")))

REF_API int ref_v8_det_floats() { return (int)(sizeof(Detection) / sizeof(float)); }

// YoloLayerPlugin::enqueue with the reference's own constructor (yololayer.cu:28-45, :167-172)
REF_API int ref_v8_plugin_enqueue(int nc, int nk, float kthr, int netw, int neth, int max_out, int seg, int pose, int obb,
                                  const int* strides, int nstr, int batch, const void* const* inputs_dev,
                                  float* output_dev, void* stream) {
    nvinfer1::YoloLayerPlugin p(nc, nk, kthr, netw, neth, max_out, seg != 0, pose != 0, obb != 0, strides, nstr);
    void* outs[1] = {output_dev};
    int rc = p.enqueue(batch, inputs_dev, outs, nullptr, static_cast<cudaStream_t>(stream));
    cudaError_t e = cudaStreamSynchronize(static_cast<cudaStream_t>(stream));
    return rc != 0 ? rc : (int)e;
}

// host nms() (postprocess.cpp:94-121); res_out must hold max rows * det_floats
REF_API int ref_v8_nms(float* output_host, float conf_thresh, float nms_thresh, float* res_out) {
    std::vector<Detection> res;
    nms(res, output_host, conf_thresh, nms_thresh);
    for (size_t i = 0; i < res.size(); ++i) memcpy(res_out + i * (sizeof(Detection) / 4), &res[i], sizeof(Detection));
    return (int)res.size();
}

// cuda_decode + cuda_nms (postprocess.cu:168-179), batch 1 like the reference driver (yolov8_det.cpp:106-112)
REF_API int ref_v8_cuda_decode_nms(float* predict_dev, int num_bboxes, float conf_thresh, float* parray_dev,
                                   int max_objects, float nms_thresh, void* stream) {
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    cudaMemsetAsync(parray_dev, 0, sizeof(float) * (1 + max_objects * bbox_element), st);
    cuda_decode(predict_dev, num_bboxes, conf_thresh, parray_dev, max_objects, st);
    cuda_nms(parray_dev, nms_thresh, max_objects, st);
    return (int)cudaStreamSynchronize(st);
}

// cuda_preprocess (preprocess.cu:89-117) incl. its pinned staging buffers
REF_API int ref_v8_preprocess(unsigned char* src_host, int w, int h, float* dst_dev, int dw, int dh, void* stream) {
    static bool inited = false;
    if (!inited) {
        cuda_preprocess_init(kMaxInputImageSize);
        inited = true;
    }
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    cuda_preprocess(src_host, w, h, dst_dev, dw, dh, st);
    return (int)cudaStreamSynchronize(st);
}
}
