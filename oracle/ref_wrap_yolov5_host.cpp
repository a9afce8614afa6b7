// ref_wrap_yolov5_host.cpp -- TEST INFRASTRUCTURE.  CPU-only wrapper around yolov5/src/postprocess.cpp nms().
#include <cstring>
#include <vector>

#include "postprocess.h"
#include "types.h"

extern "C" {
__attribute__((visibility("default"))) int ref_v5_det_floats() { return (int)(sizeof(Detection) / sizeof(float)); }
__attribute__((visibility("default"))) int ref_v5_nms(float* output_host, float conf_thresh, float nms_thresh, float* res_out) {
    std::vector<Detection> res;
    nms(res, output_host, conf_thresh, nms_thresh);
    for (size_t i = 0; i < res.size(); ++i) memcpy(res_out + i * (sizeof(Detection) / 4), &res[i], sizeof(Detection));
    return (int)res.size();
}
}
