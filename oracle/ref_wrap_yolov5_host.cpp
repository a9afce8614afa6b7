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
// get_rect (postprocess.cpp:4-36): box in network-input pixels -> cv::Rect in the original image (kInputW x kInputH = 640 x 640)
__attribute__((visibility("default"))) void ref_v5_get_rect(int img_w, int img_h, float* bbox, int* rect_out) {
    cv::Mat img(img_h, img_w, CV_8UC3, nullptr);
    cv::Rect r = get_rect(img, bbox);
    rect_out[0] = r.x;
    rect_out[1] = r.y;
    rect_out[2] = r.width;
    rect_out[3] = r.height;
}
}
