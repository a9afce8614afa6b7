// ref_wrap_retina_host.cpp -- TEST INFRASTRUCTURE.  CPU-only wrapper around retinaface/common.hpp nms().
#include <cstring>
#include <vector>

#include "common.hpp"

extern "C" {
__attribute__((visibility("default"))) int ref_retina_nms(float* output_host, float nms_thresh, float* res_out) {
    std::vector<decodeplugin::Detection> res;
    nms(res, output_host, nms_thresh);
    for (size_t i = 0; i < res.size(); ++i) memcpy(res_out + i * 15, &res[i], sizeof(decodeplugin::Detection));
    return (int)res.size();
}
// get_rect_adapt_landmark (common.hpp:65-89): lmk is modified in place
__attribute__((visibility("default"))) void ref_retina_get_rect_adapt_landmark(int img_w, int img_h, int input_w, int input_h, float* bbox,
                                                                                float* lmk, int* rect_out) {
    cv::Mat img(img_h, img_w, CV_8UC3, nullptr);
    cv::Rect r = get_rect_adapt_landmark(img, input_w, input_h, bbox, lmk);
    rect_out[0] = r.x, rect_out[1] = r.y, rect_out[2] = r.width, rect_out[3] = r.height;
}
}
