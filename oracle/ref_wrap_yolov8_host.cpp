// ref_wrap_yolov8_host.cpp -- TEST INFRASTRUCTURE.  CPU-only wrapper around the reference's host nms()
// (yolov8/src/postprocess.cpp:94-129) so that the oracle can be pinned WITHOUT a GPU, and so that
// bench.py can time the reference's own CPU code (cpu_baseline.kind = "reference" for the NMS part).
#include <cstring>
#include <vector>

#include "postprocess.h"
#include "types.h"

// defined in yolov8/src/postprocess.cpp (:38, :207) without a declaration in postprocess.h
cv::Rect get_rect_adapt_landmark(cv::Mat& img, float bbox[4], float lmk[kNumberOfPoints * 3]);
cv::Mat scale_mask(cv::Mat mask, cv::Mat img);

extern "C" {
__attribute__((visibility("default"))) int ref_v8_det_floats() { return (int)(sizeof(Detection) / sizeof(float)); }
__attribute__((visibility("default"))) int ref_v8_nms(float* output_host, float conf_thresh, float nms_thresh, float* res_out) {
    std::vector<Detection> res;
    nms(res, output_host, conf_thresh, nms_thresh);
    for (size_t i = 0; i < res.size(); ++i) memcpy(res_out + i * (sizeof(Detection) / 4), &res[i], sizeof(Detection));
    return (int)res.size();
}
// nms_obb (postprocess.cpp:357-385): greedy, probiou >= thresh
__attribute__((visibility("default"))) int ref_v8_nms_obb(float* output_host, float conf_thresh, float nms_thresh, float* res_out) {
    std::vector<Detection> res;
    nms_obb(res, output_host, conf_thresh, nms_thresh);
    for (size_t i = 0; i < res.size(); ++i) memcpy(res_out + i * (sizeof(Detection) / 4), &res[i], sizeof(Detection));
    return (int)res.size();
}
__attribute__((visibility("default"))) int ref_v8_batch_nms(float* output_host, int batch, int output_size, float conf_thresh,
                                                            float nms_thresh, float* res_out, int* counts, int max_rows) {
    std::vector<std::vector<Detection>> rb;
    batch_nms(rb, output_host, batch, output_size, conf_thresh, nms_thresh);
    for (int b = 0; b < batch; ++b) {
        counts[b] = (int)rb[b].size();
        for (size_t i = 0; i < rb[b].size() && (int)i < max_rows; ++i)
            memcpy(res_out + ((size_t)b * max_rows + i) * (sizeof(Detection) / 4), &rb[b][i], sizeof(Detection));
    }
    return 0;
}
// get_rect_adapt_landmark (postprocess.cpp:38-69): box + 17 keypoints (x, y, conf) back to the original image, in place
__attribute__((visibility("default"))) void ref_v8_get_rect_adapt_landmark(int img_w, int img_h, float* bbox, float* lmk, int* rect_out) {
    cv::Mat img(img_h, img_w, CV_8UC3, nullptr);
    cv::Rect r = get_rect_adapt_landmark(img, bbox, lmk);
    rect_out[0] = r.x;
    rect_out[1] = r.y;
    rect_out[2] = r.width;
    rect_out[3] = r.height;
}
// scale_mask (postprocess.cpp:207-226): the crop rectangle it takes out of the 640x640 mask and the size it resizes to
// (recorded by the OpenCV shim; the resize itself is pinned to cv2 in tests/test_oracle_cpu.py)
__attribute__((visibility("default"))) void ref_v8_scale_mask_rect(int img_w, int img_h, int* out6) {
    cv::Mat img(img_h, img_w, CV_8UC3, nullptr);
    cv::Mat mask(kInputH, kInputW, CV_32F, nullptr);
    (void)scale_mask(mask, img);
    for (int i = 0; i < 6; ++i) out6[i] = cv::Mat::shim_last_crop()[i];
}
// process_decode_ptr_host (postprocess.cpp:131-147): rows of the compact buffer with keep flag 1 -> Detection rows
__attribute__((visibility("default"))) int ref_v8_process_decode_ptr_host(const float* decode_ptr_host, int bbox_elem, int count, float* res_out) {
    std::vector<Detection> res;
    cv::Mat img;
    process_decode_ptr_host(res, decode_ptr_host, bbox_elem, img, count);
    for (size_t i = 0; i < res.size(); ++i) memcpy(res_out + i * 6, &res[i], 6 * sizeof(float));
    return (int)res.size();
}
// process_decode_ptr_host_obb (postprocess.cpp:273-290): kept rows incl. the angle (Detection::angle is the struct's last float)
__attribute__((visibility("default"))) int ref_v8_process_decode_ptr_host_obb(const float* decode_ptr_host, int bbox_elem, int count, float* res_out) {
    std::vector<Detection> res;
    cv::Mat img;
    process_decode_ptr_host_obb(res, decode_ptr_host, bbox_elem, img, count);
    for (size_t i = 0; i < res.size(); ++i) {
        memcpy(res_out + i * 7, &res[i], 6 * sizeof(float));
        res_out[i * 7 + 6] = res[i].angle;
    }
    return (int)res.size();
}
// get_rect (postprocess.cpp:4-36): box in network-input pixels -> cv::Rect in the original image (kInputW x kInputH = 640 x 640)
__attribute__((visibility("default"))) void ref_v8_get_rect(int img_w, int img_h, float* bbox, int* rect_out) {
    cv::Mat img(img_h, img_w, CV_8UC3, nullptr);
    cv::Rect r = get_rect(img, bbox);
    rect_out[0] = r.x;
    rect_out[1] = r.y;
    rect_out[2] = r.width;
    rect_out[3] = r.height;
}
}
