// ref_wrap_rcnn.cu -- TEST INFRASTRUCTURE.  extern "C" entry points around the REFERENCE's Faster R-CNN free
// functions rpnDecode / rpnNms / predictorDecode / batchedNms (rcnn/*.cu), compiled from /root/reference.
// Each wrapper does the "null workspace -> size" query, allocates the workspace and runs the function.
#include <cuda_runtime_api.h>

#include <vector>

#include "BatchedNmsPlugin.h"
#include "MaskRcnnInferencePlugin.h"
#include "PredictorDecodePlugin.h"
#include "RoiAlignPlugin.h"
#include "RpnDecodePlugin.h"
#include "RpnNmsPlugin.h"

extern "C" {
#define REF_API __attribute__((visibility("default")))

REF_API int ref_rpn_decode(int batch, const float* scores_dev, const float* deltas_dev, float* out_scores_dev,
                           float* out_boxes_dev, int height, int width, int image_height, int image_width, float stride,
                           const float* anchors, int num_anchors, int top_n) {
    std::vector<float> anc(anchors, anchors + num_anchors * 4);
    const void* ins[2] = {scores_dev, deltas_dev};
    void* outs[2] = {out_scores_dev, out_boxes_dev};
    int ws = nvinfer1::rpnDecode(batch, nullptr, nullptr, height, width, image_height, image_width, stride, anc, top_n,
                                 nullptr, 0, nullptr);
    void* w = nullptr;
    cudaMalloc(&w, ws);
    int rc = nvinfer1::rpnDecode(batch, ins, outs, height, width, image_height, image_width, stride, anc, top_n, w, ws, 0);
    cudaError_t e = cudaDeviceSynchronize();
    cudaFree(w);
    return rc != 0 ? rc : (int)e;
}

REF_API int ref_rpn_nms(int batch, const float* scores_dev, const float* boxes_dev, float* out_boxes_dev, int pre, int post,
                        float thresh) {
    const void* ins[2] = {scores_dev, boxes_dev};
    void* outs[1] = {out_boxes_dev};
    int ws = nvinfer1::rpnNms(batch, nullptr, nullptr, pre, post, thresh, nullptr, 0, nullptr);
    void* w = nullptr;
    cudaMalloc(&w, ws);
    int rc = nvinfer1::rpnNms(batch, ins, outs, pre, post, thresh, w, ws, 0);
    cudaError_t e = cudaDeviceSynchronize();
    cudaFree(w);
    return rc != 0 ? rc : (int)e;
}

REF_API int ref_predictor_decode(int batch, const float* scores_dev, const float* deltas_dev, const float* props_dev,
                                 float* out_scores_dev, float* out_boxes_dev, float* out_classes_dev, int num_boxes,
                                 int num_classes, int image_height, int image_width, const float* w4) {
    std::vector<float> wv(w4, w4 + 4);
    const void* ins[3] = {scores_dev, deltas_dev, props_dev};
    void* outs[3] = {out_scores_dev, out_boxes_dev, out_classes_dev};
    int ws = nvinfer1::predictorDecode(batch, nullptr, nullptr, num_boxes, num_classes, image_height, image_width, wv,
                                       nullptr, 0, nullptr);
    void* w = nullptr;
    cudaMalloc(&w, ws);
    int rc = nvinfer1::predictorDecode(batch, ins, outs, num_boxes, num_classes, image_height, image_width, wv, w, ws, 0);
    cudaError_t e = cudaDeviceSynchronize();
    cudaFree(w);
    return rc != 0 ? rc : (int)e;
}

REF_API int ref_batched_nms(int method, int batch, const float* scores_dev, const float* boxes_dev, const float* classes_dev,
                            float* out_scores_dev, float* out_boxes_dev, float* out_classes_dev, int count, int dets,
                            float thresh) {
    const void* ins[3] = {scores_dev, boxes_dev, classes_dev};
    void* outs[3] = {out_scores_dev, out_boxes_dev, out_classes_dev};
    int ws = nvinfer1::batchedNms(method, batch, nullptr, nullptr, count, dets, thresh, nullptr, 0, nullptr);
    void* w = nullptr;
    cudaMalloc(&w, ws);
    int rc = nvinfer1::batchedNms(method, batch, ins, outs, count, dets, thresh, w, ws, 0);
    cudaError_t e = cudaDeviceSynchronize();
    cudaFree(w);
    return rc != 0 ? rc : (int)e;
}

// roiAlign (rcnn/RoiAlign.cu:150-183: one launch + cudaDeviceSynchronize per image)
REF_API int ref_roi_align(int batch, const float* rois_dev, const float* features_dev, float* out_dev, int pooler_resolution,
                          float spatial_scale, int sampling_ratio, int num_proposals, int out_channels, int feature_h,
                          int feature_w) {
    const void* ins[2] = {rois_dev, features_dev};
    void* outs[1] = {out_dev};
    int rc = nvinfer1::roiAlign(batch, ins, outs, pooler_resolution, spatial_scale, sampling_ratio, num_proposals, out_channels,
                                feature_h, feature_w, 0);
    cudaError_t e = cudaDeviceSynchronize();
    return rc != 0 ? rc : (int)e;
}

// maskRcnnInference (rcnn/MaskRcnnInference.cu:35-63)
REF_API int ref_mask_rcnn_inference(int batch, const float* indices_dev, const float* masks_dev, float* out_dev,
                                    int detections_per_im, int output_size, int num_classes) {
    const void* ins[2] = {indices_dev, masks_dev};
    void* outs[1] = {out_dev};
    int rc = nvinfer1::maskRcnnInference(batch, ins, outs, detections_per_im, output_size, num_classes, 0);
    cudaError_t e = cudaDeviceSynchronize();
    return rc != 0 ? rc : (int)e;
}
}
