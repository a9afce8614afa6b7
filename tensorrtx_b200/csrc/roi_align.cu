// roi_align.cu -- RoIAlign + MaskRcnnInference for sm_100a.
// Replaces roiAlign / RoIAlignForward (rcnn/RoiAlign.cu:29-183) and maskRcnnInference / MaskRcnnInferenceKernel
// (rcnn/MaskRcnnInference.cu:8-63), which launch once per image and call cudaDeviceSynchronize() after every launch
// (RoiAlign.cu:178, MaskRcnnInference.cu:59) -- a full-device stall inside enqueue().  Here: one launch for the whole
// batch on the caller's stream, no synchronisation, no allocation.
//
// RoIAlign: a CTA owns one proposal and a chunk of channels.  The sample coordinates of RoIAlign are separable --
// (y_low, y_high, ly, hy, in-range) depends on (ph, iy) only, the x part on (pw, ix) only -- so the CTA tabulates them
// once in shared memory (<= 14*g entries each) instead of recomputing them for each of the 1024 channels; a thread
// then owns one output bin and walks its channels: 4 taps per sample from the L2-resident feature map, and the 196
// outputs of a (proposal, channel) pair are stored contiguously (coalesced).
// Roofline: HBM-write-bound: N*C*P*P*4 bytes per image (803 MB at N=1000, C=1024, P=14).
//
// Arithmetic mirrors the reference kernel statement by statement and is left to nvcc's default FMA contraction,
// exactly like the reference's own build (tests/test_vs_reference_gpu.py compares against the reference kernel).
#include "common.cuh"

namespace trtx {

constexpr int kRoiMaxPooled = 32;   // pooler_resolution
constexpr int kRoiMaxTable = 256;   // pooler_resolution * sampling grid, per axis (larger grids: coordinates on the fly)
constexpr int kRoiChannelsPerCta = 64;

struct RoiArgs {
    const float4* rois;  // [B, N]
    const float* feat;   // [B, C, H, W]
    float* out;          // [B, N, C, P, P]
    int N, C, H, W, P, sampling_ratio;
    float spatial_scale;
};

struct AxisTap {
    int low, high;  // clamped indices
    float l, h;     // fractions
    int valid;      // 0: the reference's bilinear_interpolate returns 0 for this coordinate
};

// one axis of bilinear_interpolate (RoiAlign.cu:29-69); `size` = height or width
__device__ __forceinline__ AxisTap axis_tap(float y, int size) {
    AxisTap t;
    t.valid = !(y < -1.0 || y > size);
    if (y <= 0) y = 0;
    int y_low = static_cast<int>(y);
    int y_high;
    if (y_low >= size - 1) {
        y_high = y_low = size - 1;
        y = (float)y_low;
    } else {
        y_high = y_low + 1;
    }
    t.low = y_low;
    t.high = y_high;
    t.l = y - y_low;
    t.h = 1. - t.l;
    return t;
}

__global__ void __launch_bounds__(256) roi_align_kernel(const __grid_constant__ RoiArgs a) {
    __shared__ AxisTap s_y[kRoiMaxTable], s_x[kRoiMaxTable];
    const int n = blockIdx.x, b = blockIdx.z;
    const int c_begin = blockIdx.y * kRoiChannelsPerCta, c_end = min(a.C, c_begin + kRoiChannelsPerCta);
    const int P = a.P, PP = P * P;
    const float4 roi = a.rois[(size_t)b * a.N + n];

    // RoiAlign.cu:104-127 -- "Do not using rounding; this implementation detail is critical"
    float roi_offset = 0.5f;
    float roi_start_w = roi.x * a.spatial_scale - roi_offset;
    float roi_start_h = roi.y * a.spatial_scale - roi_offset;
    float roi_end_w = roi.z * a.spatial_scale - roi_offset;
    float roi_end_h = roi.w * a.spatial_scale - roi_offset;
    float roi_width = roi_end_w - roi_start_w;
    float roi_height = roi_end_h - roi_start_h;
    float bin_size_h = static_cast<float>(roi_height) / static_cast<float>(P);
    float bin_size_w = static_cast<float>(roi_width) / static_cast<float>(P);
    int roi_bin_grid_h = (a.sampling_ratio > 0) ? a.sampling_ratio : ceil(roi_height / P);
    int roi_bin_grid_w = (a.sampling_ratio > 0) ? a.sampling_ratio : ceil(roi_width / P);
    const float count = roi_bin_grid_h * roi_bin_grid_w;
    const bool tabled = roi_bin_grid_h > 0 && roi_bin_grid_w > 0 && P * roi_bin_grid_h <= kRoiMaxTable &&
                        P * roi_bin_grid_w <= kRoiMaxTable;
    if (tabled) {
        for (int i = threadIdx.x; i < P * roi_bin_grid_h; i += blockDim.x) {
            const int ph = i / roi_bin_grid_h, iy = i - ph * roi_bin_grid_h;
            const float y = roi_start_h + ph * bin_size_h +
                            static_cast<float>(iy + .5f) * bin_size_h / static_cast<float>(roi_bin_grid_h);  // :133-135
            s_y[i] = axis_tap(y, a.H);
        }
        for (int i = threadIdx.x; i < P * roi_bin_grid_w; i += blockDim.x) {
            const int pw = i / roi_bin_grid_w, ix = i - pw * roi_bin_grid_w;
            const float x = roi_start_w + pw * bin_size_w +
                            static_cast<float>(ix + .5f) * bin_size_w / static_cast<float>(roi_bin_grid_w);  // :137-139
            s_x[i] = axis_tap(x, a.W);
        }
    }
    __syncthreads();

    // thread = (bin, channel lane): bins are the fast index, so a warp writes consecutive floats
    const int nsub = max(1, (int)blockDim.x / PP);
    const int bin = threadIdx.x % PP, csub = threadIdx.x / PP;
    if (csub >= nsub) return;
    const int ph = bin / P, pw = bin - ph * P;
    const size_t plane = (size_t)a.H * a.W;
    const float* fb = a.feat + (size_t)b * a.C * plane;
    float* ob = a.out + (((size_t)b * a.N + n) * a.C) * PP + bin;
    constexpr int CU = 4;  // channels per thread and pass: the taps (shared-memory table reads, weights) are channel-independent
    for (int c0 = c_begin + csub * CU; c0 < c_end; c0 += nsub * CU) {
        float output_val[CU];
#pragma unroll
        for (int k = 0; k < CU; ++k) output_val[k] = 0.f;
        const float* fm = fb + (size_t)c0 * plane;
        const int nch = min(CU, c_end - c0);
        for (int iy = 0; iy < roi_bin_grid_h; iy++) {
            AxisTap ty;
            if (tabled) {
                ty = s_y[ph * roi_bin_grid_h + iy];
            } else {  // degenerate or huge sampling grids: coordinates on the fly, as the reference does
                const float y = roi_start_h + ph * bin_size_h +
                                static_cast<float>(iy + .5f) * bin_size_h / static_cast<float>(roi_bin_grid_h);
                ty = axis_tap(y, a.H);
            }
            for (int ix = 0; ix < roi_bin_grid_w; ix++) {
                AxisTap tx;
                if (tabled) {
                    tx = s_x[pw * roi_bin_grid_w + ix];
                } else {
                    const float x = roi_start_w + pw * bin_size_w +
                                    static_cast<float>(ix + .5f) * bin_size_w / static_cast<float>(roi_bin_grid_w);
                    tx = axis_tap(x, a.W);
                }
                if (!(ty.valid && tx.valid)) continue;  // bilinear_interpolate returns 0: `output_val += 0`
                const float ly = ty.l, hy = ty.h, lx = tx.l, hx = tx.h;
                const float w1 = hy * hx, w2 = hy * lx, w3 = ly * hx, w4 = ly * lx;
                const int o1 = ty.low * a.W + tx.low, o2 = ty.low * a.W + tx.high, o3 = ty.high * a.W + tx.low,
                          o4 = ty.high * a.W + tx.high;
#pragma unroll
                for (int k = 0; k < CU; ++k) {
                    if (k < nch) {
                        const float* f = fm + (size_t)k * plane;
                        const float v1 = __ldg(f + o1), v2 = __ldg(f + o2), v3 = __ldg(f + o3), v4 = __ldg(f + o4);
                        const float val = w1 * v1 + w2 * v2 + w3 * v3 + w4 * v4;  // :76
                        output_val[k] += val;
                    }
                }
            }
        }
#pragma unroll
        for (int k = 0; k < CU; ++k)
            if (k < nch) ob[(size_t)(c0 + k) * PP] = output_val[k] / count;  // :147 (NaN for a degenerate proposal, like the reference)
    }
}

// MaskRcnnInferenceKernel (MaskRcnnInference.cu:8-33): out[d] = sigmoid(masks[d, class(d)]) -- one thread per OUTPUT
// element instead of one per (detection, class, pixel); detections whose class index is outside [0, num_classes)
// are left unwritten, as in the reference.
__global__ void mask_rcnn_inference_kernel(const float* __restrict__ indices, const float* __restrict__ masks,
                                           float* __restrict__ out, int total_dets, int num_classes, int ss) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (size_t)total_dets * ss) return;
    const int ind = (int)(i / ss), px = (int)(i - (size_t)ind * ss);
    const int ind_class = indices[ind];  // float -> int, :19
    if (ind_class < 0 || ind_class >= num_classes) return;
    const float maskVal = masks[((size_t)ind * num_classes + ind_class) * ss + px];
    out[i] = logist(maskVal);
}

}  // namespace trtx

using namespace trtx;

extern "C" {

TRTX_API int trtx_roi_align(int batch, const float* rois_dev, const float* features_dev, float* out_dev, int pooler_resolution,
                            float spatial_scale, int sampling_ratio, int num_proposals, int out_channels, int feature_h,
                            int feature_w, trtx_stream_t stream) {
    if (batch <= 0 || !rois_dev || !features_dev || !out_dev) return TRTX_ERR_INVALID;
    if (pooler_resolution <= 0 || num_proposals <= 0 || out_channels <= 0 || feature_h <= 0 || feature_w <= 0) return TRTX_ERR_INVALID;
    if (pooler_resolution > kRoiMaxPooled || pooler_resolution * pooler_resolution > 256) return TRTX_ERR_UNSUPPORTED;
    if (batch > 65535) return TRTX_ERR_UNSUPPORTED;
    RoiArgs a;
    a.rois = reinterpret_cast<const float4*>(rois_dev);
    a.feat = features_dev;
    a.out = out_dev;
    a.N = num_proposals;
    a.C = out_channels;
    a.H = feature_h;
    a.W = feature_w;
    a.P = pooler_resolution;
    a.sampling_ratio = sampling_ratio;
    a.spatial_scale = spatial_scale;
    dim3 grid(num_proposals, (out_channels + kRoiChannelsPerCta - 1) / kRoiChannelsPerCta, batch);
    if (grid.y > 65535) return TRTX_ERR_UNSUPPORTED;
    roi_align_kernel<<<grid, 256, 0, static_cast<cudaStream_t>(stream)>>>(a);
    return check_launch();
}

TRTX_API int trtx_mask_rcnn_inference(int batch, const float* indices_dev, const float* masks_dev, float* out_dev,
                                      int detections_per_im, int output_size, int num_classes, trtx_stream_t stream) {
    if (batch <= 0 || !indices_dev || !masks_dev || !out_dev) return TRTX_ERR_INVALID;
    if (detections_per_im <= 0 || output_size <= 0 || num_classes <= 0) return TRTX_ERR_INVALID;
    const int ss = output_size * output_size;
    const long long total = (long long)batch * detections_per_im * ss;
    const long long blocks = (total + 255) / 256;
    if (blocks > 0x7fffffffLL) return TRTX_ERR_UNSUPPORTED;
    mask_rcnn_inference_kernel<<<(unsigned)blocks, 256, 0, static_cast<cudaStream_t>(stream)>>>(
        indices_dev, masks_dev, out_dev, batch * detections_per_im, num_classes, ss);
    return check_launch();
}

}  // extern "C"
