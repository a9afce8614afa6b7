// preprocess.cu -- batched letterbox warp-affine + BGR->RGB + /255 + HWC->CHW (+fp16) for sm_100a.
// Replaces warpaffine_kernel / cuda_preprocess / cuda_batch_preprocess, yolov8/src/preprocess.cu:7-127
// (identical bodies in yolov5/7/9/10/11/12/13/26), which launch once per image and synchronise the
// stream after every image (:119-127).  Here: ONE launch for the whole batch, per-image descriptors
// and affine matrices in the kernel parameter block (no H2D copy of metadata), each thread produces 4
// horizontally adjacent destination pixels so that the three planar stores are 128-bit (fp32) or
// 64-bit (fp16) and fully coalesced.
//
// Roofline: HBM-bound; algorithmic bytes per image = src_w*src_h*3 (u8 read once) +
// 3*dst_w*dst_h*sizeof(out) (6 144 000 B for 640x640 -> 640x640 fp32).
//
// Arithmetic follows the reference kernel statement by statement (including its `+0.5f` source
// offset without a matching `-0.5f`, preprocess.cu:22-23) using round-to-nearest intrinsics so the
// result is independent of FMA contraction.
#include "common.cuh"

namespace trtx {

constexpr int kMaxImagesPerLaunch = 128;

struct PreImage {
    const uint8_t* src;
    int sw, sh, pitch;
    float m[6];  // d2s
};
struct PreArgs {
    PreImage img[kMaxImagesPerLaunch];
    int dw, dh;
    int stage_bytes;  // dynamic shared memory given to every block
};

// x / 255.0f, correctly rounded, in 3 FP instructions instead of the IEEE division subroutine:
// q0 = x*r, e = fma(-255, q0, x), q = fma(e, r, q0) with r = RN(1/255).  Verified EXHAUSTIVELY against
// x/255.0f for every float in [0, 65536) (1.2e9 values, tools/verify_div255.c) -- bit-identical.
__device__ __forceinline__ float div255(float x) {
    const float r = 1.0f / 255.0f;
    const float q0 = __fmul_rn(x, r);
    return __fmaf_rn(__fmaf_rn(-255.0f, q0, x), r, q0);
}

// int -> float without the XU (conversion) pipe: for 0 <= i < 2^23, (2^23 + i) is exactly representable with the
// integer sitting in the mantissa, so OR-ing i into the bits of 2^23 and subtracting 2^23 gives float(i) exactly.
// ncu showed the first version of this kernel bound by the quarter-rate XU pipe (12 u8->f32 conversions per
// pixel, sm__inst_executed_pipe_xu at its peak); these run on the full-rate ALU/FMA pipes instead.
__device__ __forceinline__ float u23_to_float(uint32_t i) { return __uint_as_float(0x4B000000u | i) - 8388608.0f; }
__device__ __forceinline__ float ldg_u8f(const uint8_t* p) { return u23_to_float((uint32_t)__ldg(p)); }

__device__ __forceinline__ float lds_u8f(uint32_t saddr) {
    uint32_t v;
    asm volatile("ld.shared.u8 %0, [%1];" : "=r"(v) : "r"(saddr));
    return u23_to_float(v);
}

constexpr int kStageBytesMax = 40 * 1024;  // cap on the staged source band of one block (keeps >= 5 blocks per SM)

// Byte fetch of source pixel channel `c` at byte offset `o` of a row: staged band (shared) or the image (global)
template <bool STAGED>
struct RowRef {
    const uint8_t* g;  // global row base (direct path)
    uint32_t s;        // shared address such that s + x*3 is the staged byte of pixel x (staged path)
    __device__ __forceinline__ float at(int o) const {
        if constexpr (STAGED) return lds_u8f(s + (uint32_t)o);
        else return ldg_u8f(g + o);
    }
};

// One thread = 4 horizontally adjacent destination pixels of one row.  The letterbox matrix has no rotation
// (m[1] = m[3] = -0.0f, preprocess.cu:99-104), so the source row pair, the vertical weights and the row validity are
// computed once per thread; `m3*dx` only contributes a signed zero.
template <typename OutT, bool STAGED>
__device__ __forceinline__ void letterbox_pixels(const PreImage& im, const PreArgs& a, OutT* __restrict__ base, size_t area, int dx0,
                                                 int dy, bool y_out, bool r0ok, bool r1ok, float ly, float hy,
                                                 RowRef<STAGED> row0, RowRef<STAGED> row1) {
    const float cv = 128.0f;  // const_value_st (:115)
    const float ym = __fmul_rn(im.m[1], (float)dy);
    float r[4], g[4], bl[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int dx = dx0 + i;
        float c0 = 0.f, c1 = 0.f, c2 = 0.f;
        if (dx < a.dw) {
            const float src_x = __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(im.m[0], u23_to_float((uint32_t)dx)), ym), im.m[2]), 0.5f);  // :22
            if (y_out || src_x <= -1 || src_x >= im.sw) {
                c0 = c1 = c2 = cv;
            } else {
                const int x_low = (int)floorf(src_x);
                const int x_high = x_low + 1;
                const float lx = __fsub_rn(src_x, u23_to_float((uint32_t)(x_low + 1)) - 1.0f), hx = __fsub_rn(1.0f, lx);  // x_low >= -1
                const float w1 = __fmul_rn(hy, hx), w2 = __fmul_rn(hy, lx), w3 = __fmul_rn(ly, hx), w4 = __fmul_rn(ly, lx);
                const bool xl = x_low >= 0, xh = x_high < im.sw;
                float v1[3], v2[3], v3[3], v4[3];
                if (r0ok && r1ok && xl && xh) {
                    // interior pixel (all but the letterbox border): the two source pixels of a row are 6 contiguous bytes
                    const int o = x_low * 3;
#pragma unroll
                    for (int k = 0; k < 3; ++k) {
                        v1[k] = row0.at(o + k);
                        v2[k] = row0.at(o + 3 + k);
                        v3[k] = row1.at(o + k);
                        v4[k] = row1.at(o + 3 + k);
                    }
                } else {
                    const int o1 = (xl ? x_low : 0) * 3, o2 = (xh ? x_high : 0) * 3;
#pragma unroll
                    for (int k = 0; k < 3; ++k) {
                        v1[k] = (r0ok && xl) ? row0.at(o1 + k) : cv;
                        v2[k] = (r0ok && xh) ? row0.at(o2 + k) : cv;
                        v3[k] = (r1ok && xl) ? row1.at(o1 + k) : cv;
                        v4[k] = (r1ok && xh) ? row1.at(o2 + k) : cv;
                    }
                }
                // :59-61, left-to-right sums
                c0 = __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(w1, v1[0]), __fmul_rn(w2, v2[0])), __fmul_rn(w3, v3[0])), __fmul_rn(w4, v4[0]));
                c1 = __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(w1, v1[1]), __fmul_rn(w2, v2[1])), __fmul_rn(w3, v3[1])), __fmul_rn(w4, v4[1]));
                c2 = __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(w1, v1[2]), __fmul_rn(w2, v2[2])), __fmul_rn(w3, v3[2])), __fmul_rn(w4, v4[2]));
            }
        }
        // bgr -> rgb, /255 (:64-74)
        r[i] = div255(c2);
        g[i] = div255(c1);
        bl[i] = div255(c0);
    }
    const bool vec_ok = (dx0 + 3 < a.dw) && (a.dw % 4 == 0);
    if constexpr (sizeof(OutT) == 4) {
        if (vec_ok) {
            *reinterpret_cast<float4*>(base) = make_float4(r[0], r[1], r[2], r[3]);
            *reinterpret_cast<float4*>(base + area) = make_float4(g[0], g[1], g[2], g[3]);
            *reinterpret_cast<float4*>(base + 2 * area) = make_float4(bl[0], bl[1], bl[2], bl[3]);
        } else {
            for (int i = 0; i < 4 && dx0 + i < a.dw; ++i) {
                base[i] = r[i];
                base[area + i] = g[i];
                base[2 * area + i] = bl[i];
            }
        }
    } else {
        if (vec_ok && (area % 4 == 0)) {
            __half2 r01 = __floats2half2_rn(r[0], r[1]), r23 = __floats2half2_rn(r[2], r[3]);
            __half2 g01 = __floats2half2_rn(g[0], g[1]), g23 = __floats2half2_rn(g[2], g[3]);
            __half2 b01 = __floats2half2_rn(bl[0], bl[1]), b23 = __floats2half2_rn(bl[2], bl[3]);
            uint2 pr = make_uint2(*reinterpret_cast<uint32_t*>(&r01), *reinterpret_cast<uint32_t*>(&r23));
            uint2 pg = make_uint2(*reinterpret_cast<uint32_t*>(&g01), *reinterpret_cast<uint32_t*>(&g23));
            uint2 pb = make_uint2(*reinterpret_cast<uint32_t*>(&b01), *reinterpret_cast<uint32_t*>(&b23));
            *reinterpret_cast<uint2*>(base) = pr;
            *reinterpret_cast<uint2*>(base + area) = pg;
            *reinterpret_cast<uint2*>(base + 2 * area) = pb;
        } else {
            for (int i = 0; i < 4 && dx0 + i < a.dw; ++i) {
                base[i] = __float2half_rn(r[i]);
                base[area + i] = __float2half_rn(g[i]);
                base[2 * area + i] = __float2half_rn(bl[i]);
            }
        }
    }
}

// src coordinate of a destination coordinate, exactly as the per-pixel code computes it
__device__ __forceinline__ float src_coord(float m_a, float m_b, float m_c, int da, int db) {
    return __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(m_a, (float)da), __fmul_rn(m_b, (float)db)), m_c), 0.5f);
}

// Block = 64 x 4 threads = a 256 x 4 destination tile of one image.  The block first stages the source band its tile
// samples from (rows ys0..ys1, pixels xs0..xs1) in shared memory with coalesced 32-bit loads -- every source byte is
// fetched from L2/HBM once per block instead of through 48 twelve-sector byte loads per thread (ncu: 11.8 sectors per
// request, 48x L1 amplification in the first version) -- then every thread bilinearly samples from shared memory, where
// byte reads at a 12-byte lane stride are conflict-free.  Tiles whose band does not fit (very large down-scales) sample
// straight from global memory.
template <typename OutT>
__global__ void __launch_bounds__(256) letterbox_kernel(const __grid_constant__ PreArgs a, OutT* __restrict__ dst,
                                                        int first_image) {
    extern __shared__ __align__(16) unsigned char band[];
    __shared__ int s_geo[8];  // xs0, ys0, rows, stride, staged
    const int b = blockIdx.z;
    const PreImage& im = a.img[b];
    const int tile_x0 = blockIdx.x * blockDim.x * 4, tile_y0 = blockIdx.y * blockDim.y;
    const int tid = threadIdx.y * blockDim.x + threadIdx.x;
    if (tid == 0) {
        const int tx1 = min(a.dw, tile_x0 + (int)blockDim.x * 4) - 1, ty1 = min(a.dh, tile_y0 + (int)blockDim.y) - 1;
        // m[0], m[4] > 0: source coordinates grow with the destination coordinates
        int xs0 = (int)floorf(src_coord(im.m[0], im.m[1], im.m[2], tile_x0, tile_y0));
        int xs1 = (int)floorf(src_coord(im.m[0], im.m[1], im.m[2], tx1, tile_y0)) + 1;
        int ys0 = (int)floorf(src_coord(im.m[3], im.m[4], im.m[5], tile_x0, tile_y0));
        int ys1 = (int)floorf(src_coord(im.m[3], im.m[4], im.m[5], tile_x0, ty1)) + 1;
        xs0 = max(xs0, 0);
        ys0 = max(ys0, 0);
        xs1 = min(xs1, im.sw - 1);
        ys1 = min(ys1, im.sh - 1);
        const int rows = ys1 - ys0 + 1, row_bytes = (xs1 - xs0 + 1) * 3;
        const int stride = ((row_bytes + 3 + 3) & ~3) + 4;  // room for the <= 3 byte misalignment of a row start
        const bool nonempty = rows > 0 && row_bytes > 0;
        s_geo[0] = xs0;
        s_geo[1] = ys0;
        s_geo[2] = nonempty ? rows : 0;
        s_geo[3] = stride;
        s_geo[4] = (!nonempty || (size_t)rows * stride <= (size_t)a.stage_bytes) ? 1 : 0;
        s_geo[5] = row_bytes;
    }
    __syncthreads();
    const int xs0 = s_geo[0], ys0 = s_geo[1], rows = s_geo[2], stride = s_geo[3], row_bytes = s_geo[5];
    const bool staged = s_geo[4] != 0;
    if (staged && rows > 0) {
        const uint8_t* img_end = im.src + (size_t)(im.sh - 1) * im.pitch + (size_t)im.sw * 3;
        const int wpr = stride >> 2;
        for (int idx = tid; idx < rows * wpr; idx += 256) {
            const int r = idx / wpr, wd = idx - r * wpr;
            const uint8_t* g0 = im.src + (size_t)(ys0 + r) * im.pitch + (size_t)xs0 * 3;
            const uint8_t* ga = reinterpret_cast<const uint8_t*>(reinterpret_cast<uintptr_t>(g0) & ~(uintptr_t)3) + 4 * wd;
            if (ga >= g0 + row_bytes) continue;  // past the bytes this row needs
            uint32_t w;
            if (ga >= im.src && ga + 4 <= img_end) {
                w = __ldg(reinterpret_cast<const uint32_t*>(ga));
            } else {  // first / last word of the image buffer: assemble from the bytes that exist
                w = 0;
                for (int k = 0; k < 4; ++k)
                    if (ga + k >= im.src && ga + k < img_end) w |= (uint32_t)__ldg(ga + k) << (8 * k);
            }
            *reinterpret_cast<uint32_t*>(band + (size_t)r * stride + 4 * wd) = w;
        }
    }
    __syncthreads();

    const int dy = tile_y0 + threadIdx.y;
    const int dx0 = tile_x0 + threadIdx.x * 4;
    if (dy >= a.dh || dx0 >= a.dw) return;
    const size_t area = (size_t)a.dw * a.dh;
    OutT* base = dst + (size_t)(first_image + b) * 3 * area + (size_t)dy * a.dw + dx0;

    // preprocess.cu:23 (src_y), evaluated with the thread's first dx: m[3]*dx is +-0 for every dx
    const float src_y = src_coord(im.m[3], im.m[4], im.m[5], dx0, dy);
    const bool y_out = src_y <= -1 || src_y >= im.sh;
    const int y_low = (int)floorf(src_y);
    const int y_high = y_low + 1;
    const float ly = __fsub_rn(src_y, (float)y_low);
    const float hy = __fsub_rn(1.0f, ly);
    const bool r0ok = y_low >= 0, r1ok = y_high < im.sh;
    if (staged) {
        // shared address of "pixel 0" of a source row: band row start + misalignment of the row - xs0*3
        const uint32_t sb = (uint32_t)__cvta_generic_to_shared(band);
        auto row_addr = [&](int y) -> uint32_t {
            const int r = min(max(y - ys0, 0), max(rows - 1, 0));
            const uint32_t mis = (uint32_t)(reinterpret_cast<uintptr_t>(im.src + (size_t)(ys0 + r) * im.pitch + (size_t)xs0 * 3) & 3);
            return sb + (uint32_t)(r * stride) + mis - (uint32_t)(xs0 * 3);
        };
        RowRef<true> q0{nullptr, row_addr(y_low)}, q1{nullptr, row_addr(y_high)};
        letterbox_pixels<OutT, true>(im, a, base, area, dx0, dy, y_out, r0ok, r1ok, ly, hy, q0, q1);
    } else {
        RowRef<false> q0{im.src + (size_t)(r0ok ? y_low : 0) * im.pitch, 0}, q1{im.src + (size_t)(r1ok ? y_high : 0) * im.pitch, 0};
        letterbox_pixels<OutT, false>(im, a, base, area, dx0, dy, y_out, r0ok, r1ok, ly, hy, q0, q1);
    }
}

}  // namespace trtx

using namespace trtx;

extern "C" {

// preprocess.cu:98-110 + cv::invertAffineTransform (CV_32F branch)
TRTX_API void trtx_letterbox_matrix(int sw, int sh, int dw, int dh, float d2s[6]) {
    float a = dh / (float)sh, b = dw / (float)sw;
    float scale = a < b ? a : b;  // std::min
    float s2d[6];
    s2d[0] = scale;
    s2d[1] = 0;
    s2d[2] = (float)(-scale * sw * 0.5 + dw * 0.5);
    s2d[3] = 0;
    s2d[4] = scale;
    s2d[5] = (float)(-scale * sh * 0.5 + dh * 0.5);
    /* cv::invertAffineTransform, CV_32F branch, as OpenCV >= 4.x evaluates it with softfloat/softdouble
     * (pinned against cv2 4.13 on 3000 sizes, tests/test_oracle_cpu.py): the determinant and the
     * A = M*D products are FLOAT operations (softfloat::operator* converts D to float), 1/D and the
     * translation terms are double. */
    float p0 = s2d[0] * s2d[4], p1 = s2d[1] * s2d[3];
    float det = p0 - p1;
    double D = (double)det;
    D = D != 0 ? 1. / D : 0;
    float Df = (float)D;
    float fA11 = s2d[4] * Df, fA22 = s2d[0] * Df, fA12 = (-s2d[1]) * Df, fA21 = (-s2d[3]) * Df;
    double A11 = fA11, A22 = fA22, A12 = fA12, A21 = fA21;
    double b1 = -A11 * s2d[2] - A12 * s2d[5];
    double b2 = -A21 * s2d[2] - A22 * s2d[5];
    d2s[0] = (float)A11;
    d2s[1] = (float)A12;
    d2s[2] = (float)b1;
    d2s[3] = (float)A21;
    d2s[4] = (float)A22;
    d2s[5] = (float)b2;
}

TRTX_API int trtx_preprocess_batch_enqueue(const trtx_image_desc* images_host, int batch, void* dst_dev, int dst_w,
                                           int dst_h, int out_dtype, trtx_stream_t stream) {
    if (!images_host || batch <= 0 || !dst_dev || dst_w <= 0 || dst_h <= 0) return TRTX_ERR_INVALID;
    if (out_dtype != TRTX_F32 && out_dtype != TRTX_F16) return TRTX_ERR_INVALID;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    for (int first = 0; first < batch; first += kMaxImagesPerLaunch) {
        const int n = batch - first < kMaxImagesPerLaunch ? batch - first : kMaxImagesPerLaunch;
        PreArgs a;
        a.dw = dst_w;
        a.dh = dst_h;
        for (int i = 0; i < n; ++i) {
            const trtx_image_desc& d = images_host[first + i];
            if (!d.data_dev || d.width <= 0 || d.height <= 0 || d.pitch < 3 * d.width) return TRTX_ERR_INVALID;
            a.img[i].src = d.data_dev;
            a.img[i].sw = d.width;
            a.img[i].sh = d.height;
            a.img[i].pitch = d.pitch;
            trtx_letterbox_matrix(d.width, d.height, dst_w, dst_h, a.img[i].m);
        }
        // shared memory for the staged band of a 256 x 4 tile: (256/scale + 3) pixels x (4/scale + 3) rows, max over images
        size_t need = 0;
        for (int i = 0; i < n; ++i) {
            const double inv = a.img[i].m[0] > 0 ? (double)a.img[i].m[0] : 1.0;  // d2s scale = source pixels per destination pixel
            const size_t row_bytes = ((size_t)(256 * inv) + 4) * 3;
            const size_t rows = (size_t)(4 * inv) + 4;
            const size_t bytes = rows * (((row_bytes + 6) & ~(size_t)3) + 4);
            if (bytes > need) need = bytes;
        }
        a.stage_bytes = (int)(need < (size_t)kStageBytesMax ? need : (size_t)kStageBytesMax);
        dim3 block(64, 4, 1);
        dim3 grid((dst_w + 4 * 64 - 1) / (4 * 64), (dst_h + 3) / 4, n);
        if (out_dtype == TRTX_F32)
            letterbox_kernel<float><<<grid, block, a.stage_bytes, st>>>(a, static_cast<float*>(dst_dev), first);
        else
            letterbox_kernel<__half><<<grid, block, a.stage_bytes, st>>>(a, static_cast<__half*>(dst_dev), first);
        int rc = check_launch();
        if (rc) return rc;
    }
    return TRTX_OK;
}

}  // extern "C"
