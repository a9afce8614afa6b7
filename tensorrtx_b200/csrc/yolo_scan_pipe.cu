// yolo_scan_pipe.cu -- the YoloLayer scan as a persistent, TMA-fed pipeline (sm_100a).
//
// One CTA per SM, resident for the whole launch.  A producer thread streams tiles of the level tensors
// (all C channel rows x 128 anchors, 43 KB for YOLOv8 fp32) into a ring of shared-memory stages with ONE
// TMA tensor copy per tile (cp.async.bulk.tensor.3d over a [B, C, g] tensor map, box [1, C, 128],
// completion counted on an mbarrier -- the TMA engine moves the bytes, no registers are tied up and
// ~170 KB per SM are in flight from the first cycle; columns past the end of a level are zero-filled
// by the engine).  Every stage has its OWN consumer warp: warp w waits for stage w, scans all class
// rows of its 128 anchors out of shared memory (conflict-free 128-bit LDS, 4 anchors per lane), gates,
// compacts with a warp scan, decodes the boxes from the 4 box rows that are already in the stage, writes
// the 32-byte candidate records and releases the stage.  Consumers never synchronise with each other, so
// the stages are processed concurrently and the TMA latency of one stage hides behind the others.
//
// Two earlier layouts are kept in the history for the record (profiles/r01b_sweep.log, r01c_sweep.log):
// one 512-byte cp.async.bulk per channel row (84 descriptors per tile throttle the copy engine to
// 1.7 TB/s) and one tensor copy per tile but all consumer warps cooperating on each tile (tiles are
// then processed strictly one after another and the per-tile latency chain caps it at 2.2 TB/s).
// tools/tma_bench.cu measures the copy-engine ceiling of this access pattern: 5.3 TB/s.
//
// The arithmetic is the same bit-exact running-max scheme as yolo_decode.cu::scan_classes.
// HBM traffic = algorithmic bytes (every row is read exactly once); candidates add <= 4%.
#include <cuda.h>  // CUtensorMap (types only; the encoder is fetched through cudaGetDriverEntryPoint)

#include "yolo_layout.cuh"

namespace trtx {

constexpr int kTileAnchors = 128;
constexpr int kMaxStages = 8;

// ---- mbarrier / bulk-copy PTX ----------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t tx) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(tx) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    const uint32_t addr = smem_u32(bar);
    uint32_t done;
    do {
        asm volatile(
                "{\n"
                ".reg .pred p;\n"
                "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
                "selp.u32 %0, 1, 0, p;\n"
                "}\n"
                : "=r"(done)
                : "r"(addr), "r"(parity)
                : "memory");
    } while (!done);
}
// TMA: global [B, C, g] box -> shared, bytes counted on `bar`
__device__ __forceinline__ void tma_load_3d(void* dst_smem, const CUtensorMap* map, int c0, int c1, int c2, uint64_t* bar) {
    asm volatile(
            "cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];" ::"r"(
                    smem_u32(dst_smem)),
            "l"(map), "r"(c0), "r"(c1), "r"(c2), "r"(smem_u32(bar))
            : "memory");
}
struct alignas(64) TmaMaps {
    CUtensorMap m[TRTX_MAX_LEVELS];
};
template <typename T>
__device__ __forceinline__ float4 lds4(const T* p);
template <>
__device__ __forceinline__ float4 lds4<float>(const float* p) {
    return *reinterpret_cast<const float4*>(p);
}
template <>
__device__ __forceinline__ float4 lds4<__half>(const __half* p) {
    const uint2 u = *reinterpret_cast<const uint2*>(p);
    const float2 a = __half22float2(*reinterpret_cast<const __half2*>(&u.x));
    const float2 b = __half22float2(*reinterpret_cast<const __half2*>(&u.y));
    return make_float4(a.x, a.y, b.x, b.y);
}

struct TileRef {
    int b, t, l, tile, col0, ncols;
};
__device__ __forceinline__ TileRef tile_ref(const YoloArgs& a, int T) {
    TileRef r;
    r.b = T / a.tiles_per_image;
    r.t = T - r.b * a.tiles_per_image;
    int l = 0;
    while (l + 1 < a.num_levels && r.t >= a.lv[l + 1].tile_begin) ++l;
    r.l = l;
    r.tile = r.t - a.lv[l].tile_begin;
    r.col0 = r.tile * kTileAnchors;
    r.ncols = min(kTileAnchors, a.lv[l].g - r.col0);
    return r;
}

template <typename T>
__global__ void __launch_bounds__(32 * (kMaxStages + 1), 1)
        yolo_v8_scan_pipe_kernel(const __grid_constant__ YoloArgs a, const __grid_constant__ TmaMaps maps, int total_tiles,
                                 int stages, int stage_bytes) {
    extern __shared__ __align__(128) unsigned char smem[];
    unsigned char* stage_base = smem;
    uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + (size_t)stages * stage_bytes);
    uint64_t* empty_bar = full_bar + kMaxStages;

    const int lane = threadIdx.x & 31;
    const int warp = threadIdx.x >> 5;
    if (threadIdx.x == 0) {
        for (int s = 0; s < stages; ++s) {
            mbar_init(&full_bar[s], 1);   // producer's arrive.expect_tx
            mbar_init(&empty_bar[s], 1);  // the stage's consumer warp
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();

    if (warp == stages) {
        // ------------------------------ producer (one elected thread) ------------------------------
        if (lane == 0) {
            int it = 0;
            for (int Tg = blockIdx.x; Tg < total_tiles; Tg += gridDim.x, ++it) {
                const int s = it % stages;
                const uint32_t ph = (uint32_t)(it / stages) & 1u;
                mbar_wait(&empty_bar[s], ph ^ 1u);
                const TileRef r = tile_ref(a, Tg);
                unsigned char* dst = stage_base + (size_t)s * stage_bytes;
                mbar_arrive_expect_tx(&full_bar[s], (uint32_t)stage_bytes);  // OOB columns are zero-filled and counted
                for (int ch = 0; ch < a.C; ch += 256)                          // TMA box dims are limited to 256
                    tma_load_3d(dst + (size_t)ch * kTileAnchors * sizeof(T), &maps.m[r.l], r.col0, ch, r.b, &full_bar[s]);
            }
        }
        return;
    }

    // ------------------------------ consumer warp `warp` owns stage `warp` ------------------------------
    const int s = warp;
    const T* tile = reinterpret_cast<const T*>(stage_base + (size_t)s * stage_bytes);
    int it = s;
    for (int Tg = blockIdx.x + s * gridDim.x; Tg < total_tiles; Tg += stages * gridDim.x, it += stages) {
        const uint32_t ph = (uint32_t)(it / stages) & 1u;
        const TileRef r = tile_ref(a, Tg);
        const LevelArg& L = a.lv[r.l];
        const bool active = lane * 4 < r.ncols;

        mbar_wait(&full_bar[s], ph);
        Best<4> st;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            st.bx[j] = a.x_lo;
            st.bp[j] = 0.0f;
            st.bc[j] = 0;
        }
        if (active) {
            const T* p = tile + (size_t)4 * kTileAnchors + lane * 4;
            constexpr int U = 8;
            int c = 0;
            for (; c + U <= a.nc; c += U, p += U * kTileAnchors) {
                float4 v[U];
#pragma unroll
                for (int u = 0; u < U; ++u) v[u] = lds4<T>(p + u * kTileAnchors);
#pragma unroll
                for (int u = 0; u < U; ++u) {
                    const bool any = (v[u].x > st.bx[0]) | (v[u].y > st.bx[1]) | (v[u].z > st.bx[2]) | (v[u].w > st.bx[3]);
                    if (any) {
                        update_one<4>(st, 0, v[u].x, c + u);
                        update_one<4>(st, 1, v[u].y, c + u);
                        update_one<4>(st, 2, v[u].z, c + u);
                        update_one<4>(st, 3, v[u].w, c + u);
                    }
                }
            }
            for (; c < a.nc; ++c, p += kTileAnchors) {
                const float4 v = lds4<T>(p);
                update_one<4>(st, 0, v.x, c);
                update_one<4>(st, 1, v.y, c);
                update_one<4>(st, 2, v.z, c);
                update_one<4>(st, 3, v.w, c);
            }
        }
        unsigned flags = 0;
#pragma unroll
        for (int j = 0; j < 4; ++j)
            if (active && !(st.bp[j] < a.gate)) flags |= 1u << j;  // yololayer.cu:203
        int total;
        int off = warp_excl_scan(__popc(flags), lane, &total);
        if (lane == 0) a.tile_count[(size_t)r.b * a.tiles_per_image + r.t] = total;
        if (flags) {
            const float4 d0 = lds4<T>(tile + 0 * kTileAnchors + lane * 4);
            const float4 d1 = lds4<T>(tile + 1 * kTileAnchors + lane * 4);
            const float4 d2 = lds4<T>(tile + 2 * kTileAnchors + lane * 4);
            const float4 d3 = lds4<T>(tile + 3 * kTileAnchors + lane * 4);
            const float dd[4][4] = {{d0.x, d0.y, d0.z, d0.w}, {d1.x, d1.y, d1.z, d1.w},
                                    {d2.x, d2.y, d2.z, d2.w}, {d3.x, d3.y, d3.z, d3.w}};
            const size_t slot0 = (size_t)r.b * a.slots_per_image + L.slot_begin + (size_t)r.col0;
            const float fs = (float)L.stride;
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                if (flags & (1u << j)) {
                    const int e = r.col0 + lane * 4 + j;
                    const int row = e / L.gw, col = e - row * L.gw;
                    // yololayer.cu:217-220
                    const float x1 = ((float)col + 0.5f - dd[0][j]) * fs;
                    const float y1 = ((float)row + 0.5f - dd[1][j]) * fs;
                    const float x2 = ((float)col + 0.5f + dd[2][j]) * fs;
                    const float y2 = ((float)row + 0.5f + dd[3][j]) * fs;
                    store_record(a.cand, slot0 + off, x1, y1, x2, y2, st.bp[j], st.bc[j], L.slot_begin + e);
                    ++off;
                }
            }
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(&empty_bar[s]);  // stage free: the producer may refill it
    }
}

static int g_pipe_max_stages = kMaxStages;  // tuning knob 3: cap on stages (= consumer warps)

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static EncodeTiledFn get_encoder() {
    static EncodeTiledFn fn = nullptr;  // process-wide driver entry point; resolving it twice is harmless
    if (!fn) {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
            q == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<EncodeTiledFn>(p);
    }
    return fn;
}

template <typename T>
static int launch_pipe(const YoloArgs& a, const YoloLayout& L, int batch, cudaStream_t st) {
    EncodeTiledFn enc = get_encoder();
    if (!enc) return TRTX_ERR_UNSUPPORTED;
    int dev = 0, sms = 0, max_smem = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    cudaDeviceGetAttribute(&max_smem, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev);
    const int stage_bytes = a.C * kTileAnchors * (int)sizeof(T);
    const int fixed = 2 * kMaxStages * (int)sizeof(uint64_t);
    int stages = g_pipe_max_stages < kMaxStages ? g_pipe_max_stages : kMaxStages;
    while (stages >= 2 && (size_t)stages * stage_bytes + fixed + 128 > (size_t)max_smem) --stages;
    if (stages < 2) return TRTX_ERR_UNSUPPORTED;
    const size_t smem = (size_t)stages * stage_bytes + fixed;
    TmaMaps maps;
    memset(&maps, 0, sizeof(maps));
    for (int l = 0; l < a.num_levels; ++l) {
        const cuuint64_t gdim[3] = {(cuuint64_t)a.lv[l].g, (cuuint64_t)a.C, (cuuint64_t)batch};
        const cuuint64_t gstr[2] = {(cuuint64_t)a.lv[l].g * sizeof(T), (cuuint64_t)a.C * a.lv[l].g * sizeof(T)};
        const cuuint32_t box[3] = {(cuuint32_t)kTileAnchors, (cuuint32_t)(a.C < 256 ? a.C : 256), 1u};
        const cuuint32_t estr[3] = {1u, 1u, 1u};
        const CUresult r = enc(&maps.m[l], sizeof(T) == 4 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32 : CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 3,
                               const_cast<void*>(a.lv[l].in), gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                               CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS) return TRTX_ERR_UNSUPPORTED;  // caller falls back to the register-path scan
    }
    const int total_tiles = batch * L.tiles_per_image;
    const int grid = total_tiles < sms ? total_tiles : sms;
    auto kern = yolo_v8_scan_pipe_kernel<T>;
    cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    kern<<<grid, 32 * (stages + 1), smem, st>>>(a, maps, total_tiles, stages, stage_bytes);
    return check_launch();
}

void yolo_pipe_set_consumers(int n) { g_pipe_max_stages = n < 2 ? 2 : n; }

int yolo_scan_pipe_launch(const YoloArgs& a, const YoloLayout& L, int in_dtype, int batch, cudaStream_t st) {
    if (a.variant != TRTX_YOLO_V8 || L.vec != 4 || L.tile_cells != kTileAnchors) return TRTX_ERR_UNSUPPORTED;
    // TMA needs 16-byte aligned bases and strides: fp32 rows are (g % 4 == 0 => ok); fp16 rows need g % 8 == 0
    for (int l = 0; l < a.num_levels; ++l) {
        if (reinterpret_cast<uintptr_t>(a.lv[l].in) % 16 != 0) return TRTX_ERR_UNSUPPORTED;
        if (in_dtype == TRTX_F16 && a.lv[l].g % 8 != 0) return TRTX_ERR_UNSUPPORTED;
    }
    if (a.C > 256 && (a.C % 256) != 0) return TRTX_ERR_UNSUPPORTED;  // keep the expect_tx byte count exact
    if (in_dtype == TRTX_F32) return launch_pipe<float>(a, L, batch, st);
    return launch_pipe<__half>(a, L, batch, st);
}

}  // namespace trtx
