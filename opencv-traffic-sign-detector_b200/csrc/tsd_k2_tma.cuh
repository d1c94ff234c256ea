// tsd_k2_tma.cuh -- K2 (cv2.resize of the crop, DET:123-124,570-572) with the crop staged in shared memory by the Tensor Memory
// Accelerator.
//
// k2_crop_resize_v2 gathers every tap straight from the frame: 12 byte loads per destination pixel row and lane through L1, the
// kernel is bound by the LSU / instruction issue.  Here ONE elected lane per window issues cp.async.bulk.tensor (TMA, SASS UTMALDG)
// for the window's ROI -- boxes of P bytes x 8 rows out of a 3-D tensor map over the frames (uint8 [F][H][3W]; the x coordinate
// is the byte offset 3*x1 rounded down to 16, so the ROI arrives with pitch P and up to 15 bytes of slack in front) -- into the warp's staging buffer behind an
// mbarrier, and the resize runs from shared memory with the SAME arithmetic (11-bit fixed point, x coefficient clamp, y row clip)
// restructured so that the horizontal pass of a source row is computed once and reused by the next destination row when both touch
// it (for the typical 1.4x down-scale 36 instead of 50 row passes).  The 1888-byte window is assembled in shared memory and
// leaves as 128-bit stores.  Crops that do not fit the staging buffer (wider than 85 px or taller than 64 rows: ~1/3 of the
// synthetic windows) take k2_window_gather in the same warp.
#pragma once
#include <cuda.h>
#include "tsd_kernels.cuh"

namespace tsd {

constexpr int kTmaWarps = 4;
__host__ __device__ constexpr int tma_out_bytes(int D) { return (D * D * 3 + 127) & ~127; }      // assembled window rounded up to 128
__host__ __device__ constexpr int tma_warp_bytes(int D) { return kTmaStageBytes + tma_out_bytes(D); }

struct alignas(64) K2TensorMaps { CUtensorMap m[4]; };   // box widths 64, 128, 192, 256 bytes x 8 rows x 1 frame

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_%=:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra DONE_%=;\n"
        "bra WAIT_%=;\n"
        "DONE_%=:\n"
        "}\n" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ void tma_load_3d(void* smem_dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1, int c2) {
    asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
                 ::"r"(smem_u32(smem_dst)), "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2) : "memory");
}

// One horizontal pass: the three channels of destination column `lane` from one staged source row (already >> 4, as cv2's vertical
// pass consumes them).
__device__ __forceinline__ void k2_hrow(const uint8_t* rowp, int xd1, int xa0, int xa1, int (&h)[3]) {
#pragma unroll
    for (int k = 0; k < 3; k++) h[k] = ((int)rowp[k] * xa0 + (int)rowp[xd1 + k] * xa1) >> 4;
}

template <int D>
__global__ void __launch_bounds__(kTmaWarps * 32) k2_crop_resize_tma_kernel(
    const __grid_constant__ K2TensorMaps maps, const uint8_t* __restrict__ frames, int H, int W, int64_t row_stride, int64_t frame_stride,
    const int4* __restrict__ coords, const int32_t* __restrict__ win_frame, const int32_t* __restrict__ n_ptr, int n_max,
    uint8_t* __restrict__ windows, int out_stride) {
    constexpr int C = 3;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    __shared__ int4 s_y[kTmaWarps][32];
    __shared__ __align__(8) uint64_t s_bar[kTmaWarps];
    const int lane = threadIdx.x & 31, wl = threadIdx.x >> 5;
    constexpr int kTmaOutBytes = tma_out_bytes(D);
    // staging buffer, 128-byte aligned for the TMA (the launch asks for 128 spare bytes; the per-warp size is a multiple of 128)
    uint8_t* const sb = smem_raw + ((128u - (smem_u32(smem_raw) & 127u)) & 127u) + (size_t)wl * tma_warp_bytes(D);
    uint8_t* const so = sb + kTmaStageBytes;                               // the assembled window
    uint64_t* const bar = &s_bar[wl];
    if (lane == 0) {
        mbar_init(bar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncwarp();
    uint32_t phase = 0;
    const int n = n_ptr ? min(*n_ptr, n_max) : n_max;
    const int nwarps = (gridDim.x * blockDim.x) >> 5;
    const bool act = lane < D;
    const int li = act ? lane : 0;
    for (int w = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; w < n; w += nwarps) {
        const int4 c = coords[w];
        const int fr = win_frame[w];                         // (issued with the coords load, not behind it)
        const int cx = min(c.x, W), cy = min(c.y, H);
        const int cw = min(c.z, W) - cx, ch = min(c.w, H) - cy;
        // the box starts at the 16-byte boundary at or below the ROI's first byte (an x coordinate that is not a multiple of 16 bytes
        // faulted with "illegal instruction" on B200): `rem` bytes of slack in front of every staged row
        if (!k2_tma_takes(cx, cw, ch)) continue;             // (warp-uniform) wide / tall / empty crops: k2_crop_resize_v2, launched beside this kernel
        const int rem = (cx * C) & 15;
        const int P = ((cw * C + rem + 63) >> 6) << 6, nops = (ch + kTmaBoxRows - 1) / kTmaBoxRows;
        if (lane == 0) {                                     // the ROI travels as nops boxes of P bytes x 8 rows (rows past the frame: zero fill)
            const CUtensorMap* map = &maps.m[(P >> 6) - 1];
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");      // the warp's reads of the previous ROI precede the new copy
            mbar_expect_tx(bar, (uint32_t)(P * kTmaBoxRows * nops));
            for (int i = 0; i < nops; i++) tma_load_3d(sb + i * kTmaBoxRows * P, map, bar, cx * C - rem, cy + i * kTmaBoxRows, fr);
        }
        // coefficient tables while the copy is in flight (float32 rounding as in OpenCV); lane doubles as dx and as dy
        int xs0, xd1, xa0, xa1;
        {
            const double scale = 1.0 / ((double)D / (double)cw);
            float f = (float)(((double)li + 0.5) * scale - 0.5);
            int s = (int)floorf(f); f -= (float)s;
            if (s < 0) { f = 0.f; s = 0; }
            if (s >= cw - 1) { f = 0.f; s = cw - 1; }
            xs0 = s * C;
            xd1 = (min(s + 1, cw - 1) - s) * C;
            xa0 = __float2int_rn((1.f - f) * 2048.f);
            xa1 = __float2int_rn(f * 2048.f);
        }
        {
            const double scale = 1.0 / ((double)D / (double)ch);
            float f = (float)(((double)li + 0.5) * scale - 0.5);
            int s = (int)floorf(f); f -= (float)s;
            s_y[wl][lane] = make_int4(min(max(s, 0), ch - 1), min(max(s + 1, 0), ch - 1), __float2int_rn((1.f - f) * 2048.f), __float2int_rn(f * 2048.f));
        }
        for (int i = D * D * C + lane; i < kTmaOutBytes; i += 32) so[i] = 0;       // zero pad of the internal layout
        __syncwarp();
        mbar_wait(bar, phase);
        phase ^= 1;
        uint8_t* const od = so + li * C;
        if (cw == D && ch == D) {                           // same size: copy
#pragma unroll 5
            for (int dy = 0; dy < D; dy++) {
                const uint8_t* p = sb + rem + dy * P + li * C;
                if (act) { od[dy * D * C] = p[0]; od[dy * D * C + 1] = p[1]; od[dy * D * C + 2] = p[2]; }
            }
        } else if (cw == 2 * D && ch == 2 * D) {            // INTER_AREA 2x2 fast path
#pragma unroll 5
            for (int dy = 0; dy < D; dy++) {
                const uint8_t* q0 = sb + rem + (2 * dy) * P + 2 * li * C;
                const uint8_t* q1 = q0 + P;
                if (act) {
#pragma unroll
                    for (int k = 0; k < C; k++) od[dy * D * C + k] = (uint8_t)((q0[k] + q0[C + k] + q1[k] + q1[C + k] + 2) >> 2);
                }
            }
        } else {
            // general path: the horizontal pass of a source row lives in registers and is reused by the next destination row when
            // both touch it (row indices are warp-uniform, so the branches are too)
            const uint8_t* px = sb + rem + xs0;
            int cur0 = -1, cur1 = -1, h0[3] = {0, 0, 0}, h1[3] = {0, 0, 0};
#pragma unroll 5
            for (int dy = 0; dy < D; dy++) {
                const int4 yc = s_y[wl][dy];
                const int r0 = yc.x, r1 = yc.y, b0 = yc.z, b1 = yc.w;
                if (r0 != cur0) {
                    if (r0 == cur1) { h0[0] = h1[0]; h0[1] = h1[1]; h0[2] = h1[2]; }
                    else k2_hrow(px + r0 * P, xd1, xa0, xa1, h0);
                    cur0 = r0;
                }
                if (r1 != cur1) {
                    if (r1 == cur0) { h1[0] = h0[0]; h1[1] = h0[1]; h1[2] = h0[2]; }
                    else k2_hrow(px + r1 * P, xd1, xa0, xa1, h1);
                    cur1 = r1;
                }
                if (act) {
#pragma unroll
                    for (int k = 0; k < C; k++) od[dy * D * C + k] = (uint8_t)((((b0 * h0[k]) >> 16) + ((b1 * h1[k]) >> 16) + 2) >> 2);
                }
            }
        }
        __syncwarp();
        {                                                    // the assembled window (zero pad included) as 128-bit stores
            const uint4* s4 = reinterpret_cast<const uint4*>(so);
            uint4* g4 = reinterpret_cast<uint4*>(windows + (int64_t)w * out_stride);
            for (int i = lane; i < (out_stride >> 4); i += 32) g4[i] = s4[i];
        }
        __syncwarp();                                        // staging buffer and window are free for the next ROI
    }
}

// cuTensorMapEncodeTiled through the runtime's driver entry point (the library links cudart statically and not libcuda).
typedef CUresult (*tsd_encode_tiled_fn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                        const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

// maps over frames uint8 [F][H][3W] (strides in bytes): false if the layout does not satisfy TMA's rules (16-byte aligned base and
// strides) or the driver refuses; the caller then uses k2_crop_resize_v2.
static bool k2_build_tensor_maps(K2TensorMaps* out, const uint8_t* frames, int nframes, int H, int W, int64_t row_stride, int64_t frame_stride) {
    static tsd_encode_tiled_fn encode = nullptr;
    static bool tried = false;
    if (!tried) {
        tried = true;
        void* fn = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess) encode = (tsd_encode_tiled_fn)fn;
        cudaGetLastError();
    }
    if (!encode || ((uintptr_t)frames & 15) || (row_stride & 15) || (frame_stride & 15) || nframes < 1) return false;
    const cuuint64_t dims[3] = {(cuuint64_t)W * 3, (cuuint64_t)H, (cuuint64_t)nframes};
    const cuuint64_t strides[2] = {(cuuint64_t)row_stride, (cuuint64_t)frame_stride};
    const cuuint32_t estr[3] = {1, 1, 1};
    for (int i = 0; i < 4; i++) {
        const cuuint32_t box[3] = {(cuuint32_t)(64 * (i + 1)), (cuuint32_t)kTmaBoxRows, 1};
        if (encode(&out->m[i], CU_TENSOR_MAP_DATA_TYPE_UINT8, 3, (void*)frames, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
            return false;
    }
    return true;
}

}  // namespace tsd
