// tools/pcie_patterns.cu -- what shape of SM-issued read of PAGE-LOCKED HOST memory gets how much of the PCIe link on this box.
// Background: the ROI staging copy of tsd_detect_frames (stage_copy_kernel) moves ~439 MB of scattered 32-byte sectors per 1024 frames
// at ~36 GB/s while cudaMemcpyAsync (copy engine, large requests) reaches ~55 GB/s on the same box.  32-byte read completions carry
// ~24 bytes of TLP overhead: 63 GB/s x 32 / 56 = 36 GB/s -- so the question is which access shapes make the memory system issue
// larger PCIe reads.  Every pattern copies runs of `run` bytes out of every `pitch` bytes of a 1 GiB pinned buffer into device memory.
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o build/pcie_patterns tools/pcie_patterns.cu && build/pcie_patterns
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdint>
#include <cstdlib>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); exit(1); } } while (0)

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// (a) lanes of a warp read consecutive 16-byte pieces of a run (the staging copy's shape); hint: 0 none, 1 L2::64B, 2 L2::128B, 3 L2::256B
template <int HINT, int UNROLL>
__global__ void __launch_bounds__(256) k_ld16(const uint8_t* __restrict__ src, uint8_t* __restrict__ dst, int64_t nruns, int run, int pitch) {
    const int lane = threadIdx.x & 31;
    const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5, nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    const int per = run / 16 < 32 ? run / 16 : 32;           // 16-byte pieces of a run one warp instruction covers
    const int sub = run / 16 / per;                          // runs longer than 512 bytes: `sub` units of 512 bytes
    const int rpw = 32 / per;                                // units per warp instruction
    const int64_t nunits = nruns * sub;
    for (int64_t r0 = warp * rpw * UNROLL; r0 < nunits; r0 += nwarps * rpw * UNROLL) {
        uint4 v[UNROLL];
        int64_t o[UNROLL];
#pragma unroll
        for (int u = 0; u < UNROLL; u++) {
            const int64_t un = r0 + (int64_t)u * rpw + lane / per;
            o[u] = (un < nunits && lane < rpw * per) ? (un / sub) * pitch + (un % sub) * 512 + (lane % per) * 16 : -1;
            if (o[u] >= 0) {
                const uint4* p = reinterpret_cast<const uint4*>(src + o[u]);
                if (HINT == 0) asm volatile("ld.global.cs.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v[u].x), "=r"(v[u].y), "=r"(v[u].z), "=r"(v[u].w) : "l"(p));
                if (HINT == 1) asm volatile("ld.global.L2::64B.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v[u].x), "=r"(v[u].y), "=r"(v[u].z), "=r"(v[u].w) : "l"(p));
                if (HINT == 2) asm volatile("ld.global.L2::128B.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v[u].x), "=r"(v[u].y), "=r"(v[u].z), "=r"(v[u].w) : "l"(p));
                if (HINT == 3) asm volatile("ld.global.L2::256B.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v[u].x), "=r"(v[u].y), "=r"(v[u].z), "=r"(v[u].w) : "l"(p));
            }
        }
#pragma unroll
        for (int u = 0; u < UNROLL; u++) if (o[u] >= 0) *reinterpret_cast<uint4*>(dst + o[u]) = v[u];
    }
}

// (b) one elected lane per warp: cp.async.bulk global -> shared (TMA 1-D bulk copy, run bytes) behind an mbarrier, then
// cp.async.bulk shared -> global into the device mirror.  DEPTH runs in flight per warp.
template <int DEPTH>
__global__ void __launch_bounds__(128) k_bulk(const uint8_t* __restrict__ src, uint8_t* __restrict__ dst, int64_t nruns, int run, int pitch) {
    extern __shared__ __align__(128) uint8_t smem[];
    __shared__ __align__(8) uint64_t bars[4][DEPTH];
    const int lane = threadIdx.x & 31, wl = threadIdx.x >> 5;
    const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5, nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    uint8_t* buf = smem + (size_t)wl * DEPTH * run;
    if (lane == 0) {
        for (int d = 0; d < DEPTH; d++) asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(&bars[wl][d])), "r"(1) : "memory");
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncwarp();
    if (lane != 0) return;
    uint32_t phase = 0;
    for (int64_t r0 = warp * DEPTH; r0 < nruns; r0 += nwarps * DEPTH) {
        int nd = 0;
        for (int d = 0; d < DEPTH && r0 + d < nruns; d++, nd++) {
            asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(&bars[wl][d])), "r"(run) : "memory");
            asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                         ::"r"(smem_u32(buf + (size_t)d * run)), "l"(src + (r0 + d) * pitch), "r"(run), "r"(smem_u32(&bars[wl][d])) : "memory");
        }
        for (int d = 0; d < nd; d++) {
            asm volatile("{\n.reg .pred p;\nW_%=:\nmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n@p bra D_%=;\nbra W_%=;\nD_%=:\n}\n"
                         ::"r"(smem_u32(&bars[wl][d])), "r"(phase) : "memory");
            asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst + (r0 + d) * pitch), "r"(smem_u32(buf + (size_t)d * run)), "r"(run) : "memory");
        }
        asm volatile("cp.async.bulk.commit_group;" ::: "memory");
        asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");      // the buffers are free again
        phase ^= 1;
    }
    asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
}

template <typename F>
static void timeit(const char* name, int run, int pitch, int64_t nruns, F launch) {
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    launch();
    CK(cudaDeviceSynchronize());
    CK(cudaEventRecord(e0));
    for (int i = 0; i < 3; i++) launch();
    CK(cudaEventRecord(e1));
    CK(cudaDeviceSynchronize());
    float ms = 0;
    CK(cudaEventElapsedTime(&ms, e0, e1));
    ms /= 3;
    printf("%-34s run %5d B pitch %5d B : %7.2f MB in %7.3f ms = %6.2f GB/s\n", name, run, pitch, nruns * (double)run / 1e6, ms, nruns * (double)run / ms / 1e6);
    fflush(stdout);
}

int main() {
    const size_t N = (size_t)1 << 30;
    uint8_t *h = nullptr, *d = nullptr;
    CK(cudaHostAlloc((void**)&h, N, cudaHostAllocDefault));
    for (size_t i = 0; i < N; i += 4096) h[i] = (uint8_t)i;
    CK(cudaMalloc((void**)&d, N));
    int sms = 0;
    CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0));
    {   // copy engine
        cudaEvent_t e0, e1;
        CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
        CK(cudaMemcpyAsync(d, h, N, cudaMemcpyHostToDevice));
        CK(cudaEventRecord(e0));
        CK(cudaMemcpyAsync(d, h, N, cudaMemcpyHostToDevice));
        CK(cudaEventRecord(e1));
        CK(cudaDeviceSynchronize());
        float ms = 0;
        CK(cudaEventElapsedTime(&ms, e0, e1));
        printf("cudaMemcpyAsync 1 GiB pinned H2D: %.2f GB/s\n", N / ms / 1e6);
    }
    const int grid = sms * 8;
    struct Shape { int run, pitch; } shapes[] = {{32, 256}, {64, 256}, {128, 256}, {160, 1024}, {256, 512}, {512, 1024}, {1024, 2048}, {4096, 4096}};
    for (auto s : shapes) {
        const int64_t nr = (int64_t)N / s.pitch < (int64_t)(256 << 20) / s.run ? (int64_t)N / s.pitch : (int64_t)(256 << 20) / s.run;
        timeit("ld 16B/lane .cs, 4 in flight", s.run, s.pitch, nr, [&] { k_ld16<0, 4><<<grid, 256>>>(h, d, nr, s.run, s.pitch); });
        timeit("ld 16B/lane .cs, 8 in flight", s.run, s.pitch, nr, [&] { k_ld16<0, 8><<<grid, 256>>>(h, d, nr, s.run, s.pitch); });
        if (s.run >= 64) timeit("ld 16B/lane L2::64B", s.run, s.pitch, nr, [&] { k_ld16<1, 4><<<grid, 256>>>(h, d, nr, s.run, s.pitch); });
        if (s.run >= 128) timeit("ld 16B/lane L2::128B", s.run, s.pitch, nr, [&] { k_ld16<2, 4><<<grid, 256>>>(h, d, nr, s.run, s.pitch); });
        if (s.run >= 256) timeit("ld 16B/lane L2::256B", s.run, s.pitch, nr, [&] { k_ld16<3, 4><<<grid, 256>>>(h, d, nr, s.run, s.pitch); });
        if (s.run * 4 * 4 <= 96 * 1024) {
            CK(cudaFuncSetAttribute(k_bulk<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, 4 * 4 * s.run));
            timeit("cp.async.bulk g->s->g, 4 per warp", s.run, s.pitch, nr, [&] { k_bulk<4><<<sms * 4, 128, 4 * 4 * s.run>>>(h, d, nr, s.run, s.pitch); });
        }
        if (s.run * 4 * 16 <= 96 * 1024) {
            CK(cudaFuncSetAttribute(k_bulk<16>, cudaFuncAttributeMaxDynamicSharedMemorySize, 4 * 16 * s.run));
            timeit("cp.async.bulk g->s->g, 16 per warp", s.run, s.pitch, nr, [&] { k_bulk<16><<<sms * 4, 128, 4 * 16 * s.run>>>(h, d, nr, s.run, s.pitch); });
        }
        CK(cudaGetLastError());
    }
    return 0;
}
