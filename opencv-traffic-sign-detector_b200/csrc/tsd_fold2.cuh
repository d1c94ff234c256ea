// tsd_fold2.cuh -- K5 (cleanDuplicatedDetections, DET:177-223) split into a throughput part and a latency part.
//
//  k5_pairs_kernel : ALL pairs (i < j) of a frame's windows -> class byte M[f][j][i] in {0 none, 1 delete, 2 merge}
//                    for the histogram pass (cv2.compareHist CORREL of the two min-max normalised H-S histograms,
//                    DET:200-202,575-586).  Embarrassingly parallel: one warp per item j (its dense histogram in
//                    shared memory), no block barriers; the sparse histograms of the earlier windows stream through L1.
//                    The dot product is first taken on the integer bin counts (exact), which bounds the f64 CORREL
//                    value to ~1e-7; only pairs closer than 1e-6 to a threshold take the exact f64 path, so every
//                    class equals the one the exact evaluation gives.
//  k5_fold2_kernel : the sequential fold itself (one CTA per frame).  Warp 0 walks the items in MSER order with
//                    warp-synchronous code (no block barriers on the common path): it looks the classes up in M,
//                    finds the first merge with a ballot, marks deletions, applies the pop-by-pixel-equality rule.
//                    Only a MERGE wakes the helper warps (named barriers): average the pixels, rebuild the merged
//                    item's histogram, re-classify it against the remaining survivors, and -- once the item is final --
//                    against every later item (column j > m of M), because M's precomputed entries describe the
//                    un-merged histogram.  Pass 2 (corner similarity, DET:209-213) is a table lookup per survivor.
#pragma once
#include "tsd_kernels.cuh"

namespace tsd {

constexpr int kPairWarps = 8;
constexpr int kClsUnsure = 3;

__device__ __forceinline__ int classify(double sim, double tol, double lo) {
    return sim > tol ? 1 : ((lo <= sim && sim <= tol) ? 2 : 0);
}

// cv2.compareHist(h1, h2, HISTCMP_CORREL) from the sparse dot product and the cached moments (symmetric in h1, h2)
__device__ __forceinline__ double correl_from(double s12, const WinMeta& x, const WinMeta& y) {
    const double scale = 1.0 / (double)kHistBins;
    double num = s12 - x.s1 * y.s1 * scale;
    double den2 = x.A * y.A;                                 // (s11 - s1*s1*scale) * (s22 - s2*s2*scale)
    return fabs(den2) > DBL_EPSILON ? num / sqrt(den2) : 1.0;
}

// Class from the EXACT integer dot product I = sum_b cnt_x[b]*cnt_y[b].  The f64 value the reference computes is
// s12 = sum fl32(cnt_x a_x) * fl32(cnt_y a_y) = a_x a_y I (1 + e), |e| <= 2^-23 + O(1e-16); the class is decided here
// when the approximation is further than 1e-6 (relative to sqrt(denom2)) from both thresholds, else kClsUnsure.
__device__ __forceinline__ int classify_from_int(int I, const WinMeta& x, const WinMeta& y, double tol, double lo) {
    const double scale = 1.0 / (double)kHistBins;
    const double den2 = x.A * y.A;
    if (!(fabs(den2) > DBL_EPSILON)) return classify(1.0, tol, lo);
    if (!(den2 > 0.0) || !(x.A > 0.0) || !(y.A > 0.0)) return kClsUnsure;
    const double num = (double)x.a * (double)y.a * (double)I - x.s1 * y.s1 * scale;
    const double r = x.rA * y.rA;                            // sqrt(denom2) to ~1e-16
    const double m = 2e-6 * r + 1e-300;
    const double hi_t = tol * r, lo_t = lo * r;
    if (num > hi_t + m) return 1;
    if (num < lo_t - m) return 0;
    if (num > lo_t + m && num < hi_t - m) return 2;
    return kClsUnsure;
}

// warp-wide exact s12 (f64) of a dense histogram (counts, owner's scale a_d) against a sparse one
template <typename DenseT>
__device__ __forceinline__ double exact_s12_warp(const DenseT* dense, float a_d, const uint32_t* eo, int nnz_o, float a_o) {
    const int lane = threadIdx.x & 31;
    double s12 = 0;
    for (int e = lane; e < nnz_o; e += 32) {
        const uint32_t v = eo[e];
        const float hd = (float)dense[v >> 16] * a_d;
        const float ho = (float)(v & 0xffffu) * a_o;
        s12 += (double)hd * (double)ho;
    }
    return warp_sum(s12);
}

template <typename DenseT>
__device__ __forceinline__ int int_dot_warp(const DenseT* dense, const uint32_t* eo, int nnz_o) {
    const int lane = threadIdx.x & 31;
    int acc = 0;
#pragma unroll 4
    for (int e = lane; e < nnz_o; e += 32) {
        const uint32_t v = eo[e];
        acc += (int)dense[v >> 16] * (int)(v & 0xffffu);
    }
    return warp_sum_i(acc);
}

// full class of one pair, warp-wide (used by the fold's merge jobs; plain loads: data may have been rewritten in-kernel)
template <typename DenseT>
__device__ __forceinline__ int pair_class_warp(const DenseT* dense, const WinMeta& md, const uint32_t* eo, const WinMeta& mo,
                                               double tol, double lo) {
    const int I = int_dot_warp(dense, eo, mo.nnz);
    int c = classify_from_int(I, md, mo, tol, lo);
    if (c == kClsUnsure) c = classify(correl_from(exact_s12_warp(dense, md.a, eo, mo.nnz, mo.a), md, mo), tol, lo);
    return c;
}

// Cauchy-Schwarz bound from the group energies: can the pair (x, y) reach the merge band at all?
// ub >= s12 (energies are inflated by 1e-5), so (ub - s1x s1y / N) / sqrt(Ax Ay) >= CORREL; below lo -> class 0.
__device__ __forceinline__ bool prunable(float ub, const WinMeta& x, double s1y, double Ay, double rAy, double lo) {
    const double den2 = x.A * Ay;
    if (!(den2 > DBL_EPSILON) || !(x.A > 0.0) || !(Ay > 0.0)) return false;
    const double num_ub = (double)ub * 1.00001 - x.s1 * s1y * (1.0 / (double)kHistBins);
    return num_ub < lo * (x.rA * rAy) * 0.99999 - 1e-300;
}

__global__ void __launch_bounds__(kPairWarps * 32) k5_pairs_kernel(const uint32_t* __restrict__ entries, const WinMeta* __restrict__ meta,
                                                                   const float* __restrict__ E_T, int64_t e_stride,
                                                                   const int32_t* __restrict__ offsets, int nframes, int npx, int R,
                                                                   int tiles_per_frame, double tol, double lo, uint8_t* __restrict__ M) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int f = blockIdx.x / tiles_per_frame, tile = blockIdx.x - f * tiles_per_frame;
    if (f >= nframes) return;
    const int base = offsets[f], n = offsets[f + 1] - base;
    const int wid = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int jl = tile * kPairWarps + wid;
    if (jl >= n || jl == 0 || n > R) return;                // warps are independent: no block barrier below
    uint16_t* dense = reinterpret_cast<uint16_t*>(smem_raw) + (size_t)wid * (kHistBins + 8);
    for (int b = lane; b < (kHistBins + 8) / 2; b += 32) reinterpret_cast<uint32_t*>(dense)[b] = 0;
    __syncwarp();
    const WinMeta mj = meta[base + jl];
    {
        const uint32_t* ej = entries + (int64_t)(base + jl) * npx;
        for (int e = lane; e < mj.nnz; e += 32) { const uint32_t v = __ldg(ej + e); dense[v >> 16] = (uint16_t)(v & 0xffffu); }
    }
    const float Ej = lane < kHistGroups ? __ldg(E_T + (int64_t)lane * e_stride + base + jl) : 0.f;
    __syncwarp();
    uint8_t* Mrow = M + (int64_t)f * R * R + (int64_t)jl * R;
    for (int i0 = 0; i0 < jl; i0 += 32) {
        const int il = i0 + lane;
        const bool valid = il < jl;
        const int wi = base + (valid ? il : 0);
        // lane-parallel bound for 32 earlier windows at once
        float ub = 0.f;
#pragma unroll
        for (int g = 0; g < kHistGroups; g++) ub += __shfl_sync(0xffffffffu, Ej, g) * __ldg(E_T + (int64_t)g * e_stride + wi);
        const WinMeta mi = meta[wi];
        int c = 0;
        bool need = false;
        if (valid) {
            const double den2 = mj.A * mi.A;
            if (!(fabs(den2) > DBL_EPSILON)) c = classify(1.0, tol, lo);          // compareHist's degenerate branch
            else need = !prunable(ub, mj, mi.s1, mi.A, mi.rA, lo);
        }
        unsigned todo = __ballot_sync(0xffffffffu, need);
        int Ik = 0;
        while (todo) {
            const int k = __ffs(todo) - 1;
            todo &= todo - 1;
            const int nnz_k = __shfl_sync(0xffffffffu, mi.nnz, k);
            const uint32_t* ek = entries + (int64_t)(base + i0 + k) * npx;
            int acc = 0;
#pragma unroll 4
            for (int e = lane; e < nnz_k; e += 32) {
                const uint32_t v = __ldg(ek + e);
                acc += (int)dense[v >> 16] * (int)(v & 0xffffu);
            }
            acc = warp_sum_i(acc);
            if (lane == k) Ik = acc;
        }
        if (need) c = classify_from_int(Ik, mj, mi, tol, lo);
        unsigned unsure = __ballot_sync(0xffffffffu, need && c == kClsUnsure);
        while (unsure) {                                    // rare: within 1e-6 of a threshold -> exact f64 evaluation
            const int k = __ffs(unsure) - 1;
            unsure &= unsure - 1;
            const int nnz_k = __shfl_sync(0xffffffffu, mi.nnz, k);
            const float a_k = __shfl_sync(0xffffffffu, mi.a, k);
            const double s12 = exact_s12_warp(dense, mj.a, entries + (int64_t)(base + i0 + k) * npx, nnz_k, a_k);
            if (lane == k) c = classify(correl_from(s12, mj, mi), tol, lo);
        }
        if (valid) Mrow[il] = (uint8_t)c;
    }
}

// ---- the fold ------------------------------------------------------------------------------------------------------
constexpr int kFold2Threads = 128;
enum { JOB_EXIT = 0, JOB_MERGE1 = 1, JOB_COLUMN = 2, JOB_MERGE2 = 3 };

struct Fold2Job {
    int type, slot, other, jl, from, to, n, base, clear_old;
};

template <int RMAX>
struct Fold2Smem {
    HistSmem h;
    int4 scoords[RMAX];          // coords of the survivors, by list position
    int4 incoords[RMAX];         // pass 2: coords of the input items, by input position
    uint32_t shash[RMAX];        // low 32 bits of the survivors' pixel hash, by list position
    uint16_t list[RMAX];
    uint16_t list_in[RMAX];
    uint8_t cls[RMAX];
    uint8_t mark[RMAX];
    uint8_t row[2][RMAX];        // pass 1: prefetched rows of M (current / next item)
    Fold2Job job;
};

__device__ __forceinline__ void named_bar(int id, int nthreads) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory"); }

__device__ __forceinline__ unsigned long long pixel_hash_block(const uint8_t* px, int npx, double* red_scratch) {
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5, nw = blockDim.x >> 5;
    unsigned long long hsh = 0;
    for (int p = tid; p < npx; p += blockDim.x)
        hsh += mix64(((unsigned long long)p << 24) | (unsigned long long)(px[3 * p] | (px[3 * p + 1] << 8) | (px[3 * p + 2] << 16)));
    unsigned hl = (unsigned)hsh, hh = (unsigned)(hsh >> 32);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        unsigned long long other = ((unsigned long long)__shfl_xor_sync(0xffffffffu, hh, o) << 32) | __shfl_xor_sync(0xffffffffu, hl, o);
        hsh += other; hl = (unsigned)hsh; hh = (unsigned)(hsh >> 32);
    }
    unsigned long long* ru = reinterpret_cast<unsigned long long*>(red_scratch);
    __syncthreads();
    if (lane == 0) ru[wid] = hsh;
    __syncthreads();
    unsigned long long t = 0;
    for (int i = 0; i < nw; i++) t += ru[i];
    __syncthreads();
    return t;
}

// executed by ALL threads of the CTA between the two named barriers
template <int RMAX>
__device__ void fold2_run_job(Fold2Smem<RMAX>& sm, const FoldParams& P, int R, uint8_t* Mframe) {
    const Fold2Job jb = sm.job;
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5, nw = kFold2Threads >> 5;
    const int nbytes = P.npx * 3;
    uint8_t* ipx = P.windows + (int64_t)jb.slot * nbytes;
    uint32_t* ient = P.entries + (int64_t)jb.slot * P.npx;
    if (jb.type == JOB_MERGE1 || jb.type == JOB_MERGE2) {
        // cv2.addWeighted(item, .5, other, .5, 0): round-half-even of (a+b)/2 (DET:219)
        const uint8_t* kpx = P.windows + (int64_t)jb.other * nbytes;
        if (jb.type == JOB_MERGE1 && jb.clear_old) clear_dense_block(sm.h.dense, ient, P.meta[jb.slot].nnz);
        for (int i = tid; i < nbytes; i += kFold2Threads) {
            int s = ipx[i] + kpx[i];
            ipx[i] = (uint8_t)((s >> 1) + ((s & 1) & ((s >> 1) & 1)));
        }
        __syncthreads();
    }
    if (jb.type == JOB_MERGE1) {
        build_hist_block(ipx, P.npx, sm.h.sdiv, sm.h.hdiv, sm.h.hbin, sm.h.sbin, sm.h.dense, sm.h.red, ient, P.meta + jb.slot);
        // re-classify the updated item against the survivors after the merge position
        const WinMeta mi = P.meta[jb.slot];
        for (int p = jb.from + wid; p < jb.to; p += nw) {
            const int sj = jb.base + sm.list[p];
            const int c = pair_class_warp(sm.h.dense, mi, P.entries + (int64_t)sj * P.npx, P.meta[sj], P.hist_tol, P.hist_lo);
            if (lane == 0) sm.cls[p] = (uint8_t)c;
        }
    } else if (jb.type == JOB_COLUMN) {
        // the item is final: its class against every LATER item j2 (those rows of M described the un-merged histogram)
        const WinMeta ms = P.meta[jb.slot];
        for (int j2 = jb.jl + 1 + wid; j2 < jb.n; j2 += nw) {
            const int sj = jb.base + j2;
            const int c = pair_class_warp(sm.h.dense, ms, P.entries + (int64_t)sj * P.npx, P.meta[sj], P.hist_tol, P.hist_lo);
            if (lane == 0) Mframe[(int64_t)j2 * R + jb.jl] = (uint8_t)c;
        }
        __syncthreads();
        clear_dense_block(sm.h.dense, ient, ms.nnz);
    } else if (jb.type == JOB_MERGE2) {
        // keep the pixel hash current for the pop-by-equality rule
        const unsigned long long t = pixel_hash_block(ipx, P.npx, sm.h.red);
        if (tid == 0) P.meta[jb.slot].hash = t;
    }
}

// warp-level: pop-by-pixel-equality + order-preserving compaction of the list (DET:183-185,471-477).  Returns new S.
template <int RMAX>
__device__ int fold2_apply_deletions(Fold2Smem<RMAX>& sm, const FoldParams& P, int base, int S, int nbytes) {
    const int lane = threadIdx.x & 31;
    __syncwarp();
    for (int b0 = 0; b0 < S; b0 += 32) {
        unsigned marks = __ballot_sync(0xffffffffu, b0 + lane < S && sm.mark[b0 + lane] == 1);
        while (marks) {                                     // marks in increasing list position
            const int p = b0 + __ffs(marks) - 1;
            marks &= marks - 1;
            const uint32_t h32 = sm.shash[p];
            int victim = p;
            for (int q0 = 0; q0 < p && victim == p; q0 += 32) {
                const int q = q0 + lane;
                unsigned m = __ballot_sync(0xffffffffu, q < p && sm.mark[q] != 2 && sm.shash[q] == h32);
                while (m) {                                 // verify candidates in position order (hash collisions are possible)
                    const int qq = q0 + __ffs(m) - 1;
                    const uint8_t* A = P.windows + (int64_t)(base + sm.list[qq]) * nbytes;
                    const uint8_t* B = P.windows + (int64_t)(base + sm.list[p]) * nbytes;
                    bool eq = true;
                    for (int i = lane; i < nbytes; i += 32) eq &= A[i] == B[i];
                    if (__all_sync(0xffffffffu, eq)) { victim = qq; break; }
                    m &= m - 1;
                }
            }
            __syncwarp();
            if (lane == 0) { sm.mark[victim] = 2; if (victim != p) sm.mark[p] = 0; }
            __syncwarp();
        }
    }
    int outp = 0;
    for (int b = 0; b < S; b += 32) {
        const int p = b + lane;
        const bool live = p < S && sm.mark[p] != 2;
        const uint16_t v = p < S ? sm.list[p] : 0;
        const uint32_t hs = p < S ? sm.shash[p] : 0;
        const int4 c = p < S ? sm.scoords[p] : make_int4(0, 0, 0, 0);
        const unsigned m = __ballot_sync(0xffffffffu, live);
        __syncwarp();
        if (live) { const int o = outp + __popc(m & ((1u << lane) - 1)); sm.list[o] = v; sm.scoords[o] = c; sm.shash[o] = hs; }
        outp += __popc(m);
        __syncwarp();
    }
    for (int p = lane; p < S; p += 32) sm.mark[p] = 0;
    __syncwarp();
    return outp;
}

template <int RMAX>
__global__ void __launch_bounds__(kFold2Threads, RMAX <= 256 ? 8 : 1) k5_fold2_kernel(FoldParams P, int nframes, uint8_t* M, int R, int sim_cut) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    Fold2Smem<RMAX>& sm = *reinterpret_cast<Fold2Smem<RMAX>*>(smem_raw);
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    const int nbytes = P.npx * 3;
    load_tables_block(sm.h, P.tab);
    __syncthreads();
    for (int f = blockIdx.x; f < nframes; f += gridDim.x) {
        const int base = P.offsets[f], n = P.offsets[f + 1] - base;
        uint8_t* Mframe = M + (int64_t)f * R * R;
        if (wid != 0) {
            // helpers: sleep on barrier 1 until the controller posts a job
            while (true) {
                named_bar(1, kFold2Threads);
                if (sm.job.type == JOB_EXIT) break;
                fold2_run_job<RMAX>(sm, P, R, Mframe);
                named_bar(2, kFold2Threads);
            }
            __syncthreads();                                // frame done (matches the controller's)
            continue;
        }
        // ---------------- controller warp ----------------
        for (int p = lane; p < RMAX; p += 32) sm.mark[p] = 0;
        __syncwarp();
        int S = 0;
        if (n <= RMAX) {
            int nin = n;
            for (int pass = 0; pass < 2; pass++) {
                const bool by_coords = pass == 1;
                if (by_coords ? !P.do_coords : !P.do_hist) continue;
                const bool from_list = by_coords && P.do_hist;
                if (from_list) {
                    nin = S;
                    for (int p = lane; p < S; p += 32) { sm.list_in[p] = sm.list[p]; sm.incoords[p] = sm.scoords[p]; }
                } else if (by_coords) {
                    for (int p = lane; p < n; p += 32) sm.incoords[p] = P.coords[base + p];
                }
                __syncwarp();
                S = 0;
                const double tol = by_coords ? P.coord_tol : P.hist_tol, lo = by_coords ? P.coord_lo : P.hist_lo;
                if (!by_coords && nin > 1) {                // prefetch row 1 of M (row 0 has no earlier windows)
                    for (int b = lane; b < R && b < RMAX; b += 32) sm.row[1][b] = Mframe[(int64_t)1 * R + b];
                }
                for (int it = 0; it < nin; it++) {
                    const int jl = from_list ? sm.list_in[it] : it;
                    const int slot = base + jl;
                    int4 ic = by_coords ? sm.incoords[it] : P.coords[slot];   // issued early, consumed after the scan
                    uint32_t myhash = (uint32_t)P.meta[slot].hash;
                    bool dirty = false, any_del = false;
                    int start = 0;
                    const uint8_t* row = sm.row[it & 1];
                    const bool pf = !by_coords && it + 1 < nin;
                    __syncwarp();
                    if (pf) {                               // rows are R <= RMAX bytes; each lane carries up to RMAX/32 of them
#pragma unroll
                        for (int t = 0; t < RMAX / 128; t++) {
                            const int b = (t * 32 + lane) * 4;
                            if (b < R) reinterpret_cast<uint32_t*>(sm.row[(it + 1) & 1])[t * 32 + lane] =
                                *reinterpret_cast<const uint32_t*>(Mframe + (int64_t)(it + 1) * R + b);
                        }
                    }
                    while (start < S) {
                        int fm = S;
                        for (int p0 = start; p0 < S; p0 += 32) {
                            const int p = p0 + lane;
                            int c = 0;
                            if (p < S) {
                                if (by_coords) c = classify(coord_sim(ic, sm.scoords[p], P.simtab, sim_cut), tol, lo);
                                else c = dirty ? sm.cls[p] : row[sm.list[p]];
                            }
                            const unsigned mm = __ballot_sync(0xffffffffu, c == 2);
                            const int first = mm ? p0 + __ffs(mm) - 1 : S;
                            if (c == 1 && p < first) sm.mark[p] = 1;
                            any_del |= __any_sync(0xffffffffu, c == 1 && p < first) != 0;
                            if (mm) { fm = first; break; }
                        }
                        if (fm >= S) break;
                        // ---- merge with the survivor at list position fm (DET:217-221) ----
                        const int sk = base + sm.list[fm];
                        const int4 kc = sm.scoords[fm];
                        ic = make_int4((ic.x + kc.x) >> 1, (ic.y + kc.y) >> 1, (ic.z + kc.z) >> 1, (ic.w + kc.w) >> 1);   // Python // (coords >= 0)
                        if (lane == 0) {
                            sm.mark[fm] = 1;
                            sm.job.type = by_coords ? JOB_MERGE2 : JOB_MERGE1;
                            sm.job.slot = slot; sm.job.other = sk; sm.job.jl = jl;
                            sm.job.from = fm + 1; sm.job.to = S; sm.job.n = n; sm.job.base = base;
                            sm.job.clear_old = dirty ? 1 : 0;   // first merge of this item: the dense histogram is still empty
                        }
                        any_del = true;
                        __syncwarp();
                        named_bar(1, kFold2Threads);
                        fold2_run_job<RMAX>(sm, P, R, Mframe);
                        named_bar(2, kFold2Threads);
                        dirty = true;
                        start = fm + 1;
                    }
                    if (dirty && !by_coords) {
                        if (lane == 0) { sm.job.type = JOB_COLUMN; sm.job.slot = slot; sm.job.jl = jl; sm.job.n = n; sm.job.base = base; }
                        __syncwarp();
                        named_bar(1, kFold2Threads);
                        fold2_run_job<RMAX>(sm, P, R, Mframe);
                        named_bar(2, kFold2Threads);
                        if (pf && lane == 0) sm.row[(it + 1) & 1][jl] = Mframe[(int64_t)(it + 1) * R + jl];   // patch the prefetched row
                    }
                    if (dirty) { if (lane == 0) P.coords[slot] = ic; myhash = (uint32_t)P.meta[slot].hash; }
                    if (any_del) S = fold2_apply_deletions<RMAX>(sm, P, base, S, nbytes);
                    if (lane == 0) { sm.list[S] = (uint16_t)jl; sm.scoords[S] = ic; sm.shash[S] = myhash; }
                    S++;
                    __syncwarp();
                }
            }
            if (!(P.do_hist || P.do_coords)) {
                for (int p = lane; p < n; p += 32) sm.list[p] = (uint16_t)p;
                S = n;
                __syncwarp();
            }
            for (int p = lane; p < S; p += 32) P.list[base + p] = base + sm.list[p];
            if (lane == 0) P.out_count[f] = S;
        } else if (lane == 0) {
            P.out_count[f] = -1;                            // frame too large for this instantiation (host picks RMAX)
        }
        __syncwarp();
        if (lane == 0) sm.job.type = JOB_EXIT;
        __syncwarp();
        named_bar(1, kFold2Threads);
        __syncthreads();
    }
}

}  // namespace tsd
