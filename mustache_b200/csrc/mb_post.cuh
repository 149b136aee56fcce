// Block post-processing on the device: what mustache() does right after the scale-space loop (mustache.py:774-811), so
// that only the surviving candidates leave the GPU instead of every found record.
//   1. Benjamini-Hochberg per block (mustache.py:778, statsmodels fdrcorrection method 'indep'): the block's found p-values
//      sorted ascending (mb_sort.cuh), q_(i) = min_{j >= i} p_(j) / ((j + 1) / m) clipped to 1 -- the same IEEE divisions the
//      host formula performs, a reverse running minimum, scattered back to the records.
//   2. Selection o < pt (mustache.py:791-798).
//   3. Sparsity filter (mustache.py:800-811) with numpy's slice semantics: a window that starts at a negative index is
//      empty, a window running past the tile is clipped but still divided by (2s+1)^2; `nonsparse = x != 0`.
//   4. For every selected pixel the 3 x 3 neighbourhood of the dense `o` / `so` matrices (1 off the mask, 2 / 1 on the mask
//      but never updated, q / sigma where found; mustache.py:789-795), which is all the clustering step (:830-848) reads,
//      and the pixel's value in the 2-filled tile for the enrichment filter (:822-828).
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

#include "mb_sort.cuh"
#include "mb_normalize.cuh"

// keys = bit patterns of the p-values (positive doubles order like their bits), payload = record slot in the block
__global__ void __launch_bounds__(256)
bh_keys_kernel(const unsigned long long* __restrict__ rec_count, long long rec_cap, const double* __restrict__ rec_p,
               unsigned long long* __restrict__ keys, unsigned* __restrict__ vals) {
    const int b = blockIdx.y;
    unsigned long long m = rec_count[b];
    if (m > (unsigned long long)rec_cap) m = rec_cap;
    for (unsigned long long r = blockIdx.x * 256ULL + threadIdx.x; r < m; r += (unsigned long long)gridDim.x * 256ULL) {
        const size_t o = (size_t)b * rec_cap + r;
        keys[o] = (unsigned long long)__double_as_longlong(rec_p[o]);
        vals[o] = (unsigned)r;
    }
}

__device__ __forceinline__ double bh_raw(unsigned long long key, long long i, long long m) {
    // ps / ecdf with ecdf = arange(1, m + 1) / float(m)
    return __ddiv_rn(__longlong_as_double((long long)key), __ddiv_rn((double)(i + 1), (double)m));
}

// minimum of raw over every tile of RS_TILE sorted p-values: tmin[seg * ntiles + tile]
__global__ void __launch_bounds__(RS_THREADS)
bh_tilemin_kernel(const unsigned long long* __restrict__ keys, const unsigned long long* __restrict__ rec_count, long long rec_cap,
                  int ntiles, double* __restrict__ tmin) {
    __shared__ double sh[RS_WARPS];
    const int b = blockIdx.y, tile = blockIdx.x;
    long long m = (long long)min(rec_count[b], (unsigned long long)rec_cap);
    const double kInf = __longlong_as_double(0x7ff0000000000000LL);
    double mn = kInf;
    const long long t0 = (long long)tile * RS_TILE;
    for (int it = 0; it < RS_ITEMS; ++it) {
        const long long i = t0 + it * RS_THREADS + threadIdx.x;
        if (i < m) mn = fmin(mn, bh_raw(keys[(size_t)b * rec_cap + i], i, m));
    }
    for (int o = 16; o > 0; o >>= 1) mn = fmin(mn, __shfl_xor_sync(0xffffffffu, mn, o));
    if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = mn;
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int w = 1; w < RS_WARPS; ++w) mn = fmin(mn, sh[w]);
        tmin[(size_t)b * ntiles + tile] = mn;
    }
}

// tmin[tile] <- min over the tiles AFTER it (exclusive suffix minimum).  One thread per block: ntiles is small.
__global__ void __launch_bounds__(64)
bh_suffix_kernel(int nblk, int ntiles, double* __restrict__ tmin) {
    const int b = blockIdx.x * 64 + threadIdx.x;
    if (b >= nblk) return;
    double run = __longlong_as_double(0x7ff0000000000000LL);
    for (int t = ntiles - 1; t >= 0; --t) {
        const double v = tmin[(size_t)b * ntiles + t];
        tmin[(size_t)b * ntiles + t] = run;
        run = fmin(run, v);
    }
}

// q of every record: reverse running minimum inside the tile joined with the suffix minimum of the later tiles
__global__ void __launch_bounds__(RS_THREADS)
bh_q_kernel(const unsigned long long* __restrict__ keys, const unsigned* __restrict__ vals,
            const unsigned long long* __restrict__ rec_count, long long rec_cap, int ntiles, const double* __restrict__ tsuf,
            double* __restrict__ rec_q) {
    __shared__ double th[RS_THREADS];
    const int b = blockIdx.y, tile = blockIdx.x;
    const long long m = (long long)min(rec_count[b], (unsigned long long)rec_cap);
    const long long t0 = (long long)tile * RS_TILE;
    if (t0 >= m) return;
    const double kInf = __longlong_as_double(0x7ff0000000000000LL);
    // thread t owns the RS_ITEMS consecutive ranks t0 + t * RS_ITEMS ..
    const long long i0 = t0 + (long long)threadIdx.x * RS_ITEMS;
    double raw[RS_ITEMS];
    double mn = kInf;
#pragma unroll
    for (int k = RS_ITEMS - 1; k >= 0; --k) {
        const long long i = i0 + k;
        raw[k] = i < m ? bh_raw(keys[(size_t)b * rec_cap + i], i, m) : kInf;
        mn = fmin(mn, raw[k]);
    }
    th[threadIdx.x] = mn;
    __syncthreads();
    // exclusive suffix minimum over the threads (Hillis-Steele, reversed)
    double after = kInf;
    for (int o = 1; o < RS_THREADS; o <<= 1) {
        const double other = (threadIdx.x + o < RS_THREADS) ? th[threadIdx.x + o] : kInf;
        __syncthreads();
        th[threadIdx.x] = fmin(th[threadIdx.x], other);
        __syncthreads();
    }
    after = (threadIdx.x + 1 < RS_THREADS) ? th[threadIdx.x + 1] : kInf;
    double run = fmin(after, tsuf[(size_t)b * ntiles + tile]);
#pragma unroll
    for (int k = RS_ITEMS - 1; k >= 0; --k) {
        const long long i = i0 + k;
        if (i < m) {
            run = fmin(run, raw[k]);
            rec_q[(size_t)b * rec_cap + vals[(size_t)b * rec_cap + i]] = run > 1.0 ? 1.0 : run;
        }
    }
}

// slot map: band-layout int32 tile holding, for every found pixel, its record slot (-1 elsewhere; memset 0xFF before)
__global__ void __launch_bounds__(256)
post_slotmap_kernel(const unsigned long long* __restrict__ rec_count, long long rec_cap, const int* __restrict__ rec_row,
                    const int* __restrict__ rec_col, int n, int wc, int* __restrict__ slot) {
    const int b = blockIdx.y;
    unsigned long long m = rec_count[b];
    if (m > (unsigned long long)rec_cap) m = rec_cap;
    for (unsigned long long r = blockIdx.x * 256ULL + threadIdx.x; r < m; r += (unsigned long long)gridDim.x * 256ULL) {
        const size_t o = (size_t)b * rec_cap + r;
        const int i = rec_row[o], j = rec_col[o];
        slot[((size_t)b * n + i) * wc + (j - i - 4)] = (int)r;
    }
}

struct PostOut {
    int* block;
    int* row;
    int* col;
    int* flags;            // bit 0: passes the sparsity filter (nonsparse)
    double* q;
    double* sigma;
    double* cval;          // value of the pixel in the 2-filled tile (mustache.py:703-706)
    double* o9;            // [cand][9] dense `o` over the 3 x 3 neighbourhood, row-major
    double* so9;           // [cand][9] dense `so`
    double* pair9;         // differential runs only (else nullptr): [cand][9] dense `pair` of the candidate's own map,
    double* vself9;        //   `v` of its own map and `v` of the other map of the block pair (diff_mustache.py:445-453):
    double* vother9;       //   1 off that map's mask, pPair / vAll where found, 2 / 0 on the mask but never updated
    unsigned long long* count;     // candidates emitted (may exceed cap: only cap are written)
    long long cap;
};

// number of mask pixels in rows [r0, r1) x cols [c0, c1) of the tile, summed over the lanes of a warp
__device__ __forceinline__ int post_window_count(const double* __restrict__ rawb, int n, int wc, int dhi, int r0, int r1, int c0,
                                                 int c1, int lane) {
    const int w = c1 - c0;
    int cnt = 0;
    if (w > 0 && r1 > r0) {
        const int total = (r1 - r0) * w;
        for (int t = lane; t < total; t += 32) {
            const int r = r0 + t / w, c = c0 + t % w, d = c - r;
            if (d >= 4 && d <= dhi && rawb[(size_t)r * wc + (d - 4)] != 0.0) ++cnt;
        }
    }
    for (int o = 16; o > 0; o >>= 1) cnt += __shfl_xor_sync(0xffffffffu, cnt, o);
    return cnt;
}

// selection o < pt (mustache.py:791), one thread per record: the selected records' (block, slot) go to the candidate list
__global__ void __launch_bounds__(256)
post_select_kernel(const unsigned long long* __restrict__ rec_count, long long rec_cap, const double* __restrict__ rec_q, double pt,
                   PostOut out, int* __restrict__ cand_slot) {
    const int b = blockIdx.y;
    unsigned long long m = rec_count[b];
    if (m > (unsigned long long)rec_cap) m = rec_cap;
    for (unsigned long long r = blockIdx.x * 256ULL + threadIdx.x; r < m; r += (unsigned long long)gridDim.x * 256ULL) {
        if (rec_q[(size_t)b * rec_cap + r] < pt) {
            const unsigned long long pos = atomicAdd(out.count, 1ULL);
            if (pos < (unsigned long long)out.cap) {
                out.block[pos] = b;
                cand_slot[pos] = (int)r;
            }
        }
    }
}

// one warp per selected record: sparsity filter, neighbourhood patches
__global__ void __launch_bounds__(256)
post_candidates_kernel(long long rec_cap, const int* __restrict__ rec_row, const int* __restrict__ rec_col,
                       const double* __restrict__ rec_q, const double* __restrict__ rec_sigma, const double* __restrict__ rec_v,
                       const double* __restrict__ rec_pair, const double* __restrict__ raw, const int* __restrict__ slot,
                       const int* __restrict__ cand_slot, int n, int wc, int dhi, int dpx, double st, PostOut out) {
    const int lane = threadIdx.x & 31;
    unsigned long long total = *out.count;
    if (total > (unsigned long long)out.cap) total = out.cap;
    for (unsigned long long pos = blockIdx.x * 8ULL + (threadIdx.x >> 5); pos < total; pos += (unsigned long long)gridDim.x * 8ULL) {
        const int b = out.block[pos];
        const size_t o = (size_t)b * rec_cap + cand_slot[pos];
        const double* rawb = raw + (size_t)b * n * wc;
        const int* slotb = slot + (size_t)b * n * wc;
        const double q = rec_q[o];
        const int x = rec_row[o], y = rec_col[o];
        const double sg = rec_sigma[o];
        // sparsity filter (mustache.py:800-811)
        bool keep = x != 0;
        {
            const int s = (int)ceil(sg);
            int cnt1 = 0, cnt2 = 0;
            if (x - s >= 0 && y - s >= 0) cnt1 = post_window_count(rawb, n, wc, dhi, x - s, min(x + s + 1, n), y - s, min(y + s + 1, n), lane);
            const int s2 = 2 * s;
            if (x - s2 >= 0 && y - s2 >= 0) cnt2 = post_window_count(rawb, n, wc, dhi, x - s2, min(x + s2 + 1, n), y - s2, min(y + s2 + 1, n), lane);
            const double c1 = (double)cnt1 / (double)((2 * s + 1) * (2 * s + 1));
            const double c2 = (double)cnt2 / (double)((2 * s2 + 1) * (2 * s2 + 1));
            if (c1 < st || c2 < 0.6) keep = false;
        }
        if (lane == 0) {
            out.row[pos] = x;
            out.col[pos] = y;
            out.flags[pos] = keep ? 1 : 0;
            out.q[pos] = q;
            out.sigma[pos] = sg;
            const int d = y - x;
            out.cval[pos] = (d <= 4 || d >= dpx + 1) ? 2.0 : rawb[(size_t)x * wc + (d - 4)];
        }
        if (lane < 9) {
            const int rr = x + lane / 3 - 1, cc = y + lane % 3 - 1, d = cc - rr;
            double ov = 1.0, sv = 1.0;                            // off the mask (or off the tile): np.ones_like(c)
            if (rr >= 0 && rr < n && cc >= 0 && cc < n && d >= 4 && d <= dhi && rawb[(size_t)rr * wc + (d - 4)] != 0.0) {
                const int sl = slotb[(size_t)rr * wc + (d - 4)];
                if (sl >= 0) {
                    ov = rec_q[(size_t)b * rec_cap + sl];
                    sv = rec_sigma[(size_t)b * rec_cap + sl];
                } else {
                    ov = 2.0;                                     // pAll stays 2, Scales stays 1 (mustache.py:708-709)
                }
            }
            out.o9[pos * 9 + lane] = ov;
            out.so9[pos * 9 + lane] = sv;
            if (out.pair9 != nullptr) {                           // blocks 2k / 2k+1 are the two maps of pair k
                const int bo = b ^ 1;
                const double* rawo = raw + (size_t)bo * n * wc;
                const int* sloto = slot + (size_t)bo * n * wc;
                double pv = 1.0, vs = 1.0, vo = 1.0;              // np.ones_like off the masks
                const bool inside = rr >= 0 && rr < n && cc >= 0 && cc < n && d >= 4 && d <= dhi;
                if (inside && rawb[(size_t)rr * wc + (d - 4)] != 0.0) {
                    const int sl = slotb[(size_t)rr * wc + (d - 4)];
                    pv = sl >= 0 ? rec_pair[(size_t)b * rec_cap + sl] : 2.0;          // pPair initialised to 2 (:290)
                    vs = sl >= 0 ? rec_v[(size_t)b * rec_cap + sl] : 0.0;             // vAll initialised to 0 (:292)
                }
                if (inside && rawo[(size_t)rr * wc + (d - 4)] != 0.0) {
                    const int sl = sloto[(size_t)rr * wc + (d - 4)];
                    vo = sl >= 0 ? rec_v[(size_t)bo * rec_cap + sl] : 0.0;
                }
                out.pair9[pos * 9 + lane] = pv;
                out.vself9[pos * 9 + lane] = vs;
                out.vother9[pos * 9 + lane] = vo;
            }
        }
    }
}

// ---------------------------------------------------------------------------------------------------------------
// Enrichment filter (mustache.py:816-828): keep a candidate iff c[x, y] > 2 * mean(non-zero entries of the (y - x)-th
// diagonal of the 2-filled tile).  Diagonals <= 4 and >= dpx + 1 are constant 2 (mean exactly 2); for the others the
// non-zero entries are the mask pixels of that diagonal in row order, and np.mean is numpy's pairwise sum over that
// compacted sequence divided by its length -- reproduced bit for bit: a warp compacts the diagonal into a scratch line
// (ballot + prefix, row order), then eight lanes run numpy's pairwise summation (mb_normalize.cuh) over it.
// ---------------------------------------------------------------------------------------------------------------
// marks the (block, diagonal) lines the sparsity-passing candidates need: need[b * wc + (d - 4)] = 1
__global__ void __launch_bounds__(256)
enrich_mark_kernel(const unsigned long long* __restrict__ count, long long cap, const int* __restrict__ cblock,
                   const int* __restrict__ crow, const int* __restrict__ ccol, const int* __restrict__ cflags, int wc, int dpx,
                   int* __restrict__ need) {
    unsigned long long total = *count;
    if (total > (unsigned long long)cap) total = cap;
    for (unsigned long long c = blockIdx.x * 256ULL + threadIdx.x; c < total; c += (unsigned long long)gridDim.x * 256ULL) {
        const int d = ccol[c] - crow[c];
        if ((cflags[c] & 1) && d > 4 && d < dpx + 1) need[(size_t)cblock[c] * wc + (d - 4)] = 1;
    }
}

// gives every marked line (need == 1) a slot in the scratch pool: need[line] becomes slot + 2, lines[slot] = line
__global__ void __launch_bounds__(256)
enrich_list_kernel(int nlines, int* __restrict__ need, int* __restrict__ lines, unsigned* __restrict__ nlist, unsigned max_slots) {
    for (int l = blockIdx.x * 256 + threadIdx.x; l < nlines; l += gridDim.x * 256) {
        if (need[l]) {
            const unsigned slot = atomicAdd(nlist, 1u);
            if (slot < max_slots) {
                lines[slot] = l;
                need[l] = (int)slot + 2;                     // >= 2: slot + 2
            } else {
                need[l] = -1;                                // no room in this round: handled by the next one
            }
        }
    }
}

// one warp per listed line: compaction of the diagonal's non-zero values in row order, then numpy's pairwise mean
__global__ void __launch_bounds__(256)
enrich_mean_kernel(const double* __restrict__ raw, int n, int wc, const int* __restrict__ lines, const unsigned* __restrict__ nlist,
                   unsigned max_slots, double* __restrict__ scratch, double* __restrict__ mean) {
    const int lane = threadIdx.x & 31;
    unsigned total = *nlist;
    if (total > max_slots) total = max_slots;
    for (unsigned slot = blockIdx.x * 8 + (threadIdx.x >> 5); slot < total; slot += gridDim.x * 8) {
        const int line = lines[slot], b = line / wc, kidx = line - b * wc, k = kidx + 4;
        const double* col = raw + (size_t)b * n * wc + kidx;
        double* out = scratch + (size_t)slot * n;
        const int rows = n - k;                                  // entries (i, i + k), i = 0 .. n - k - 1
        int m = 0;
        constexpr int INFLIGHT = 8;                              // strided (one sector per element) loads in flight per lane
        for (int i0 = 0; i0 < rows; i0 += 32 * INFLIGHT) {
            double v[INFLIGHT];
#pragma unroll
            for (int q = 0; q < INFLIGHT; ++q) {
                const int i = i0 + q * 32 + lane;
                v[q] = i < rows ? col[(size_t)i * wc] : 0.0;
            }
#pragma unroll
            for (int q = 0; q < INFLIGHT; ++q) {
                const unsigned nzb = __ballot_sync(0xffffffffu, v[q] != 0.0);
                if (v[q] != 0.0) out[m + __popc(nzb & ((1u << lane) - 1u))] = v[q];
                m += __popc(nzb);
            }
        }
        __syncwarp();
        double mu = 0.0;
        if (lane < 8) mu = __ddiv_rn(nz_pairwise8(out, (long long)m, NzIdentity(), lane, 0xffu), (double)m);   // 0 / 0 = NaN as np.mean
        if (lane == 0) mean[slot] = mu;
    }
}

// flags bit 1: the candidate passes the enrichment filter
__global__ void __launch_bounds__(256)
enrich_apply_kernel(const unsigned long long* __restrict__ count, long long cap, const int* __restrict__ cblock,
                    const int* __restrict__ crow, const int* __restrict__ ccol, const double* __restrict__ cval, int wc, int dpx,
                    const int* __restrict__ need, const double* __restrict__ mean, int* __restrict__ cflags) {
    unsigned long long total = *count;
    if (total > (unsigned long long)cap) total = cap;
    for (unsigned long long c = blockIdx.x * 256ULL + threadIdx.x; c < total; c += (unsigned long long)gridDim.x * 256ULL) {
        if (!(cflags[c] & 1) || (cflags[c] & 4)) continue;       // failed the sparsity filter, or already decided
        const int d = ccol[c] - crow[c];
        double mu = 2.0;                                         // diagonals <= 4 and >= dpx + 1 are constant 2
        if (d > 4 && d < dpx + 1) {
            const int s = need[(size_t)cblock[c] * wc + (d - 4)];
            if (s < 2) continue;                                 // its line is computed in a later round
            mu = mean[s - 2];
        }
        cflags[c] |= 4 | ((cval[c] > __dmul_rn(2.0, mu)) ? 2 : 0);
    }
}

// between rounds: lines served in this round are done, deferred ones become pending again
__global__ void __launch_bounds__(256)
enrich_next_round_kernel(int nlines, int* __restrict__ need) {
    for (int l = blockIdx.x * 256 + threadIdx.x; l < nlines; l += gridDim.x * 256) need[l] = need[l] == -1 ? 1 : 0;
}
