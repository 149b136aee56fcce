// Scale-space kernels for sm_100a (B200).  FP64 end to end, bit-exact against scipy.ndimage.gaussian_filter.
//
// Replaces the body of mustache() between mustache.py:699 and mustache.py:772 (reference: ay-lab/mustache v1.3.3):
//   K_V  (kv_kernel)   axis-0 pass of every Gaussian of the chain            mustache.py:719,725,734,751 (first half of
//                      scipy gaussian_filter: correlate1d along axis 0, mode='reflect')
//   K_H  (kh_kernel)   axis-1 pass + DoG, written to HBM                     mustache.py:728,738,754 (second half of
//                      gaussian_filter + the subtraction)
//   K_S  (ks_kernel)   zero-padded 3x3 maxima + 5-clause extremum test + running best + per-level |L| min/sum
//                      mustache.py:740-743,757 (maximum_filter) :760-768 (test + state update)
//   reduce / finalise  expon.fit (loc = min, scale = mean - min) and 1 - expon.cdf for the winners only   mustache.py:755-756
//   K_HS (khs_kernel)  opt-in: K_H and K_S in one kernel, DoG levels in shared memory (mb200_set_fusion)
// Companion headers: mb_sort.cuh (device radix sort), mb_post.cuh (BH per block, o < pt, sparsity filter: mustache.py:774-811),
// mb_normalize.cuh (normalize_sparse, mustache.py:622-686), mb_parse.h (native text reader, mustache.py:254-263).
//
// Data layout in HBM ("band layout"): a block is an N x N tile but only diagonals d = j - i in [4, dhi] can hold data
// (reader keeps |j-i| <= dpx+1, mask needs j-i >= 4), so a tile is stored as raw[i][d-4], i in [0,N), wc = dhi-3 doubles per
// row; (i, j) with j >= N is never read.  The 2-fills (mustache.py:703-706) are applied on the fly when a tile is staged
// into shared memory.  Axis-0 results are stored as V[step][block][i][j-i-vlo] for diagonals [vlo, vlo+wv), vlo = 2-rmax.
//
// Arithmetic: scipy's NI_Correlate1D symmetric branch,  out = x[0]*w[0]; for j=-R..-1: out += (x[j]+x[-j])*w[j],
// with separate multiply and add (__dmul_rn/__dadd_rn are never contracted into FMA).  An FP64 instruction holds a
// sub-partition's dispatch for two cycles and every other instruction costs most of a cycle on top (tools/fp64_peak.cu),
// so the tap loops are written for the fewest instructions around the FP64 ones: register windows instead of reloads,
// running pointers with compile-time offsets, no remainder loops.
#pragma once
#include <cstdint>
#include <cuda.h>
#include <cuda_runtime.h>

#define MB_MAX_STEPS 64
#define MB_MAX_TAPS 1536
#define MB_FLAG_RESTART 1   // chain is cut before this Gaussian: no DoG is formed with the previous one
#define MB_FLAG_SCORE 2     // after forming this step's DoG, score the DoG before it (ring centre)
#define MB_FLAG_DIFFREF 4   // this step's DoG is L_2 of its octave (the only difference-stack DoG diff_mustache uses)

struct MbStep {
    int radius;
    int tap_off;     // taps[tap_off + j], j = 0..radius, weight at distance j
    int flags;
    int score_idx;   // 0-based index among scored steps (valid when MB_FLAG_SCORE); in the difference chain of
                     // diff_mustache: slot of this step's DoG in dout (valid when MB_FLAG_DIFFREF)
};

// kh_kernel stages the axis-0 tile of every step in a byte-granular ring of shared memory: small-radius steps have small
// boxes, so more of them are in flight.  Placement is computed on the host (mb_engine.cu: plan_kh_ring).
struct MbStage {
    int off;         // offset of the step's box in the ring, in doubles (multiple of 16 -> 128-byte aligned)
    int dep;         // latest earlier step whose box overlaps this one (-1: none): must be released before the copy
};

struct MbProgram {
    int n_steps;
    int n_scored;
    int rmax;
    int pad;
    MbStep st[MB_MAX_STEPS];
    MbStage stage[MB_MAX_STEPS];
    MbStage stage_f[MB_MAX_STEPS];  // same for the (smaller) ring of the fused khs_kernel
    int score_id[MB_MAX_STEPS];   // per scored index: octave*12 + i  (reference's scales[o][i])
    double taps[MB_MAX_TAPS];
};

// kv_kernel's view of the chain: the steps sorted by radius and cut into groups of up to KV_GMAX that share their pair
// sums (mb_engine.cu: plan_kv picks the cut that minimises the FP64 work).  Inside a group every step runs over the
// group's largest radius with its weights zero-padded: tapsT[tap_off + j*n + slot] = weight of `slot` at distance j, 0
// for j beyond the slot's own radius.  Adding (pair sum) * 0 before a step's first real tap leaves its accumulator
// unchanged, so the padded taps cost FP64 instructions but not exactness.
#define KV_GMAX 5
// Chains with large radii (the 4-octave ladder) are cut into groups of at most KV_GSMALL: 0.7 % more FP64 instructions than
// groups of 5, but 12 instead of 20 accumulators -- the kernel fits 64 registers and three CTAs per SM (the 61 KB tile is
// then the limit).  Measured on the 10k tile: axis-0 pass 5.67 -> 5.37 ms; on the 2-octave shapes the small groups are
// slower (0.85 -> 0.97 ms on 24 x 2000^2), so those keep groups of 5 at two CTAs per SM.
#define KV_GSMALL 3
#define KV_GSMALL_RMIN 24   // chains whose largest radius reaches this use the small groups
#define KV_MAX_GROUPS 32
#define KV_MAX_TAPS_T 2048
struct KvGroup {
    int n;              // steps in the group (1..KV_GMAX)
    int rmax;           // largest radius of the group
    int tap_off;
    int step[KV_GMAX];  // chain step per slot
};
struct KvPlan {
    int n_groups;
    int rmax;
    int gmax;           // largest group of this plan (KV_GSMALL or KV_GMAX): selects the kernel variant
    int pad;
    KvGroup grp[KV_MAX_GROUPS];
    double tapsT[KV_MAX_TAPS_T];
};

// TMA descriptors (cuTensorMapEncodeTiled, built by the host per batch geometry).  Both scratch arrays are addressed
// through a skewed 3-D view (x = column index, y = image row, z = step * nblk + block) whose row stride is one element
// shorter than the band row, so that a rectangular box of the view is a rectangular (i, j) tile of the image:
//   V: element (i, j) sits at  plane + i*wv + (j - i - vlo) = plane + i*(wv-1) + (j - vlo)   ->  x = j - vlo
//   L: element (i, j) sits at  plane + i*wl + (j - i - 2)   = plane + i*(wl-1) + (j - 2)     ->  x = j - 2
// Rows outside the image are out of bounds in y and arrive as zeros.  One descriptor per step for V (the box is as
// wide as that step's filter support), one for L.  The descriptors live in global memory (64-byte aligned) and are
// written by the host only.
struct MbTensorMaps {
    CUtensorMap v[MB_MAX_STEPS];
    CUtensorMap l;
    CUtensorMap vf[MB_MAX_STEPS];   // axis-0 boxes of the fused khs_kernel (its tile may be 128 columns wide)
};

struct MbGeom {
    int n;        // tile side
    int dpx;      // distance_in_px
    int intra;    // chromosome == chromosome2 (upper 2-fill applies)
    int dhi;      // last stored diagonal
    int wc;       // dhi - 3
    int vlo;      // first diagonal of V storage (2 - rmax)
    int wv;       // diagonals in V storage
    int nblk;     // blocks in this pass
    int ncta_h;   // CTAs per block in kh_kernel (for the partial-statistics layout)
    int dbg_step; // -1 or the step whose Gaussian is dumped to dbgG
    long long rec_cap;              // record capacity per block
    const double* raw;              // [nblk][n][wc]
    double* V;                      // [n_steps][nblk][n][wv]
    double* L;                      // [n_steps][nblk][n][wl]  DoG formed at each step, diagonals 2..dhi+2 (nullptr: not stored)
    int wl;                         // diagonals per row of L (even)
    long long plane_v;              // elements between consecutive (step, block) planes of V (>= n*wv, even)
    long long plane_l;              // same for L
    double* part_min;               // [nblk][n_scored][ncta_h]
    double* part_sum;               // [nblk][n_scored][ncta_h]
    unsigned long long* rec_count;  // [nblk]
    int* rec_row;                   // [nblk][rec_cap]
    int* rec_col;
    double* rec_v;
    int* rec_sidx;
    double* dbgG;                   // dense [n][n] (block 0 of the pass) or nullptr
    double* dbgL;                   // dense [n][n]: DoG formed at dbg_step
    double fill;                    // value of the constant regions (2.0; 0.0 for the difference stack of diff_mustache)
    double* dout;                   // [nblk][ndiff][n][wc]: DoG of every MB_FLAG_DIFFREF step (difference stack) or nullptr
    int ndiff;                      // MB_FLAG_DIFFREF steps of the difference chain (octaves)
    int zstride;                    // planes per step of the V / L scratch (its capacity in blocks): plane = step * zstride + zoff + b
    int zoff;                       // first plane of this pass inside a step (two passes may be in flight in two regions)
};

// ---------------------------------------------------------------------------------------------------------------
// tile shapes
// ---------------------------------------------------------------------------------------------------------------
// rows per CTA of the axis-0 pass (template parameter TH of kv_kernel), walked as 32-row sub-blocks: 128 amortises the
// staged +/- rmax halo rows best; 64 wastes less staging on the tiles that straddle the edges of a narrow band (a
// tile's rows shift the band by one column each)
constexpr int KV_TH_WIDE = 128, KV_TH_NARROW = 64;
constexpr int KV_TW = 32;      // columns per CTA = lanes
constexpr int KV_K = 4;        // outputs per thread along the filter axis
constexpr int KV_THREADS = (32 / KV_K) * 32;      // 256: one 32-row sub-block at a time

constexpr int KH_TR = 32;      // tile rows = lanes (axis-1 pass)
#ifndef MB_KH_TC
#define MB_KH_TC 64
#endif
constexpr int KH_TC = MB_KH_TC; // tile columns (64: two CTAs per SM; 128: one CTA of 16 warps with the whole SM as staging ring)
constexpr int KH_CTAS = (KH_TC == 64) ? 2 : 1;
constexpr int KH_K = 8;
constexpr int KH_THREADS = (KH_TC / KH_K) * 32;   // 256

constexpr int KS_TR = 32;      // scoring tile rows = lanes (30 scored + 2 halo)
constexpr int KS_TC = 64;      // scoring tile columns (62 scored + 2 halo)
constexpr int KS_K = 8;
constexpr int KS_THREADS = (KS_TC / KS_K) * 32;   // 256
constexpr int KS_SR = KS_TR - 2;   // scored rows per CTA
constexpr int KS_SC = KS_TC - 2;   // scored columns per CTA
constexpr int KS_PITCH = KS_TC + 6;   // 70: box width of the staged DoG tile (tile + even start column + row stagger); with
                                      // 70 = 6 (mod 16) and the one-column stagger of rows 8-15 / 24-31 the 16 lanes (= rows) of a
                                      // half warp read 16 different 8-byte bank pairs
// kh_kernel producer: 0 = warp 0 issues the boxes at the start of its steps and blocks until the boxes they overwrite are
// released; 1 = every warp, at the start of each of its steps, issues whatever box has become free (non-blocking test of the
// release barrier, claimed with a CAS on a shared counter): the warp that releases a box last is the one that sees it free
// first, and nobody waits for somebody else's release.  Measured slower (axis-1 pass 7.66 -> 8.00 ms on the 10k tile, 1.38 ->
// 1.47 ms on 24 x 2000^2): off
#ifndef MB_KH_COOP
#define MB_KH_COOP 0
#endif
#ifndef MB_KS_DEPTH
#define MB_KS_DEPTH 6
#endif
#ifndef MB_KS_CTAS
#define MB_KS_CTAS 2
#endif
constexpr int KS_DEPTH = MB_KS_DEPTH;           // ring stages per CTA: level being scored, the two before it, three in flight

// width of the staged box of a step with radius R: the filter support of the tile plus one element (the box must start
// on an even column: TMA needs 16-byte aligned box rows), padded to 2 (mod 4) elements so that the dense rows the TMA
// writes put the 32 lanes (= rows) on 8 different 8-byte bank pairs (2-way conflicts)
__host__ __device__ inline int kh_box_width(int R) {
    const int p = KH_TC + 2 * R + 2;          // + 1 for the even start column, + 1 for the row stagger of kh_kernel
    return (p % 4 == 2) ? p : p + 2;
}
__host__ __device__ inline int kh_vbuf_pitch(int rmax) { return kh_box_width(rmax); }
// staging ring of kh_kernel: half an SM's shared memory (two CTAs per SM), at least two of the widest boxes
#ifndef MB_KH_LOOKAHEAD
#define MB_KH_LOOKAHEAD 4
#endif
constexpr int KH_LOOKAHEAD = MB_KH_LOOKAHEAD;        // at most this many steps ahead of the slowest warp (6: no measurable difference)
#ifndef MB_KH_PREFETCH
#define MB_KH_PREFETCH 0
#endif
constexpr int KH_PREFETCH = MB_KH_PREFETCH;   // boxes beyond the ring that are prefetched into L2 (0: off; measured: 4 and 8
                                              // are 2-3 % slower on both the 4-octave tile and the 2000^2 blocks, profiles/README.md)
#ifndef MB_KS_PREFETCH
#define MB_KS_PREFETCH 0
#endif
constexpr int KS_PREFETCH = MB_KS_PREFETCH;   // DoG levels beyond the ring that are prefetched into L2 (0: off)
constexpr int KH_XP = KH_K + 1;        // pitch of the per-warp 32 x 8 transpose buffer (odd: lanes index rows)
__host__ __device__ inline int kh_ring_doubles(int rmax) {
    const int widest = KH_TR * kh_vbuf_pitch(rmax);
    const int half_sm = ((KH_CTAS == 2 ? 112 : 222) * 1024) / 8 - (KH_THREADS / 32) * KH_TR * KH_XP - 2 * MB_MAX_STEPS;
    return half_sm > 2 * widest ? half_sm : 2 * widest;
}
constexpr int KV_GUARD = 3;
__host__ __device__ inline size_t kv_smem_bytes(int rmax, int th) { return (size_t)(th + 2 * rmax + KV_GUARD) * KV_TW * sizeof(double); }
__host__ __device__ inline size_t kh_smem_bytes(int rmax, int n_scored) {
    (void)n_scored;
    // ring + full/empty mbarrier per step + per-warp transpose buffers for the coalesced DoG stores
    return ((size_t)kh_ring_doubles(rmax) + 2 * MB_MAX_STEPS + (KH_THREADS / 32) * KH_TR * KH_XP) * sizeof(double);
}

__host__ __device__ inline size_t ks_smem_bytes(int n_scored) {
    return ((size_t)KS_DEPTH * KS_TR * KS_PITCH + 2 * (size_t)(n_scored > 0 ? n_scored : 1) * (KS_THREADS / 32) + 2 * KS_DEPTH)
           * sizeof(double);                                                       // stages + per-warp statistics + mbarriers
}

// khs_kernel<TC> (axis-1 pass + DoG + scoring fused, the DoG levels never leave the SM): tile of 30 x (TC - 2) scored pixels +
// halo, without the row stagger (every row must hold the same columns because a pixel is compared with the rows above and
// below), three DoG levels in shared memory written by the threads themselves (odd pitch TC + 1: the 16 lanes of a half
// warp hit 16 different bank pairs), the rest of the CTA's shared memory is the ring of axis-0 boxes.
//   TC = 64:  8 warps, two CTAs per SM, 112 KB each  -- chains up to radius ~20 (two octaves)
//   TC = 128: 16 warps, one CTA per SM, 222 KB       -- up to radius 55 (four octaves): two boxes of 32 x 242 + 3 x 32 x 129
constexpr int KF_LSTAGES = 3;
__host__ __device__ inline int kf_box_width(int R, int tc) {
    const int p = tc + 2 * R + 2;             // filter support of the tile + 1 for the even start column (+ 1 spare)
    return (p % 4 == 2) ? p : p + 2;
}
__host__ __device__ inline int kf_total_doubles(int tc) { return ((tc == 64 ? 112 : 222) * 1024) / 8; }
__host__ __device__ inline int kf_ring_doubles(int tc) {
    return kf_total_doubles(tc) - KF_LSTAGES * KS_TR * (tc + 1) - 2 * MB_MAX_STEPS;
}
__host__ __device__ inline size_t kf_smem_bytes(int tc) { return (size_t)kf_total_doubles(tc) * sizeof(double); }
// the fused kernel needs at least two of the widest boxes in its ring
__host__ __device__ inline bool kf_fits(int rmax, int tc) {
    return 2 * ((KS_TR * kf_box_width(rmax, tc) + 15) & ~15) <= kf_ring_doubles(tc);
}

// scipy 'reflect' = (d c b a | a b c d | d c b a); |overshoot| < n is guaranteed by the host (n > 2*rmax)
__device__ __forceinline__ int reflect_idx(int i, int n) {
    if (i < 0) i = -1 - i;
    if (i >= n) i = 2 * n - 1 - i;
    return i;
}

// Tile value after the 2-fills (mustache.py:703-706); (i, j) must be inside the tile.
__device__ __forceinline__ double filled_at(const MbGeom& g, const double* __restrict__ rawb, int i, int j) {
    const int d = j - i;
    if (d <= 4) return g.fill;
    if (g.intra && d >= g.dpx + 1) return g.fill;
    if (d > g.dhi) return 0.0;
    return rawb[(size_t)i * g.wc + (d - 4)];
}

// ---------------------------------------------------------------------------------------------------------------
// folded symmetric correlation, K outputs per thread, register windows sliding one element per tap
//   x[q] = ctr[q * stride];  out[k] = x[k]*w0 + sum_{j=R..1} (x[k-j] + x[k+j]) * w[j]   (that order, no FMA)
// The tap loop is unrolled K times (the window registers rotate with period K).  R is rarely a multiple of K: the
// first trip enters the unrolled body at tap u0 = (K - R % K) % K (Duff's device) with the windows loaded in the state
// u0 taps would have left them in, so there is no remainder loop and no padded tap.
// ---------------------------------------------------------------------------------------------------------------
// acc + t * w: the reference's arithmetic (scipy multiplies, then adds: two roundings) or, in the opt-in fast mode
// (mb200_set_arithmetic), one fused multiply-add (one rounding, one FP64 instruction instead of two)
template <bool FAST>
__device__ __forceinline__ double mb_mac(double acc, double t, double w) {
    return FAST ? __fma_rn(t, w, acc) : __dadd_rn(acc, __dmul_rn(t, w));
}

// Radii below K never complete a trip of the unrolled loop; for them the whole support (K + 2R values) fits in
// registers and the sum is written out with compile-time indices: no window rotation, no loop, no entry switch.
template <int R, int K, int STRIDE, bool FAST>
__device__ __forceinline__ void conv_small(const double* __restrict__ ctr, const double* __restrict__ tp, double (&acc)[K]) {
    double x[K + 2 * R];
#pragma unroll
    for (int q = 0; q < K + 2 * R; ++q) x[q] = ctr[(q - R) * STRIDE];
    const double w0 = tp[0];
#pragma unroll
    for (int k = 0; k < K; ++k) acc[k] = __dmul_rn(x[k + R], w0);
#pragma unroll
    for (int j = R; j >= 1; --j) {
        const double w = tp[j];
#pragma unroll
        for (int k = 0; k < K; ++k) acc[k] = mb_mac<FAST>(acc[k], __dadd_rn(x[k + R - j], x[k + R + j]), w);
    }
}

template <int K, int STRIDE, bool FAST>
__device__ __forceinline__ void conv_slide(const double* __restrict__ ctr, const int R, const double* __restrict__ tp,
                                           double (&acc)[K]) {
    static_assert(K == 8, "the unrolled body below is written for K = 8");
    if (R < K) {
        switch (R) {
            case 1: conv_small<1, K, STRIDE, FAST>(ctr, tp, acc); break;
            case 2: conv_small<2, K, STRIDE, FAST>(ctr, tp, acc); break;
            case 3: conv_small<3, K, STRIDE, FAST>(ctr, tp, acc); break;
            case 4: conv_small<4, K, STRIDE, FAST>(ctr, tp, acc); break;
            case 5: conv_small<5, K, STRIDE, FAST>(ctr, tp, acc); break;
            case 6: conv_small<6, K, STRIDE, FAST>(ctr, tp, acc); break;
            default: conv_small<7, K, STRIDE, FAST>(ctr, tp, acc); break;
        }
        return;
    }
    double pl[K], pr[K];
    const double w0 = tp[0];
#pragma unroll
    for (int k = 0; k < K; ++k) acc[k] = __dmul_rn(ctr[k * STRIDE], w0);
    const int u0 = (K - (R & (K - 1))) & (K - 1);
    int j = R + u0;                                    // multiple of K; taps j .. j-u0+1 do not exist and are skipped
    // running pointers: every access of the unrolled body is [pointer + compile-time offset]
    const double* xl = ctr + (K - j) * STRIDE;         // tap u reloads pl[u] = x[u + K - j]
    const double* xr = ctr + (j - 1) * STRIDE;         //              and pr[K-1-u] = x[j - u - 1]
    const double* wj = tp + j;                         // its weight is wj[-u]
    // window state at the entry point: slot p of pl was reloaded by the skipped taps u < u0 (x[p + K - j]) or still
    // holds x[p - j]; slot p of pr was reloaded by u = K-1-p < u0 (x[p + j - K]) or still holds x[p + j]
#define MB_WIN(U0)                                                                                  \
    _Pragma("unroll") for (int p = 0; p < K; ++p) {                                                 \
        pl[p] = xl[(p < (U0) ? p : p - K) * STRIDE];                                                \
        pr[p] = xr[(p + 1 - ((K - 1 - p) < (U0) ? K : 0)) * STRIDE];                                \
    }
    switch (u0) {
        case 0: MB_WIN(0) break;
        case 1: MB_WIN(1) break;
        case 2: MB_WIN(2) break;
        case 3: MB_WIN(3) break;
        case 4: MB_WIN(4) break;
        case 5: MB_WIN(5) break;
        case 6: MB_WIN(6) break;
        default: MB_WIN(7) break;
    }
#undef MB_WIN
#define MB_TAP(u)                                                                                              \
    {                                                                                                          \
        const double w = wj[-(u)];                                                                             \
        double t[K];                                                                                           \
        _Pragma("unroll") for (int k = 0; k < K; ++k) t[k] = __dadd_rn(pl[(k + u) % K], pr[(k - u + K) % K]);  \
        if (FAST) {                                                                                            \
            _Pragma("unroll") for (int k = 0; k < K; ++k) acc[k] = __fma_rn(t[k], w, acc[k]);                  \
        } else {                                                                                               \
            _Pragma("unroll") for (int k = 0; k < K; ++k) t[k] = __dmul_rn(t[k], w);                           \
            _Pragma("unroll") for (int k = 0; k < K; ++k) acc[k] = __dadd_rn(acc[k], t[k]);                    \
        }                                                                                                      \
        pl[u % K] = xl[(u) * STRIDE];                                                                          \
        pr[(K - 1 - u) % K] = xr[-(u) * STRIDE];                                                               \
    }
    switch (u0) {
        case 0: do { MB_TAP(0)
        case 1: MB_TAP(1)
        case 2: MB_TAP(2)
        case 3: MB_TAP(3)
        case 4: MB_TAP(4)
        case 5: MB_TAP(5)
        case 6: MB_TAP(6)
        case 7: MB_TAP(7)
                xl += K * STRIDE; xr -= K * STRIDE; wj -= K;
                j -= K; } while (j > 0);
    }
#undef MB_TAP
}

// ---------------------------------------------------------------------------------------------------------------
// K_V: axis-0 pass for every step of the chain.  grid = (column tiles, row tiles, blocks), two CTAs per SM.
//
// Every Gaussian of the chain is an independent convolution of the SAME tile (mustache.py:719,725,734,751 all pass `c`),
// and in scipy's folded sum  out = x0*w0; for j = R..1: out += (x[-j] + x[j]) * w[j]  the pair sum (x[-j] + x[j]) does
// not depend on the step.  The steps are therefore processed in groups (KvPlan): a thread owns KV_K consecutive rows
// of one column, walks j from the group's largest radius down to 1, forms each pair sum once and feeds it to every
// step of the group.  Per step the additions still run j = R..1 (preceded by exact zeros), so every result is
// bit-identical to scipy's; the FP64 instruction count per output drops from sum(3R+1) = 2415 to 1987 for 4 octaves
// (532 -> 431 for 2).  The tap loop has no step-dependent control flow: four taps per trip from windows loaded once.
// ---------------------------------------------------------------------------------------------------------------
template <int N, bool FAST, int PITCH = KV_TW>
__device__ __forceinline__ void kv_group(const double* __restrict__ ctr, const double* __restrict__ tp, const int rtop,
                                         const int* __restrict__ steps, double* __restrict__ vrow, const long long step_stride,
                                         const int kstride, const unsigned vmask) {
    double acc[N][KV_K];
#pragma unroll
    for (int k = 0; k < KV_K; ++k) {
        const double x = ctr[k * PITCH];
#pragma unroll
        for (int s = 0; s < N; ++s) acc[s][k] = __dmul_rn(x, tp[s]);
    }
    // windows of a trip: xl[m] = x[m - jc], xr[m] = x[jc - 3 + m]; tap j = jc - u pairs xl[k + u] with xr[k + 3 - u].
    // Running pointers keep every access of the trip at [pointer + compile-time offset].
    const double* pl = ctr - rtop * PITCH;
    const double* pr = ctr + (rtop - 3) * PITCH;
    const double* pw = tp + rtop * N;
    for (int jc = rtop; jc > 0; jc -= 4) {
        double xl[KV_K + 3], xr[KV_K + 3];
#pragma unroll
        for (int m = 0; m < KV_K + 3; ++m) {
            xl[m] = pl[m * PITCH];
            xr[m] = pr[m * PITCH];
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            if (jc - u > 0) {
                double t[KV_K];
#pragma unroll
                for (int k = 0; k < KV_K; ++k) t[k] = __dadd_rn(xl[k + u], xr[k + 3 - u]);
#pragma unroll
                for (int s = 0; s < N; ++s) {
                    const double w = pw[s - u * N];
#pragma unroll
                    for (int k = 0; k < KV_K; ++k) acc[s][k] = mb_mac<FAST>(acc[s][k], t[k], w);
                }
            }
        }
        pl += 4 * PITCH;
        pr -= 4 * PITCH;
        pw -= 4 * N;
    }
#pragma unroll
    for (int s = 0; s < N; ++s) {
        double* vout = vrow + steps[s] * step_stride;
#pragma unroll
        for (int k = 0; k < KV_K; ++k) {
            if (vmask & (1u << k)) *vout = acc[s][k];
            vout += kstride;
        }
    }
}

template <int KV_TH, bool FAST, int GM = KV_GMAX>
__global__ void __launch_bounds__(KV_THREADS, GM <= KV_GSMALL ? 4 : 2)
kv_kernel(const __grid_constant__ KvPlan plan, const MbGeom g) {
    extern __shared__ double smem[];
    double* cs = smem + KV_GUARD * KV_TW;           // [(KV_TH + 2 rmax)][KV_TW]; the last window of a radius < 3 step
    const int rmax = plan.rmax;                     //   reaches up to KV_GUARD rows above the tile (loaded, never used)
    const int b = blockIdx.z;
    const int i0 = blockIdx.y * KV_TH;
    const int vhi = g.vlo + g.wv - 1;
    int jbase = i0 + g.vlo;                         // first column any row of this row tile can need
    if (jbase < 0) jbase = 0;
    const int j0 = jbase + blockIdx.x * KV_TW;
    const int ilast = min(i0 + KV_TH, g.n) - 1;
    if (j0 >= g.n || j0 > ilast + vhi) return;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int j = j0 + lane;

    // stage the filled tile (mustache.py:703-706 applied on the fly), 'reflect' rows: one warp per tile row, 16 rows
    // (= 16 independent loads per thread) in flight at a time
    {
        const double* rawb = g.raw + (size_t)b * g.n * g.wc;
        const int rows = KV_TH + 2 * rmax;
        constexpr int NWARP = KV_THREADS / 32, BATCH = 16;
        for (int r0 = warp; r0 < rows; r0 += NWARP * BATCH) {
            double v[BATCH];
#pragma unroll
            for (int q = 0; q < BATCH; ++q) {
                const int r = r0 + q * NWARP;
                // rows past n - 1 + rmax feed no stored output: clamped so that the reflection stays inside the tile
                const int ii = reflect_idx(min(i0 - rmax + min(r, rows - 1), g.n - 1 + rmax), g.n);
                const int d = j - ii;
                const bool banded = (j < g.n) && (d > 4) && (d <= g.dhi) && !(g.intra && d >= g.dpx + 1);
                double val = g.fill;                                // d <= 4, or intra and d >= dpx + 1
                if (j >= g.n || (d > g.dhi && !(g.intra && d >= g.dpx + 1))) val = 0.0;   // never used by a stored output
                if (banded) val = rawb[ii * g.wc + (d - 4)];
                v[q] = val;
            }
#pragma unroll
            for (int q = 0; q < BATCH; ++q) {
                const int r = r0 + q * NWARP;
                if (r < rows) cs[r * KV_TW + lane] = v[q];
            }
        }
    }
    __syncthreads();

    const long long step_stride = (long long)g.zstride * g.plane_v;
    double* const vblock = g.V + (long long)(g.zoff + b) * g.plane_v;
    const int kstride = g.wv - 1;                   // output k + 1 sits kstride elements after output k
    for (int sub = 0; sub < KV_TH / 32; ++sub) {
        const int rb = sub * 32 + warp * KV_K;      // first tile row of this thread's KV_K outputs
        const int i = i0 + rb;
        if (i >= g.n) break;                        // warp-uniform
        const double* ctr = cs + (rb + rmax) * KV_TW + lane;
        const int dmin_w = j0 - (i + KV_K - 1), dmax_w = j0 + KV_TW - 1 - i;
        unsigned vmask = 0;                         // outputs inside the image and inside the V storage band
#pragma unroll
        for (int k = 0; k < KV_K; ++k) {
            const int d = j - (i + k);
            if (i + k < g.n && j < g.n && d >= g.vlo && d <= vhi) vmask |= 1u << k;
        }
        double* const vrow = vblock + ((long long)i * g.wv + (j - i - g.vlo));
        for (int gi = 0; gi < plan.n_groups; ++gi) {
            const KvGroup& gr = plan.grp[gi];
            if (dmax_w < 2 - gr.rmax || dmin_w > g.dhi + 2 + gr.rmax) continue;     // warp-uniform: nobody reads these
            const double* tp = plan.tapsT + gr.tap_off;
            if (GM <= KV_GSMALL) {
                switch (gr.n) {
                    case 1: kv_group<1, FAST>(ctr, tp, gr.rmax, gr.step, vrow, step_stride, kstride, vmask); break;
                    case 2: kv_group<2, FAST>(ctr, tp, gr.rmax, gr.step, vrow, step_stride, kstride, vmask); break;
                    default: kv_group<3, FAST>(ctr, tp, gr.rmax, gr.step, vrow, step_stride, kstride, vmask); break;
                }
            } else {
                switch (gr.n) {
                    case 1: kv_group<1, FAST>(ctr, tp, gr.rmax, gr.step, vrow, step_stride, kstride, vmask); break;
                    case 2: kv_group<2, FAST>(ctr, tp, gr.rmax, gr.step, vrow, step_stride, kstride, vmask); break;
                    case 3: kv_group<3, FAST>(ctr, tp, gr.rmax, gr.step, vrow, step_stride, kstride, vmask); break;
                    case 4: kv_group<4, FAST>(ctr, tp, gr.rmax, gr.step, vrow, step_stride, kstride, vmask); break;
                    default: kv_group<5, FAST>(ctr, tp, gr.rmax, gr.step, vrow, step_stride, kstride, vmask); break;
                }
            }
        }
    }
}

// ---------------------------------------------------------------------------------------------------------------
// K_H: axis-1 pass and DoG chain, written to HBM for K_S.  grid = (col tiles, row tiles, blocks), two CTAs per SM.
// Thread (lane = tile row, warp = group of 8 tile columns) owns 8 pixels for the whole chain and keeps the previous
// Gaussian in registers.  The axis-0 tile of every step arrives as ONE 3-D TMA box (cp.async.bulk.tensor through the
// skewed tensor map, issued by an elected lane of warp 0) in a byte-granular shared-memory ring with a full / empty
// mbarrier per step, up to KH_LOOKAHEAD steps ahead; tiles whose halo leaves the image take a generic staging path.
// ---------------------------------------------------------------------------------------------------------------
// max / min of finite doubles: one compare + select (fmax / fmin add NaN handling that costs ~3x the instructions; the
// engine rejects non-finite tiles up front, mustache.py:755 would raise on them)
__device__ __forceinline__ double dmax(double a, double b) { return a > b ? a : b; }
__device__ __forceinline__ double dmin(double a, double b) { return a < b ? a : b; }

__device__ __forceinline__ double warp_min(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = dmin(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = __dadd_rn(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
// arrival that tells whether it completed the barrier's phase: the last of the expected arrivals always sees true, an
// earlier one may too if the last slips in between its two instructions (callers make the follow-up idempotent)
__device__ __forceinline__ bool mbar_arrive_is_last(uint64_t* bar) {
    uint32_t done;
    asm volatile(
        "{\n"
        ".reg .b64 tok;\n"
        ".reg .pred P1;\n"
        "mbarrier.arrive.shared::cta.b64 tok, [%1];\n"
        "mbarrier.test_wait.shared::cta.b64 P1, [%1], tok;\n"
        "selp.u32 %0, 1, 0, P1;\n"
        "}" : "=r"(done) : "r"(smem_u32(bar)) : "memory");
    return done != 0;
}
// non-blocking: has the phase with this parity completed?
__device__ __forceinline__ bool mbar_test(uint64_t* bar, uint32_t parity) {
    uint32_t done;
    asm volatile(
        "{\n"
        ".reg .pred P1;\n"
        "mbarrier.test_wait.parity.shared::cta.b64 P1, [%1], %2;\n"
        "selp.u32 %0, 1, 0, P1;\n"
        "}" : "=r"(done) : "r"(smem_u32(bar)), "r"(parity) : "memory");
    return done != 0;
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    asm volatile(
        "{\n"
        ".reg .pred P1;\n"
        "LAB_WAIT:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n"
        "@P1 bra DONE;\n"
        "bra LAB_WAIT;\n"
        "DONE:\n"
        "}" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
// one lane of a converged warp (elect.sync): keeps the TMA operands provably uniform for the compiler
__device__ __forceinline__ bool elect_one() {
    uint32_t pred;
    asm volatile("{\n.reg .pred P;\nelect.sync _|P, 0xffffffff;\nselp.u32 %0, 1, 0, P;\n}" : "=r"(pred));
    return pred != 0;
}

// TMA tiled copy global -> shared of one box of a 3-D tensor map (SASS: UTMALDG), completion in bytes on the mbarrier
__device__ __forceinline__ void tma_load_box3d(void* dst, const CUtensorMap* map, int x, int y, int z, uint64_t* bar) {
    asm volatile(
        "cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];" ::"r"(
            smem_u32(dst)),
        "l"(map), "r"(smem_u32(bar)), "r"(x), "r"(y), "r"(z)
        : "memory");
}

// TMA prefetch of a box into L2 only (no shared memory, no barrier): lets a kernel look further ahead than its
// shared-memory ring holds boxes, so that the real copy later is an L2 hit
__device__ __forceinline__ void tma_prefetch_box3d(const CUtensorMap* map, int x, int y, int z) {
    asm volatile("cp.async.bulk.prefetch.tensor.3d.L2.global.tile [%0, {%1, %2, %3}];" ::"l"(map), "r"(x), "r"(y), "r"(z) : "memory");
}

// MODE: KH_MAIN  DoG of every step -> L (scored by ks_kernel)
//       KH_DIFF  difference stack of diff_mustache: only the DoGs of MB_FLAG_DIFFREF steps are kept, in dout
//       KH_DEBUG KH_MAIN plus the dense dumps of mb200_debug_level
constexpr int KH_MAIN = 0, KH_DIFF = 1, KH_DEBUG = 2;

template <int MODE, bool FAST>
__global__ void __launch_bounds__(KH_THREADS, KH_CTAS)
kh_kernel(const __grid_constant__ MbProgram prog, const MbTensorMaps* __restrict__ tm, const MbGeom g) {
    extern __shared__ __align__(128) double smem[];
    constexpr int NW = KH_THREADS / 32;
    const int rmax = prog.rmax;
    double* vbuf = smem;                                        // byte-granular ring of staged boxes (TMA destination)
    uint64_t* full = reinterpret_cast<uint64_t*>(vbuf + kh_ring_doubles(rmax));    // [n_steps] bytes landed (used once)
    uint64_t* empty = full + MB_MAX_STEPS;                      // [n_steps] every warp is done reading the box (used once)
    double* xbuf = reinterpret_cast<double*>(empty + MB_MAX_STEPS) + (threadIdx.x >> 5) * (KH_TR * KH_XP);   // per warp [32][9]

    const int b = blockIdx.z;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int i0 = blockIdx.y * KH_TR;                  // first row of the tile
    const int js = i0 + 2 + blockIdx.x * KH_TC;         // first column: diagonal 2 of the first row
    const int ilast = min(i0 + KH_TR, g.n) - 1;
    if (js >= g.n || js > ilast + g.dhi + 2) return;    // tile entirely right of the band / of the image
    const int i = i0 + lane;                            // this thread's image row
    // Rows 8-15 and 24-31 of every tile are shifted one column to the right: with the even row pitch of the dense TMA box,
    // lanes (= rows) r and r+8 would hit the same 8-byte bank pair; the one-column stagger puts them on the odd pairs.
    // All tiles share the stagger, so they still partition the band.
    const int stag = (lane >> 3) & 1;
    const int c0 = warp * KH_K + stag;
    const int jc0 = js + c0;                            // image column of its first pixel
    const bool row_in = i < g.n;
    // tiles whose +/- rmax column halo leaves the image need 'reflect' indexing: generic (slow) staging for those
    const bool border = (js - rmax < 0) || (js + KH_TC + 1 + rmax > g.n);

    __shared__ int next_box;                                    // MB_KH_COOP: first box nobody has issued yet
    for (int t = threadIdx.x; t < prog.n_steps; t += KH_THREADS) {
        mbar_init(&full[t], 1);
        mbar_init(&empty[t], NW);
    }
    if (threadIdx.x == 0) next_box = 0;
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    __syncthreads();
    const int n_steps = prog.n_steps;
    // Producer side (one elected lane of warp 0): one TMA box copy per step -- the 32 x (TC + 2R + 2) axis-0 tile of step s
    // into its slot of the ring, after every warp released the boxes it overlaps.
    const MbStage* stg = prog.stage;
    int next_issue = 0, next_prefetch = 0;
    auto issue_ready = [&](int p_now) {
        if (KH_PREFETCH > 0) {                                  // boxes the ring cannot hold yet: into L2
            while (next_prefetch < n_steps && next_prefetch <= p_now + KH_LOOKAHEAD + KH_PREFETCH) {
                if (next_prefetch > p_now + KH_LOOKAHEAD && elect_one())
                    tma_prefetch_box3d(&tm->v[next_prefetch], (js - prog.st[next_prefetch].radius - g.vlo) & ~1, i0,
                                       next_prefetch * g.zstride + g.zoff + b);
                ++next_prefetch;
            }
        }
        while (next_issue < n_steps && next_issue <= p_now + KH_LOOKAHEAD) {
            const int dep = stg[next_issue].dep;
            if (dep >= p_now) break;                            // the overlapped box is still ahead of this warp
            if (elect_one()) {
                if (dep >= 0) mbar_wait(&empty[dep], 0);
                const int s = next_issue;
                const int R = prog.st[s].radius;
                mbar_arrive_expect_tx(&full[s], (uint32_t)(KH_TR * kh_box_width(R)) * 8u);
                // x = column index of the first needed element in the skewed view, floored to even (16-byte aligned rows)
                tma_load_box3d(vbuf + stg[next_issue].off, &tm->v[s], (js - R - g.vlo) & ~1, i0, s * g.zstride + g.zoff + b, &full[s]);
            }
            ++next_issue;
        }
    };
    // MB_KH_COOP: any warp, non-blocking
    auto issue_free = [&](int p_now) {
        for (int round = 0; round <= KH_LOOKAHEAD; ++round) {
            const int nb = *(volatile int*)&next_box;
            if (nb >= n_steps || nb > p_now + KH_LOOKAHEAD) break;
            const int dep = stg[nb].dep;
            if (dep >= 0 && !mbar_test(&empty[dep], 0)) break;               // the box it overwrites is still being read
            if (elect_one()) {
                if (atomicCAS(&next_box, nb, nb + 1) == nb) {
                    const int R = prog.st[nb].radius;
                    mbar_arrive_expect_tx(&full[nb], (uint32_t)(KH_TR * kh_box_width(R)) * 8u);
                    tma_load_box3d(vbuf + stg[nb].off, &tm->v[nb], (js - R - g.vlo) & ~1, i0, nb * g.zstride + g.zoff + b, &full[nb]);
                }
            }
            __syncwarp();
        }
    };
    if (!border && warp == 0) {                 // first boxes in flight before the store geometry below is set up
        if (MB_KH_COOP) issue_free(0);
        else issue_ready(0);
    }
    // any pixel of this warp's 32 x 9 chunk on a diagonal the detector reads (2 .. dhi+2)?
    const bool chunk_live = (js + warp * KH_K + KH_K - i0 >= 2) && (js + warp * KH_K - (i0 + KH_TR - 1) <= g.dhi + 2) &&
                            (js + warp * KH_K < g.n);

    // Everything the DoG store needs that does not depend on the step.  The owner layout (lane = row) would store 32
    // separate 8-byte pieces per instruction, so the 32 x 8 DoG chunk of the warp takes a trip through the warp's
    // transpose buffer: store instruction q writes tile rows q, q+8, q+16, q+24 (lane>>3 selects the row), 8 columns each.
    // With the buffer pitch of 9 doubles, rows 8 apart sit 72 = 8 (mod 16) doubles apart: the 16 lanes of a half warp read
    // 16 different 8-byte bank pairs (rows q..q+3 per instruction, as in round 1, made every read a 2-way conflict).
    unsigned zmask = 0;                                 // owner side: pixels that hold a DoG (others hold the cval 0)
#pragma unroll
    for (int k = 0; k < KH_K; ++k)
        if (row_in && jc0 + k < g.n) zmask |= 1u << k;
    const int kk = lane & 7;
    const int r0 = lane >> 3;                                   // this lane writes tile rows q + 8*r0
    const int pitch = (MODE == KH_DIFF) ? g.wc : g.wl;          // row length of the destination band
    const int dlo = (MODE == KH_DIFF) ? 4 : 2;                  // first diagonal it stores
    const int dhi_st = (MODE == KH_DIFF) ? g.dhi : g.dhi + 2;   // last one
    const int jst = js + warp * KH_K + kk + (r0 & 1);           // image column it writes (rows 8-15, 24-31 are staggered)
    unsigned qmask = 0;                                 // writer side: valid (row, column) of store instruction q
#pragma unroll
    for (int q = 0; q < KH_TR / 4; ++q) {
        const int ii = i0 + q + 8 * r0;
        const int d = jst - ii;
        bool ok = ii < g.n && d >= dlo && d <= dhi_st;
        if (MODE == KH_DIFF) ok = ok && jst < g.n;
        if (ok) qmask |= 1u << q;
    }
    // element offset of q = 0: row*pitch + (d - dlo) = row*(pitch-1) + col - dlo; q adds pitch - 1
    const long long qoff0 = (long long)(i0 + 8 * r0) * (pitch - 1) + jst - dlo;
    const int qstride = pitch - 1;
    // warp-uniform: no pixel of the chunk needs a mask (true for all but the tiles on the band / image edges)
    const bool interior = __all_sync(0xffffffffu, zmask == (1u << KH_K) - 1u && qmask == (1u << (KH_TR / 4)) - 1u);


    double gA[KH_K], gB[KH_K];
#pragma unroll
    for (int k = 0; k < KH_K; ++k) gA[k] = gB[k] = 0.0;

    // Gaussian of step s into gnew; gprev holds the Gaussian of step s - 1
    auto step = [&](const int s, const double (&gprev)[KH_K], double (&gnew)[KH_K]) {
        const int R = prog.st[s].radius;
        double* vst = vbuf + (border ? 0 : stg[s].off);
        const int bw = kh_box_width(R);                          // row pitch of this step's staged box
        const int shift = border ? 0 : ((js - R - g.vlo) & 1);   // the box starts one column early when that is odd
        if (!border) {
            if (MB_KH_COOP) issue_free(s);
            else if (warp == 0) issue_ready(s);
            mbar_wait(&full[s], 0);
        } else {
            __syncthreads();                                     // previous step's readers are done with the buffer
            const double* vin = g.V + ((size_t)s * g.zstride + g.zoff + b) * g.plane_v;
            const int wlen = KH_TC + 1 + 2 * R;
            // one warp per tile row, four rows x two column groups (= eight independent loads per thread) in flight at a time
            for (int t0 = lane; t0 < wlen; t0 += 64) {
                double val[2][KH_TR / NW];
#pragma unroll
                for (int h = 0; h < 2; ++h) {
                    const int t = t0 + 32 * h;
                    const int jj = reflect_idx(js - R + min(t, wlen - 1), g.n);
#pragma unroll
                    for (int q = 0; q < KH_TR / NW; ++q) {
                        const int ii = i0 + warp + q * NW;
                        const int dd = jj - ii - g.vlo;
                        val[h][q] = (ii < g.n && dd >= 0 && dd < g.wv) ? vin[(size_t)ii * g.wv + dd] : 0.0;
                    }
                }
#pragma unroll
                for (int h = 0; h < 2; ++h) {
                    const int t = t0 + 32 * h;
                    if (t < wlen) {
#pragma unroll
                        for (int q = 0; q < KH_TR / NW; ++q) vst[(warp + q * NW) * bw + t] = val[h][q];
                    }
                }
            }
            __syncthreads();
        }
        if (chunk_live && row_in) {
            conv_slide<KH_K, 1, FAST>(vst + lane * bw + shift + c0 + R, R, prog.taps + prog.st[s].tap_off, gnew);
        } else {
#pragma unroll
            for (int k = 0; k < KH_K; ++k) gnew[k] = 0.0;
        }
        if (!border) {                                           // release the box: one arrival per warp
            __syncwarp();
            if (lane == 0) mbar_arrive(&empty[s]);
        }
        if (MODE == KH_DEBUG) {
            if (g.dbgG != nullptr && s == g.dbg_step && b == 0 && row_in) {
#pragma unroll
                for (int k = 0; k < KH_K; ++k) {
                    const int j = jc0 + k;
                    if (j < g.n) g.dbgG[(size_t)i * g.n + j] = gnew[k];
                }
            }
        }
        const int sl = s;                                        // the DoG completed now: L_s = G_{s-1} - G_s
        const int flags = prog.st[sl].flags;
        const bool formed = !(flags & MB_FLAG_RESTART);
        const bool keep = (MODE == KH_DIFF) ? (formed && (flags & MB_FLAG_DIFFREF)) : formed;
        if (keep && chunk_live) {
            double* dst = (MODE == KH_DIFF) ? g.dout + ((size_t)b * g.ndiff + prog.st[sl].score_idx) * g.n * g.wc
                                            : g.L + ((size_t)sl * g.zstride + g.zoff + b) * g.plane_l;
            dst += qoff0;
            const double* xrd = xbuf + (8 * r0) * KH_XP + kk;
            if (MODE != KH_DEBUG && interior) {
                // every pixel of the warp's chunk is inside the image and on a stored diagonal: no masks
#pragma unroll
                for (int k = 0; k < KH_K; ++k) xbuf[lane * KH_XP + k] = __dsub_rn(gprev[k], gnew[k]);
                __syncwarp();
#pragma unroll
                for (int q = 0; q < KH_TR / 4; ++q) dst[(long long)q * qstride] = xrd[q * KH_XP];
            } else {
                // columns past the image hold the maximum filter's cval 0
#pragma unroll
                for (int k = 0; k < KH_K; ++k) xbuf[lane * KH_XP + k] = (zmask & (1u << k)) ? __dsub_rn(gprev[k], gnew[k]) : 0.0;
                __syncwarp();
#pragma unroll
                for (int q = 0; q < KH_TR / 4; ++q) {
                    if (qmask & (1u << q)) {
                        const double l = xrd[q * KH_XP];
                        *dst = l;
                        if (MODE == KH_DEBUG) {
                            if (g.dbgL != nullptr && sl == g.dbg_step && b == 0) {
                                const int ii = i0 + q + 8 * r0;
                                if (jst < g.n) g.dbgL[(size_t)ii * g.n + jst] = l;
                            }
                        }
                    }
                    dst += qstride;
                }
            }
            __syncwarp();
        }
    };

    int p = 0;
    for (; p + 1 < n_steps; p += 2) {
        step(p, gB, gA);
        step(p + 1, gA, gB);
    }
    if (p < n_steps) step(p, gB, gA);
}

// ---------------------------------------------------------------------------------------------------------------
// K_S (ks_kernel): zero-padded 3x3 maxima, 5-clause extremum test, running best and |L| statistics, streaming the DoG
// levels from HBM.  grid = (column tiles, row tiles, blocks); tile = 30 x 62 scored pixels + 1-pixel halo; thread (lane =
// tile row, warp = 8 tile columns) owns 8 pixels for the whole chain and keeps their state in registers (best response,
// winning level, own value of the previous DoG, "is a 3x3 maximum" bits of the two previous DoGs).  One TMA box per
// level into a KS_DEPTH-stage ring (full/empty mbarriers), two levels in flight while one is scored.
//
// The maximum filters are never materialised.  "L == max3x3(L)" (mustache.py:762-763) is "L >= its 8 neighbours" (the
// zero padding of mode='constant' is in the staged tile: rows outside the image arrive as zeros from the TMA, columns
// past it were written as zeros by kh_kernel) -- 8 chained compares per pixel and level.  The two strict clauses
// "Lc > max3x3(Lp)" and "Lc > max3x3(Ln)" (mustache.py:764-765) only matter for the few pixels that pass the cheap
// clauses first; for those the 9 values of the next level are already in registers and the 9 of the previous level are
// still in its ring stage (a stage is released two levels after it was filled).  Everything is exact FP64 comparison.
// ---------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(KS_THREADS, MB_KS_CTAS)
ks_kernel(const __grid_constant__ MbProgram prog, const MbTensorMaps* __restrict__ tm, const MbGeom g) {
    extern __shared__ __align__(128) double smem[];
    constexpr int NW = KS_THREADS / 32;
    constexpr int PL = KS_PITCH;                                // even, == 2 (mod 4)
    constexpr int D = KS_DEPTH;
    double* lst = smem;                                         // [D][KS_TR][PL]   staged DoG tiles
    double* pmin = lst + D * KS_TR * PL;                        // [n_scored][NW] per-warp min of |L|
    double* psum = pmin + (size_t)max(prog.n_scored, 1) * NW;   // [n_scored][NW] per-warp sum of |L|
    uint64_t* full = reinterpret_cast<uint64_t*>(psum + (size_t)max(prog.n_scored, 1) * NW);   // [D]
    uint64_t* empty = full + D;                                 // [D] every warp released the stage's tile
    __shared__ int stage_lvl[KS_DEPTH];                         // level the stage holds or is loading (claims reloads)
    __shared__ int lvl_step[MB_MAX_STEPS];                      // chain step of the nl-th DoG of the stream
    __shared__ int n_levels_s;

    const int b = blockIdx.z;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int is0 = blockIdx.y * KS_SR;                 // first scored row
    const int i0 = is0 - 1;                             // tile row 0 (halo)
    const int js = is0 + 4 + blockIdx.x * KS_SC;        // first scored column of this CTA
    const int ilast = min(is0 + KS_SR, g.n) - 1;
    const int cta = blockIdx.y * gridDim.x + blockIdx.x;
    const bool active = (js < g.n) && (js <= ilast + g.dhi);
    const double kInf = __longlong_as_double(0x7ff0000000000000LL);
    if (!active) {
        for (int t = threadIdx.x; t < prog.n_scored; t += KS_THREADS) {
            const size_t o = ((size_t)b * prog.n_scored + t) * g.ncta_h + cta;
            g.part_min[o] = kInf;
            g.part_sum[o] = 0.0;
        }
        return;
    }
    const double* rawb = g.raw + (size_t)b * g.n * g.wc;
    const int i = i0 + lane;                            // this thread's image row
    // Rows 8-15 and 24-31 of every tile are shifted one column to the right (as in kh_kernel): with the even row pitch of
    // the dense TMA box, lanes (= rows) r and r+8 would otherwise hit the same 8-byte bank pair.  All tiles share the
    // stagger, so they still partition the band (the column a staggered row gives up on the left is never in the mask:
    // it sits on a diagonal < 4 for every row but the first of the tile).
    const int stag = (lane >> 3) & 1;
    const int c0 = warp * KS_K + stag;                  // first tile column of this thread
    const int jc0 = js - 1 + c0;                        // image column of its first pixel
    const bool row_in = (i >= 0) && (i < g.n);
    const bool row_scored = (lane >= 1) && (lane <= KS_SR) && row_in;

    // the staged tile is a dense [KS_TR][PL] box; the box starts on the even column below tile column 0
    const int x_first = js - 1 - 2;                     // column index (skewed view of L) of tile column 0
    const int off_c = lane * PL + (x_first & 1) + c0;
    // the column left of tile column 0 is not staged: warp 0 clamps it on the rows that start there (it only feeds the
    // halo pixel, which is never scored); on the right the box is wide enough
    const int cl = (c0 == 0) ? 0 : -1;
    const int cr = KS_K;

    if (warp == 0) {
        // steps that form a DoG, in chain order: one step per lane, ranks from a ballot (a serial loop of thread 0 here kept
        // the other 255 threads of every CTA at the barrier below for ~4 % of the kernel's time)
        int nlev = 0;
        for (int s0 = 0; s0 < prog.n_steps; s0 += 32) {
            const int s = s0 + lane;
            const bool forms = s < prog.n_steps && !(prog.st[s].flags & MB_FLAG_RESTART);
            const unsigned m = __ballot_sync(0xffffffffu, forms);
            if (forms) lvl_step[nlev + __popc(m & ((1u << lane) - 1u))] = s;
            nlev += __popc(m);
        }
        if (lane < D) {
            mbar_init(&full[lane], 1);
            mbar_init(&empty[lane], NW);
            stage_lvl[lane] = lane;
        }
        if (lane == 0) n_levels_s = nlev;
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    const int n_levels = n_levels_s;

    // Producer side: the DoG tile of the nl-th level of the stream goes into stage nl % D.  A stage is free once every
    // warp has scored the level two after the one it holds; the warp whose release completes that (the last of the NW
    // arrivals on the stage's `empty` barrier) issues the stage's next load itself, so nobody ever waits for a free stage.
    auto issue = [&](int nl) {
        const int st = nl % D;
        mbar_arrive_expect_tx(&full[st], (uint32_t)(KS_TR * PL) * 8u);
        tma_load_box3d(lst + st * (KS_TR * PL), &tm->l, x_first & ~1, i0, lvl_step[nl] * g.zstride + g.zoff + b, &full[st]);
        if (KS_PREFETCH > 0 && nl + KS_PREFETCH < n_levels_s)       // the level that will take this stage's successor: into L2
            tma_prefetch_box3d(&tm->l, x_first & ~1, i0, lvl_step[nl + KS_PREFETCH] * g.zstride + g.zoff + b);
    };

    if (threadIdx.x == 0)                                       // prologue: every stage starts loading (before the mask
        for (int q = 0; q < D && q < n_levels; ++q) issue(q);   // reads below, so that the two latencies overlap)

    // mask bits of the 8 owned pixels (mustache.py:699: c != 0 and j - i >= 4, taken before the fills)
    unsigned mask = 0;
    if (row_scored) {
#pragma unroll
        for (int k = 0; k < KS_K; ++k) {
            const int c = c0 + k, j = jc0 + k, d = j - i;
            if (c >= 1 + stag && c <= KS_SC + stag && j < g.n && d >= 4 && d <= g.dhi) {
                if (rawb[(size_t)i * g.wc + (d - 4)] != 0.0) mask |= 1u << k;
            }
        }
    }

    double vbest[KS_K], lA[KS_K], lB[KS_K];
    unsigned long long lvl = 0;                         // 8 x uint8: scored index + 1 of the winning level, 0 = none
#pragma unroll
    for (int k = 0; k < KS_K; ++k) { vbest[k] = 0.0; lA[k] = lB[k] = 0.0; }
    unsigned e_cur = 0, e_prev = 0;                     // "L == max3x3(L)" bits of the two previous DoGs

    // One DoG level (stream position nl).  lcur holds the own values of level nl-1 (the one being scored), lown receives
    // those of level nl; the two register arrays swap roles between consecutive levels (the loop is unrolled by two).
    auto level = [&](const int s, const int nl, const double (&lcur)[KS_K], double (&lown)[KS_K]) {
        const int flags = prog.st[s].flags;
        const int u = nl / D, st_i = nl - u * D;
        mbar_wait(&full[st_i], u & 1);
        const double* st = lst + st_i * (KS_TR * PL) + off_c;                           // level nl, own pixel 0
        int sp_i = st_i - 2;
        if (sp_i < 0) sp_i += D;
        const double* sp = lst + sp_i * (KS_TR * PL) + off_c;                           // level nl-2, own pixel 0
        unsigned e_new = 0;
        double tmin = kInf, tsum = 0.0;
        const bool score = (flags & MB_FLAG_SCORE) != 0;
        const int sidx = prog.st[s].score_idx;
        if (row_scored) {
            double o[KS_K + 2], up[KS_K + 2], dn[KS_K + 2];      // rows i, i-1, i+1, tile columns c0-1 .. c0+8
            o[0] = st[cl]; up[0] = st[cl - PL]; dn[0] = st[cl + PL];
            o[KS_K + 1] = st[cr]; up[KS_K + 1] = st[cr - PL]; dn[KS_K + 1] = st[cr + PL];
#pragma unroll
            for (int k = 0; k < KS_K; ++k) {
                o[k + 1] = st[k];
                up[k + 1] = st[k - PL];
                dn[k + 1] = st[k + PL];
            }
#pragma unroll
            for (int k = 0; k < KS_K; ++k) {
                const double x = o[k + 1];
                lown[k] = x;
                const bool en = (x >= o[k]) && (x >= o[k + 2]) && (x >= up[k]) && (x >= up[k + 1]) && (x >= up[k + 2]) &&
                                (x >= dn[k]) && (x >= dn[k + 1]) && (x >= dn[k + 2]);
                if (en) e_new |= 1u << k;
            }
            if (score) {
                const unsigned cand = mask & e_cur & (e_prev | e_new);      // mustache.py:762-763
#pragma unroll
                for (int k = 0; k < KS_K; ++k) {
                    const unsigned bit = 1u << k;
                    const double x = lcur[k];
                    if (mask & bit) {                                       // expon.fit over the mask, mustache.py:755
                        const double a = fabs(x);
                        tmin = dmin(tmin, a);
                        tsum = __dadd_rn(tsum, a);
                    }
                    if ((cand & bit) && x > vbest[k]) {                     // mustache.py:761
                        // mustache.py:765  Lc > max3x3(Ln): the next level's 3x3 block is in registers
                        bool ok = (x > o[k]) && (x > o[k + 1]) && (x > o[k + 2]) && (x > up[k]) && (x > up[k + 1]) &&
                                  (x > up[k + 2]) && (x > dn[k]) && (x > dn[k + 1]) && (x > dn[k + 2]);
                        if (ok) {
                            // mustache.py:764  Lc > max3x3(Lp): the previous level's tile is still staged
                            const double* q = sp + k;
                            ok = (x > q[-1]) && (x > q[0]) && (x > q[1]) && (x > q[-PL - 1]) && (x > q[-PL]) && (x > q[-PL + 1]) &&
                                 (x > q[PL - 1]) && (x > q[PL]) && (x > q[PL + 1]);
                            if (ok) {
                                vbest[k] = x;
                                lvl = (lvl & ~(0xffULL << (8 * k))) | ((unsigned long long)(sidx + 1) << (8 * k));
                            }
                        }
                    }
                }
            }
        }
        __syncwarp();
        if (nl >= 2 && lane == 0 && mbar_arrive_is_last(&empty[sp_i])) {     // level nl-2 is released by every warp:
            if (nl - 2 + D < n_levels && atomicCAS(&stage_lvl[sp_i], nl - 2, nl - 2 + D) == nl - 2) issue(nl - 2 + D);   // reload
        }
        if (score) {                                            // per-warp statistics, fixed order (deterministic)
            tmin = warp_min(tmin);
            tsum = warp_sum(tsum);
            if (lane == 0) {
                pmin[sidx * NW + warp] = tmin;
                psum[sidx * NW + warp] = tsum;
            }
        }
        e_prev = e_cur;
        e_cur = e_new;
    };

    for (int nl = 0; nl < n_levels; nl += 2) {
        level(lvl_step[nl], nl, lB, lA);        // even level: own values into lA
        if (nl + 1 < n_levels) level(lvl_step[nl + 1], nl + 1, lA, lB);
    }
    __syncthreads();
    for (int t = threadIdx.x; t < prog.n_scored; t += KS_THREADS) {
        double mn = pmin[t * NW], sm = psum[t * NW];
        for (int w = 1; w < NW; ++w) {
            mn = dmin(mn, pmin[t * NW + w]);
            sm = __dadd_rn(sm, psum[t * NW + w]);
        }
        const size_t o = ((size_t)b * prog.n_scored + t) * g.ncta_h + cta;
        g.part_min[o] = mn;
        g.part_sum[o] = sm;
    }
    // ---- emit the pixels that were ever updated (pAll != 2, mustache.py:774) ----
    if (lvl != 0) {
#pragma unroll
        for (int k = 0; k < KS_K; ++k) {
            const int id = (int)((lvl >> (8 * k)) & 0xff);
            if (id) {
                const unsigned long long slot = atomicAdd(g.rec_count + b, 1ULL);
                if (slot < (unsigned long long)g.rec_cap) {
                    const size_t o = (size_t)b * g.rec_cap + slot;
                    g.rec_row[o] = i;
                    g.rec_col[o] = jc0 + k;
                    g.rec_v[o] = vbest[k];
                    g.rec_sidx[o] = id - 1;
                }
            }
        }
    }
}

// ---------------------------------------------------------------------------------------------------------------
// K_VH (kvh_kernel, opt-in: mb200_set_fusion bit 1): kv_kernel and kh_kernel in one, for chains up to radius KVH_RMAX
// (the default two octaves): the axis-0 results never go to HBM.  A CTA owns kh_kernel's 32 x 64 tile, stages the filled
// input tile with its +/- rmax halo once ('reflect' rows AND columns, so there is no separate border path), and then walks
// the chain two steps at a time:
//   phase A  axis-0 pass of both steps for the 32 rows x (64 + 2R) columns the axis-1 pass will read, into two
//            shared-memory tiles; the two steps share scipy's folded pair sums exactly as kv_kernel's groups do
//            (lanes = columns, 4 outputs per thread down a column);
//   barrier; phase B  kh_kernel's axis-1 pass + DoG + transposed store for each of the two steps (lanes = rows, odd
//            pitch: conflict-free); barrier.
// Per output the axis-0 work is done (64 + 2R) / 64 times over (the column halo every tile recomputes) and in groups of 2
// instead of up to 5: 1.24 x the FP64 instructions of kv + kh, against 8 B written + ~10 B read per bin and step less.
// ---------------------------------------------------------------------------------------------------------------
constexpr int KVH_RMAX = 14;
constexpr int KVH_RP = KH_TC + 2 * KVH_RMAX;          // 92: pitch of the staged input tile (lanes = columns)
constexpr int KVH_VP = KH_TC + 2 * KVH_RMAX + 1;      // 93: pitch of the axis-0 tiles (odd: lanes = rows in phase B)
constexpr int KVH_MAX_TAPS = 512;
__host__ __device__ inline size_t kvh_smem_bytes() {
    return ((size_t)(KH_TR + 2 * KVH_RMAX) * KVH_RP + 2 * KH_TR * KVH_VP + (KH_THREADS / 32) * KH_TR * KH_XP + KVH_MAX_TAPS) * sizeof(double);
}

template <bool FAST>
__global__ void __launch_bounds__(KH_THREADS, 2)
kvh_kernel(const __grid_constant__ MbProgram prog, const KvPlan* __restrict__ pairs, const MbGeom g) {
    extern __shared__ __align__(128) double smem[];
    constexpr int NW = KH_THREADS / 32;
    double* rawt = smem;                                               // [32 + 2 rmax][KVH_RP]
    double* vb = rawt + (KH_TR + 2 * KVH_RMAX) * KVH_RP;               // [2][32][KVH_VP]
    double* xbuf = vb + 2 * KH_TR * KVH_VP + (threadIdx.x >> 5) * (KH_TR * KH_XP);
    double* tapsA = vb + 2 * KH_TR * KVH_VP + NW * KH_TR * KH_XP;      // the pair plan's transposed taps

    const int b = blockIdx.z;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int rmax = prog.rmax;
    const int i0 = blockIdx.y * KH_TR;
    const int js = i0 + 2 + blockIdx.x * KH_TC;
    const int ilast = min(i0 + KH_TR, g.n) - 1;
    if (js >= g.n || js > ilast + g.dhi + 2) return;

    // ---- stage the filled tile (mustache.py:703-706), rows i0 - rmax .., columns js - rmax .., scipy 'reflect' both ways ----
    {
        const double* rawb = g.raw + (size_t)b * g.n * g.wc;
        const int rows = KH_TR + 2 * rmax, cols = KH_TC + 2 * rmax;
        for (int r = warp; r < rows; r += NW) {
            // rows past n - 1 + rmax feed no stored output: clamped so that the reflection stays inside the tile
            const int ii = reflect_idx(min(i0 - rmax + r, g.n - 1 + rmax), g.n);
            for (int c = lane; c < cols; c += 32) {
                const int jj = reflect_idx(min(js - rmax + c, g.n - 1 + rmax), g.n);
                const int d = jj - ii;
                double val = g.fill;                                    // d <= 4, or intra and d >= dpx + 1
                if (d > 4 && !(g.intra && d >= g.dpx + 1)) val = d <= g.dhi ? rawb[(size_t)ii * g.wc + (d - 4)] : 0.0;
                rawt[r * KVH_RP + c] = val;
            }
        }
        const int ntap = pairs->grp[pairs->n_groups - 1].tap_off + (pairs->grp[pairs->n_groups - 1].rmax + 1) * pairs->grp[pairs->n_groups - 1].n;
        for (int t = threadIdx.x; t < ntap; t += KH_THREADS) tapsA[t] = pairs->tapsT[t];
    }
    __syncthreads();

    const int i = i0 + lane;
    const int c0 = warp * KH_K;
    const int jc0 = js + c0;
    const bool row_in = i < g.n;
    const bool chunk_live = (js + warp * KH_K + KH_K - i0 >= 2) && (js + warp * KH_K - (i0 + KH_TR - 1) <= g.dhi + 2) &&
                            (js + warp * KH_K < g.n);
    // DoG store geometry, as in kh_kernel (no row stagger here: the axis-0 tiles have an odd pitch)
    unsigned zmask = 0;
#pragma unroll
    for (int k = 0; k < KH_K; ++k)
        if (row_in && jc0 + k < g.n) zmask |= 1u << k;
    const int kk = lane & 7, r0 = lane >> 3;
    const int pitch = g.wl;
    const int jst = js + warp * KH_K + kk;
    unsigned qmask = 0;
#pragma unroll
    for (int q = 0; q < KH_TR / 4; ++q) {
        const int ii = i0 + q + 8 * r0, d = jst - ii;
        if (ii < g.n && d >= 2 && d <= g.dhi + 2) qmask |= 1u << q;
    }
    const long long qoff0 = (long long)(i0 + 8 * r0) * (pitch - 1) + jst - 2;
    const int qstride = pitch - 1;
    const bool interior = __all_sync(0xffffffffu, zmask == (1u << KH_K) - 1u && qmask == (1u << (KH_TR / 4)) - 1u);

    double gprev[KH_K];
#pragma unroll
    for (int k = 0; k < KH_K; ++k) gprev[k] = 0.0;

    for (int gi = 0; gi < pairs->n_groups; ++gi) {
        const KvGroup& gr = pairs->grp[gi];
        const int R = gr.rmax;                                  // both steps run over the pair's larger radius
        // ---- phase A: axis-0 pass of the pair, V[slot][row][c], c = 0 .. 64 + 2R - 1  <->  image column js - R + c ----
        {
            const int W = KH_TC + 2 * R;
            const double* tp = tapsA + gr.tap_off;
            for (int c = lane; c < W; c += 32) {
                const double* ctr = rawt + (warp * KV_K + rmax) * KVH_RP + (rmax - R) + c;
                double* vrow = vb + (warp * KV_K) * KVH_VP + c;
                const int slots[2] = {0, 1};
                if (gr.n == 2) kv_group<2, FAST, KVH_RP>(ctr, tp, R, slots, vrow, (long long)KH_TR * KVH_VP, KVH_VP, 0xfu);
                else kv_group<1, FAST, KVH_RP>(ctr, tp, R, slots, vrow, (long long)KH_TR * KVH_VP, KVH_VP, 0xfu);
            }
        }
        __syncthreads();
        // ---- phase B: axis-1 pass + DoG of each step of the pair ----
        for (int slot = 0; slot < gr.n; ++slot) {
            const int s = gr.step[slot];
            const int Rs = prog.st[s].radius;
            double gnew[KH_K];
            if (chunk_live && row_in) {
                conv_slide<KH_K, 1, FAST>(vb + slot * (KH_TR * KVH_VP) + lane * KVH_VP + c0 + R, Rs, prog.taps + prog.st[s].tap_off, gnew);
            } else {
#pragma unroll
                for (int k = 0; k < KH_K; ++k) gnew[k] = 0.0;
            }
            const int flags = prog.st[s].flags;
            if (!(flags & MB_FLAG_RESTART) && chunk_live) {
                double* dst = g.L + ((size_t)s * g.zstride + g.zoff + b) * g.plane_l + qoff0;
                const double* xrd = xbuf + (8 * r0) * KH_XP + kk;
                if (interior) {
#pragma unroll
                    for (int k = 0; k < KH_K; ++k) xbuf[lane * KH_XP + k] = __dsub_rn(gprev[k], gnew[k]);
                    __syncwarp();
#pragma unroll
                    for (int q = 0; q < KH_TR / 4; ++q) dst[(long long)q * qstride] = xrd[q * KH_XP];
                } else {
#pragma unroll
                    for (int k = 0; k < KH_K; ++k) xbuf[lane * KH_XP + k] = (zmask & (1u << k)) ? __dsub_rn(gprev[k], gnew[k]) : 0.0;
                    __syncwarp();
#pragma unroll
                    for (int q = 0; q < KH_TR / 4; ++q) {
                        if (qmask & (1u << q)) *dst = xrd[q * KH_XP];
                        dst += qstride;
                    }
                }
                __syncwarp();
            }
#pragma unroll
            for (int k = 0; k < KH_K; ++k) gprev[k] = gnew[k];
        }
        __syncthreads();                                        // the pair's tiles are free for the next pair
    }
}

// ---------------------------------------------------------------------------------------------------------------
// K_HS (khs_kernel): kh_kernel and ks_kernel in one -- axis-1 pass, DoG, 3x3 maxima, extremum test, running best and |L|
// statistics -- for chains whose widest axis-0 box leaves room for three DoG levels in shared memory (kf_fits: the
// default two octaves).  The DoG levels are written by the threads that computed them into a three-stage ring and scored
// from there: `L` never goes to HBM (11 writes + 11 reads of 8 B per contact-bin and octave less, half of what the
// 2-octave path moved).  Tile = 30 x 62 scored pixels + 1-pixel halo, thread = 8 pixels of one row for the whole chain: G_{s-1}, the two last DoGs and the running best in
// registers.  Two CTA barriers per level: one before a level overwrites the stage of level-3 (everybody is done
// scoring with it), one after (the level is visible).  The axis-0 boxes arrive by TMA exactly as in kh_kernel.
// ---------------------------------------------------------------------------------------------------------------
template <int TC, bool FAST>
__global__ void __launch_bounds__((TC / KS_K) * 32, TC == 64 ? 2 : 1)
khs_kernel(const __grid_constant__ MbProgram prog, const MbTensorMaps* __restrict__ tm, const MbGeom g) {
    extern __shared__ __align__(128) double smem[];
    constexpr int NW = TC / KS_K;                               // warps: 8 tile columns each
    constexpr int NT = NW * 32;
    constexpr int PL = TC + 1;
    constexpr int SC = TC - 2;                                  // scored columns per CTA
    double* vbuf = smem;                                        // ring of staged axis-0 boxes (TMA destination)
    double* lst = vbuf + kf_ring_doubles(TC);                   // [3][KS_TR][PL] DoG levels
    uint64_t* full = reinterpret_cast<uint64_t*>(lst + KF_LSTAGES * KS_TR * PL);   // [n_steps]
    uint64_t* empty = full + MB_MAX_STEPS;

    const int b = blockIdx.z;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int is0 = blockIdx.y * KS_SR;                 // first scored row
    const int i0 = is0 - 1;                             // tile row 0 (halo)
    const int js = is0 + 4 + blockIdx.x * SC;           // first scored column of this CTA
    const int jt = js - 1;                              // image column of tile column 0
    const int ilast = min(is0 + KS_SR, g.n) - 1;
    const int cta = blockIdx.y * gridDim.x + blockIdx.x;
    const double kInf = __longlong_as_double(0x7ff0000000000000LL);
    // per-warp statistics go straight to part_min / part_sum [b][scored][cta * NW + warp] (g.ncta_h = CTAs * NW)
    if (!((js < g.n) && (js <= ilast + g.dhi))) {       // nothing to score here
        for (int t = threadIdx.x; t < prog.n_scored * NW; t += NT) {
            const size_t o = ((size_t)b * prog.n_scored + t / NW) * g.ncta_h + cta * NW + t % NW;
            g.part_min[o] = kInf;
            g.part_sum[o] = 0.0;
        }
        return;
    }
    const int rmax = prog.rmax;
    const double* rawb = g.raw + (size_t)b * g.n * g.wc;
    const int i = i0 + lane;                            // this thread's image row
    const int c0 = warp * KS_K;                         // first tile column of this thread
    const int jc0 = jt + c0;                            // image column of its first pixel
    const bool row_in = (i >= 0) && (i < g.n);
    const bool row_scored = (lane >= 1) && (lane <= KS_SR) && row_in;
    const bool border = (jt - rmax < 0) || (jt + TC + 1 + rmax > g.n);
    const int n_steps = prog.n_steps;

    for (int t = threadIdx.x; t < n_steps; t += NT) {
        mbar_init(&full[t], 1);
        mbar_init(&empty[t], NW);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    __syncthreads();

    // producer (one elected lane of warp 0), as in kh_kernel: one TMA box per step into its slot of the ring
    const MbStage* stg = prog.stage_f;
    int next_issue = 0;
    auto issue_ready = [&](int p_now) {
        while (next_issue < n_steps && next_issue <= p_now + KH_LOOKAHEAD) {
            const int dep = stg[next_issue].dep;
            if (dep >= p_now) break;
            if (elect_one()) {
                if (dep >= 0) mbar_wait(&empty[dep], 0);
                const int s = next_issue;
                const int R = prog.st[s].radius;
                mbar_arrive_expect_tx(&full[s], (uint32_t)(KS_TR * kf_box_width(R, TC)) * 8u);
                tma_load_box3d(vbuf + stg[s].off, &tm->vf[s], (jt - R - g.vlo) & ~1, i0, s * g.zstride + g.zoff + b, &full[s]);
            }
            ++next_issue;
        }
    };
    if (!border && warp == 0) issue_ready(0);

    // any pixel of this warp's 32 x 9 chunk on a diagonal the detector reads (2 .. dhi+2)?
    const bool chunk_live = (jt + warp * KS_K + KS_K - i0 >= 2) && (jt + warp * KS_K - (i0 + KS_TR - 1) <= g.dhi + 2) &&
                            (jt + warp * KS_K < g.n);
    unsigned zmask = 0;                                 // pixels inside the image (the others hold the cval 0)
    unsigned mask = 0;                                  // mask bits (mustache.py:699: c != 0 and j - i >= 4, before the fills)
#pragma unroll
    for (int k = 0; k < KS_K; ++k) {
        const int c = c0 + k, j = jc0 + k, d = j - i;
        if (row_in && j < g.n) zmask |= 1u << k;
        if (row_scored && c >= 1 && c <= SC && j < g.n && d >= 4 && d <= g.dhi) {
            if (rawb[(size_t)i * g.wc + (d - 4)] != 0.0) mask |= 1u << k;
        }
    }
    const int off_c = lane * PL + c0;                   // own pixel 0 inside a DoG stage
    const int cl = (c0 == 0) ? 0 : -1;                  // the column left of tile column 0 does not exist: clamp (halo only)

    double gprev[KS_K], vbest[KS_K];
#pragma unroll
    for (int k = 0; k < KS_K; ++k) { gprev[k] = vbest[k] = 0.0; }
    double pend_min = kInf, pend_sum = 0.0;             // min / sum of |L| over the mask for the level scored at the next step
    unsigned long long lvl = 0;                         // 8 x uint8: scored index + 1 of the winning level, 0 = none
    unsigned e_cur = 0, e_prev = 0;                     // "L == max3x3(L)" bits of the two previous DoGs
    int nl = 0;                                         // DoG levels formed so far

    // step s: Gaussian of the step (gprev holds the previous one); if the step forms a DoG: level nl = gprev - gnew, staged,
    // and level nl-1 scored against its two neighbours.  One copy of the code for every step (the Gaussian moves from gnew to
    // gprev at the end: 16 moves against ~10^3 instructions per step, and half the instruction footprint of a role swap).
    for (int s = 0; s < n_steps; ++s) {
        double gnew[KS_K], dnew[KS_K];
        const int R = prog.st[s].radius;
        double* vst = vbuf + (border ? 0 : stg[s].off);
        const int bw = kf_box_width(R, TC);
        const int shift = border ? 0 : ((jt - R - g.vlo) & 1);
        if (!border) {
            if (warp == 0) issue_ready(s);
            mbar_wait(&full[s], 0);
        } else {
            __syncthreads();
            const double* vin = g.V + ((size_t)s * g.zstride + g.zoff + b) * g.plane_v;
            const int wlen = TC + 1 + 2 * R;
            for (int r = warp; r < KS_TR; r += NW) {
                const int ii = i0 + r;
                for (int t = lane; t < wlen; t += 32) {
                    double val = 0.0;
                    if (ii >= 0 && ii < g.n) {
                        const int jj = reflect_idx(jt - R + t, g.n);
                        const int dd = jj - ii - g.vlo;
                        if (dd >= 0 && dd < g.wv) val = vin[(size_t)ii * g.wv + dd];
                    }
                    vst[r * bw + t] = val;
                }
            }
            __syncthreads();
        }
        if (chunk_live && row_in) {
            conv_slide<KS_K, 1, FAST>(vst + lane * bw + shift + c0 + R, R, prog.taps + prog.st[s].tap_off, gnew);
        } else {
#pragma unroll
            for (int k = 0; k < KS_K; ++k) gnew[k] = 0.0;
        }
        if (!border) {                                           // release the box: one arrival per warp
            __syncwarp();
            if (lane == 0) mbar_arrive(&empty[s]);
        }
        const int flags = prog.st[s].flags;
        if (flags & MB_FLAG_RESTART) {                           // first Gaussian of a chain: no DoG yet
#pragma unroll
            for (int k = 0; k < KS_K; ++k) gprev[k] = gnew[k];
            continue;
        }
        // ---- DoG level nl into its stage ----
        double* st = lst + (nl % KF_LSTAGES) * (KS_TR * PL) + off_c;
        __syncthreads();                                         // everybody is done scoring with level nl-3 (same stage)
        if (!border && warp == 0) issue_ready(s + 1);            // ... and has released box s: its slot may be refilled now
#pragma unroll
        for (int k = 0; k < KS_K; ++k) {
            dnew[k] = (zmask & (1u << k)) ? __dsub_rn(gprev[k], gnew[k]) : 0.0;
            st[k] = dnew[k];
            gprev[k] = gnew[k];
        }
        __syncthreads();                                         // level nl is visible
        // ---- score level nl-1 (ks_kernel's level(), own values from registers) ----
        const double* sp = lst + ((nl + KF_LSTAGES - 2) % KF_LSTAGES) * (KS_TR * PL) + off_c;    // level nl-2
        const double* sc = lst + ((nl + KF_LSTAGES - 1) % KF_LSTAGES) * (KS_TR * PL) + off_c;    // level nl-1 (being scored)
        unsigned e_new = 0;
        const double tmin = pend_min, tsum = pend_sum;           // statistics of level nl-1, taken when it was formed
        pend_min = kInf;
        pend_sum = 0.0;
        const bool score = (flags & MB_FLAG_SCORE) != 0;
        const int sidx = prog.st[s].score_idx;
        if (row_scored) {
            double o0, o9, up[KS_K + 2], dn[KS_K + 2];            // rows i-1, i+1: tile columns c0-1 .. c0+8; row i: the two ends
            o0 = st[cl]; up[0] = st[cl - PL]; dn[0] = st[cl + PL];
            o9 = st[KS_K]; up[KS_K + 1] = st[KS_K - PL]; dn[KS_K + 1] = st[KS_K + PL];
#pragma unroll
            for (int k = 0; k < KS_K; ++k) {
                up[k + 1] = st[k - PL];
                dn[k + 1] = st[k + PL];
            }
#pragma unroll
            for (int k = 0; k < KS_K; ++k) {
                const double x = dnew[k];
                const double xl = k == 0 ? o0 : dnew[k - 1], xr = k == KS_K - 1 ? o9 : dnew[k + 1];
                const bool en = (x >= xl) && (x >= xr) && (x >= up[k]) && (x >= up[k + 1]) && (x >= up[k + 2]) &&
                                (x >= dn[k]) && (x >= dn[k + 1]) && (x >= dn[k + 2]);
                if (en) e_new |= 1u << k;
            }
#pragma unroll
            for (int k = 0; k < KS_K; ++k) {
                if (mask & (1u << k)) {                                     // expon.fit over the mask (mustache.py:755), for the
                    const double a = fabs(dnew[k]);                         // step that scores this level
                    pend_min = dmin(pend_min, a);
                    pend_sum = __dadd_rn(pend_sum, a);
                }
            }
            if (score) {
                unsigned cand = mask & e_cur & (e_prev | e_new);            // mustache.py:762-763 (few pixels get this far)
                while (cand) {
                    const int k = __ffs(cand) - 1;
                    cand &= cand - 1;
                    const double x = sc[k];                                 // own value of level nl-1
                    double vb = 0.0;
#pragma unroll
                    for (int q = 0; q < KS_K; ++q) vb = (q == k) ? vbest[q] : vb;
                    if (x > vb) {                                           // mustache.py:761
                        // mustache.py:765  Lc > max3x3(Ln): the level just staged
                        const double* qn = st + k;
                        const int ql = (k == 0) ? cl : -1;                 // left neighbour of own pixel 0: clamped like cl
                        bool ok = (x > qn[ql]) && (x > qn[0]) && (x > qn[1]) && (x > qn[-PL + ql]) && (x > qn[-PL]) && (x > qn[-PL + 1]) &&
                                  (x > qn[PL + ql]) && (x > qn[PL]) && (x > qn[PL + 1]);
                        if (ok) {
                            // mustache.py:764  Lc > max3x3(Lp): the previous level is still staged
                            const double* q = sp + k;
                            ok = (x > q[ql]) && (x > q[0]) && (x > q[1]) && (x > q[-PL + ql]) && (x > q[-PL]) && (x > q[-PL + 1]) &&
                                 (x > q[PL + ql]) && (x > q[PL]) && (x > q[PL + 1]);
                            if (ok) {
#pragma unroll
                                for (int q2 = 0; q2 < KS_K; ++q2)
                                    if (q2 == k) vbest[q2] = x;
                                lvl = (lvl & ~(0xffULL << (8 * k))) | ((unsigned long long)(sidx + 1) << (8 * k));
                            }
                        }
                    }
                }
            }
        }
        if (score) {                                            // per-warp statistics, fixed order (deterministic)
            const double wmin = warp_min(tmin), wsum = warp_sum(tsum);
            if (lane == 0) {
                const size_t o = ((size_t)b * prog.n_scored + sidx) * g.ncta_h + cta * NW + warp;
                g.part_min[o] = wmin;
                g.part_sum[o] = wsum;
            }
        }
        e_prev = e_cur;
        e_cur = e_new;
        ++nl;
    }

    if (lvl != 0) {                                     // the pixels that were ever updated (pAll != 2, mustache.py:774)
#pragma unroll
        for (int k = 0; k < KS_K; ++k) {
            const int id = (int)((lvl >> (8 * k)) & 0xff);
            if (id) {
                const unsigned long long slot = atomicAdd(g.rec_count + b, 1ULL);
                if (slot < (unsigned long long)g.rec_cap) {
                    const size_t o = (size_t)b * g.rec_cap + slot;
                    g.rec_row[o] = i;
                    g.rec_col[o] = jc0 + k;
                    g.rec_v[o] = vbest[k];
                    g.rec_sidx[o] = id - 1;
                }
            }
        }
    }
}

// ---------------------------------------------------------------------------------------------------------------
// statistics and p-values
// ---------------------------------------------------------------------------------------------------------------
// grid = (n_scored, nblk); fixed-order tree so the result is deterministic.
__global__ void __launch_bounds__(256)
reduce_stats_kernel(const double* __restrict__ part_min, const double* __restrict__ part_sum, int ncta, int n_scored,
                    const unsigned long long* __restrict__ nz_count, double* __restrict__ fit_loc,
                    double* __restrict__ fit_scale) {
    __shared__ double smn[256], ssm[256];
    const int t = blockIdx.x, b = blockIdx.y;
    const double* pm = part_min + ((size_t)b * n_scored + t) * ncta;
    const double* ps = part_sum + ((size_t)b * n_scored + t) * ncta;
    double mn = __longlong_as_double(0x7ff0000000000000LL), sm = 0.0;
    for (int c = threadIdx.x; c < ncta; c += 256) {
        mn = fmin(mn, pm[c]);
        sm = __dadd_rn(sm, ps[c]);
    }
    smn[threadIdx.x] = mn;
    ssm[threadIdx.x] = sm;
    __syncthreads();
    for (int o = 128; o > 0; o >>= 1) {
        if (threadIdx.x < o) {
            smn[threadIdx.x] = fmin(smn[threadIdx.x], smn[threadIdx.x + o]);
            ssm[threadIdx.x] = __dadd_rn(ssm[threadIdx.x], ssm[threadIdx.x + o]);
        }
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        const double loc = smn[0];
        const double mean = ssm[0] / (double)nz_count[b];
        fit_loc[(size_t)b * n_scored + t] = loc;             // expon.fit: loc = min
        fit_scale[(size_t)b * n_scored + t] = mean - loc;    //            scale = mean - loc
    }
}

// p = 1 - expon.cdf(|L|, loc, scale) = 1 - (-expm1(-(x - loc)/scale))   (mustache.py:756); winners have L > 0.
// Also resolves the scored index of every record to what the caller reads back: the score id (octave*12 + i, the
// reference's scales[o][i] index) and, when the host registered them, the detection scale sigma itself (mustache.py:767).
__global__ void __launch_bounds__(256)
finalise_kernel(const unsigned long long* __restrict__ rec_count, long long rec_cap, const double* __restrict__ rec_v,
                const int* __restrict__ rec_sidx, int n_scored, const double* __restrict__ fit_loc,
                const double* __restrict__ fit_scale, const int* __restrict__ score_id, const double* __restrict__ score_sigma,
                double* __restrict__ rec_p, int* __restrict__ rec_sid, double* __restrict__ rec_sigma) {
    const int b = blockIdx.y;
    unsigned long long n = rec_count[b];
    if (n > (unsigned long long)rec_cap) n = rec_cap;
    for (unsigned long long r = blockIdx.x * 256ULL + threadIdx.x; r < n; r += (unsigned long long)gridDim.x * 256ULL) {
        const size_t o = (size_t)b * rec_cap + r;
        const int t = rec_sidx[o];
        const double y = (fabs(rec_v[o]) - fit_loc[(size_t)b * n_scored + t]) / fit_scale[(size_t)b * n_scored + t];
        rec_p[o] = 1.0 - (-expm1(-y));
        rec_sid[o] = score_id[t];
        rec_sigma[o] = score_sigma[t];
    }
}

// Records of the batch packed block after block (mb200_pack_records): grid = (x, blocks).
__global__ void __launch_bounds__(256)
pack_records_kernel(const long long* __restrict__ offsets, long long rec_cap, const int* __restrict__ row, const int* __restrict__ col,
                    const double* __restrict__ v, const int* __restrict__ sid, const int* __restrict__ sidx,
                    const double* __restrict__ p, const double* __restrict__ sigma, const double* __restrict__ pair,
                    int* __restrict__ orow, int* __restrict__ ocol, double* __restrict__ ov, int* __restrict__ osid,
                    int* __restrict__ osidx, double* __restrict__ op, double* __restrict__ osigma, double* __restrict__ opair) {
    const int b = blockIdx.y;
    const long long o0 = offsets[b], m = offsets[b + 1] - o0;
    const size_t s0 = (size_t)b * rec_cap;
    for (long long r = blockIdx.x * 256LL + threadIdx.x; r < m; r += (long long)gridDim.x * 256LL) {
        orow[o0 + r] = row[s0 + r];
        ocol[o0 + r] = col[s0 + r];
        ov[o0 + r] = v[s0 + r];
        osid[o0 + r] = sid[s0 + r];
        osidx[o0 + r] = sidx[s0 + r];
        op[o0 + r] = p[s0 + r];
        osigma[o0 + r] = sigma[s0 + r];
        if (pair != nullptr) opair[o0 + r] = pair[s0 + r];
    }
}

// ---------------------------------------------------------------------------------------------------------------
// tile preparation
// ---------------------------------------------------------------------------------------------------------------
// COO (block-local, duplicates already resolved by the host: last write wins, mustache.py:924) -> band tile
__global__ void __launch_bounds__(256)
scatter_coo_kernel(const int* __restrict__ rows, const int* __restrict__ cols, const double* __restrict__ vals,
                   long long nnz, double* __restrict__ rawb, int n, int wc, int dhi) {
    for (long long e = blockIdx.x * 256LL + threadIdx.x; e < nnz; e += (long long)gridDim.x * 256LL) {
        const int r = rows[e], c = cols[e], d = c - r;
        if (r >= 0 && r < n && c >= 0 && c < n && d >= 4 && d <= dhi) rawb[(size_t)r * wc + (d - 4)] = vals[e];
    }
}

// same for a whole batch: entries [offsets[b], offsets[b+1]) belong to block b (mb200_upload_coo_batch)
__global__ void __launch_bounds__(256)
scatter_coo_batch_kernel(const int* __restrict__ rows, const int* __restrict__ cols, const double* __restrict__ vals,
                         const long long* __restrict__ offsets, int nblk, double* __restrict__ raw, int n, int wc, int dhi) {
    const long long nnz = offsets[nblk];
    for (long long e = blockIdx.x * 256LL + threadIdx.x; e < nnz; e += (long long)gridDim.x * 256LL) {
        int lo = 0, hi = nblk;                          // last b with offsets[b] <= e
        while (hi - lo > 1) {
            const int mid = (lo + hi) >> 1;
            if (offsets[mid] <= e) lo = mid; else hi = mid;
        }
        const int r = rows[e], c = cols[e], d = c - r;
        if (r >= 0 && r < n && c >= 0 && c < n && d >= 4 && d <= dhi) raw[((size_t)lo * n + r) * wc + (d - 4)] = vals[e];
    }
}

// dense row-major tile on the device -> band tile
__global__ void __launch_bounds__(256)
band_from_dense_kernel(const double* __restrict__ c, long long ld, double* __restrict__ rawb, int n, int wc) {
    const long long total = (long long)n * wc;
    for (long long e = blockIdx.x * 256LL + threadIdx.x; e < total; e += (long long)gridDim.x * 256LL) {
        const int i = (int)(e / wc), k = (int)(e - (long long)i * wc);
        const int j = i + 4 + k;
        rawb[e] = (j < n) ? c[(long long)i * ld + j] : 0.0;
    }
}

// mask size (mustache.py:699-701) and a finiteness check (scipy.stats.expon.fit raises on non-finite data).
// grid = (row groups, blocks): a CTA walks whole band rows, so there is no per-element index arithmetic.
__global__ void __launch_bounds__(256)
count_mask_kernel(const double* __restrict__ raw, int n, int wc, unsigned long long* __restrict__ nz_count,
                  int* __restrict__ nonfinite) {
    const int b = blockIdx.y;
    const double* rawb = raw + (size_t)b * n * wc;
    unsigned cnt = 0;
    int bad = 0;
    for (int i = blockIdx.x; i < n; i += gridDim.x) {
        const double* row = rawb + (size_t)i * wc;
        const int w = min(wc, n - 4 - i);                       // columns i+4+k < n
        for (int k = threadIdx.x; k < w; k += 256) {
            const double v = row[k];
            cnt += (v != 0.0);
            bad |= !isfinite(v);
        }
    }
    for (int o = 16; o > 0; o >>= 1) {
        cnt += __shfl_xor_sync(0xffffffffu, cnt, o);
        bad |= __shfl_xor_sync(0xffffffffu, bad, o);
    }
    if ((threadIdx.x & 31) == 0) {
        if (cnt) atomicAdd(nz_count + b, (unsigned long long)cnt);
        if (bad) atomicOr(nonfinite + b, 1);
    }
}

// ---------------------------------------------------------------------------------------------------------------
// differential path (diff_mustache.py:262-425)
// ---------------------------------------------------------------------------------------------------------------
// c = c1 - c2 on the common mask, 0 elsewhere, taken AFTER the 2-fills (diff_mustache.py:268-276): diagonals 4 and
// >= dpx+1 are 2 - 2 = 0.  raw1/raw2/rawd: [npairs][n][wc] with the two maps interleaved (block 2k, 2k+1).
__global__ void __launch_bounds__(256)
diff_tile_kernel(const double* __restrict__ raw, double* __restrict__ rawd, int n, int wc, int dpx) {
    const int pr = blockIdx.y;
    const double* r1 = raw + (size_t)(2 * pr) * n * wc;
    const double* r2 = raw + (size_t)(2 * pr + 1) * n * wc;
    double* out = rawd + (size_t)pr * n * wc;
    const long long total = (long long)n * wc;
    for (long long e = blockIdx.x * 256LL + threadIdx.x; e < total; e += (long long)gridDim.x * 256LL) {
        const int i = (int)(e / wc), k = (int)(e - (long long)i * wc);
        const int d = k + 4;
        const double a = r1[e], b = r2[e];
        out[e] = (i + d < n && d >= 5 && d <= dpx && a != 0.0 && b != 0.0) ? __dsub_rn(a, b) : 0.0;
    }
}

// norm.fit over the common mask (diff_mustache.py:371): mean and population std (ddof 0), two passes.  Each pass is split
// over DIFF_NCH CTAs per (octave, pair) -- contiguous chunks, fixed-order tree inside a CTA, partials summed in chunk order
// by diff_finish_kernel -- so the result is deterministic and the pass runs at HBM speed instead of on 26 CTAs.
constexpr int DIFF_NCH = 64;

__device__ __forceinline__ double block_sum_256(double v, double* sh) {
    v = warp_sum(v);
    __syncthreads();
    if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = v;
    __syncthreads();
    double t = 0.0;
    if (threadIdx.x < 32) {
        t = threadIdx.x < 8 ? sh[threadIdx.x] : 0.0;
        t = warp_sum(t);
        if (threadIdx.x == 0) sh[8] = t;
    }
    __syncthreads();
    return sh[8];
}

// PASS 0: partial (sum, count) of the difference DoG over the common mask; PASS 1: partial sum of squared deviations from
// mu.  grid = (DIFF_NCH, ndiff, npairs); part[((pr * ndiff + o) * DIFF_NCH + chunk) * 2 + {0, 1}]
template <int PASS>
__global__ void __launch_bounds__(256)
diff_partial_kernel(const double* __restrict__ raw, const double* __restrict__ dout, int n, int wc, const double* __restrict__ mu,
                    double* __restrict__ part) {
    __shared__ double sh[9];
    const int ch = blockIdx.x, o = blockIdx.y, pr = blockIdx.z, ndiff = gridDim.y;
    const double* r1 = raw + (size_t)(2 * pr) * n * wc;
    const double* r2 = raw + (size_t)(2 * pr + 1) * n * wc;
    const double* dd = dout + ((size_t)pr * ndiff + o) * n * wc;
    const long long total = (long long)n * wc;
    const long long lo = total * ch / DIFF_NCH, hi = total * (ch + 1) / DIFF_NCH;
    const double mean = PASS == 1 ? mu[(size_t)pr * ndiff + o] : 0.0;
    double s = 0.0, cnt = 0.0;
    for (long long e = lo + threadIdx.x; e < hi; e += 256) {
        const int i = (int)(e / wc), k = (int)(e - (long long)i * wc);
        if (i + 4 + k < n && r1[e] != 0.0 && r2[e] != 0.0) {
            if (PASS == 0) {
                s = __dadd_rn(s, dd[e]);
                cnt += 1.0;
            } else {
                const double t = __dsub_rn(dd[e], mean);
                s = __dadd_rn(s, __dmul_rn(t, t));
            }
        }
    }
    const double tot = block_sum_256(s, sh);
    const double num = PASS == 0 ? block_sum_256(cnt, sh) : 0.0;
    if (threadIdx.x == 0) {
        double* p = part + (((size_t)pr * ndiff + o) * DIFF_NCH + ch) * 2;
        p[0] = tot;
        p[1] = num;
    }
}

// PASS 0: mu = sum / count (count kept in cntv); PASS 1: sd = sqrt(sum_sq / count).  One thread per (octave, pair).
template <int PASS>
__global__ void __launch_bounds__(64)
diff_finish_kernel(const double* __restrict__ part, int nitems, double* __restrict__ mu, double* __restrict__ sd,
                   double* __restrict__ cntv) {
    const int t = blockIdx.x * 64 + threadIdx.x;
    if (t >= nitems) return;
    double s = 0.0, c = 0.0;
    for (int ch = 0; ch < DIFF_NCH; ++ch) {
        s = __dadd_rn(s, part[((size_t)t * DIFF_NCH + ch) * 2]);
        c = __dadd_rn(c, part[((size_t)t * DIFF_NCH + ch) * 2 + 1]);
    }
    if (PASS == 0) {
        cntv[t] = c;
        mu[t] = s / c;
    } else {
        sd[t] = sqrt(s / cntv[t]);
    }
}

// two-sided normal p of the difference DoG at each record's pixel (diff_mustache.py:372-385, 412, 421)
__global__ void __launch_bounds__(256)
diff_pair_kernel(const unsigned long long* __restrict__ rec_count, long long rec_cap, const int* __restrict__ rec_row,
                 const int* __restrict__ rec_col, const int* __restrict__ rec_sidx,
                 const int* __restrict__ score_id, const double* __restrict__ dout, const double* __restrict__ mu,
                 const double* __restrict__ sd, int n, int wc, int ndiff, double* __restrict__ rec_pair) {
    const int b = blockIdx.y, pr = b >> 1;
    unsigned long long m = rec_count[b];
    if (m > (unsigned long long)rec_cap) m = rec_cap;
    for (unsigned long long r = blockIdx.x * 256ULL + threadIdx.x; r < m; r += (unsigned long long)gridDim.x * 256ULL) {
        const size_t o = (size_t)b * rec_cap + r;
        const int oct = score_id[rec_sidx[o]] / 12;                 // score id = octave*12 + i
        const int i = rec_row[o], j = rec_col[o];
        const double x = dout[((size_t)pr * ndiff + oct) * n * wc + (size_t)i * wc + (j - i - 4)];
        double p = normcdf((x - mu[(size_t)pr * ndiff + oct]) / sd[(size_t)pr * ndiff + oct]);
        if (!isfinite(p)) p = 1.0;                                  // np.nan_to_num(..., nan=1, posinf=1, neginf=1)
        if (p > 0.5) p = 1.0 - p;
        rec_pair[o] = 2.0 * p;
    }
}
