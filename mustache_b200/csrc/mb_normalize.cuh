// Per-diagonal z-score normalisation of a chromosome's contact list on the device.
//
// Replaces mustache.py:622-686 (normalize_sparse), the producer of the values that get scattered into the tiles:
//   windowed branch (mustache.py:628-668): per diagonal d, a dense line of (v + 0.001), three box filters of
//     2 Mb / resolution bins (np.convolve(..., ones(w), mode='same') of the line, its square and its non-zero
//     indicator), local mean / variance with the global fallback where a window holds < 30 contacts, z-score, weight
//     1 + log30(1 + mean_d);
//   global branch (mustache.py:669-685): (v - mean_d) / std_d per diagonal.
// The reference spends O(n * w) per diagonal on the three convolutions of DENSE lines; only the windows centred on
// contacts are ever read back (mustache.py:662-663) and only contacts contribute to them, so the device works on the
// sparse contact list alone:
//   1. keys (diagonal << 24 | position) and a stable radix sort (mb_sort.cuh) group the contacts by diagonal, by position
//      inside a diagonal (for row-sorted input -- what the readers deliver -- that is the input order np.mean / np.std see);
//   2. np.mean / np.std per diagonal with numpy's pairwise summation, 8 lanes per diagonal = numpy's 8 interleaved
//      accumulators (bit-exact);
//   3. one warp per contact: binary search of its 2 Mb window in the diagonal's sorted positions, the three box sums
//      (count, sum, sum of squares of v + 0.001) over the contacts inside, lanes striding the window, fixed-order tree.
// The window sums differ from the BLAS dot product behind np.convolve in the last bits (whose order depends on the CPU
// the reference runs on), so normalised values agree to ~1e-13 relative, not bit for bit.  Every arithmetic step is an
// explicitly rounded intrinsic: nothing is contracted into FMA.
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

// numpy's pairwise sum (numpy/_core/src/umath/loops_utils.h.src, pairwise_sum_DOUBLE) of f(a[i]), i in [0, n):
// blocks of <= 128 elements with 8 interleaved accumulators, halves split at multiples of 8.
struct NzIdentity {
    __device__ __forceinline__ double operator()(double x) const { return x; }
};
struct NzSqDev {
    double mean;
    __device__ __forceinline__ double operator()(double x) const {
        const double t = __dsub_rn(x, mean);
        return __dmul_rn(t, t);
    }
};

// ---------------------------------------------------------------------------------------------------------------
// 1. keys
// ---------------------------------------------------------------------------------------------------------------
constexpr int NZ_POS_BITS = 24;                     // positions (bin indices) below 16.7 M

// key = min(|y - x|, D) << 24 | x, payload = input index.  Bucket D collects the contacts no diagonal loop visits.
__global__ void __launch_bounds__(256)
nz_keys_kernel(const int* __restrict__ x, const int* __restrict__ y, long long nnz, int D, unsigned long long* __restrict__ keys,
               unsigned* __restrict__ idx) {
    for (long long e = blockIdx.x * 256LL + threadIdx.x; e < nnz; e += (long long)gridDim.x * 256LL) {
        long long d = (long long)y[e] - (long long)x[e];
        if (d < 0) d = -d;
        if (d > D) d = D;
        keys[e] = ((unsigned long long)d << NZ_POS_BITS) | (unsigned long long)(unsigned)x[e];
        idx[e] = (unsigned)e;
    }
}

// seg[d] = first sorted contact with diagonal >= d, d = 0 .. D + 1 (seg[D + 1] = nnz)
__global__ void __launch_bounds__(256)
nz_segments_kernel(const unsigned long long* __restrict__ keys, long long nnz, int D, long long* __restrict__ seg) {
    const int d = blockIdx.x * 256 + threadIdx.x;
    if (d > D + 1) return;
    long long lo = 0, hi = nnz;                     // first index with (key >> 24) >= d
    while (lo < hi) {
        const long long mid = (lo + hi) >> 1;
        if ((long long)(keys[mid] >> NZ_POS_BITS) < d) lo = mid + 1; else hi = mid;
    }
    seg[d] = lo;
}

// values and positions in sorted order; the global branch cleans NaN / Inf first (mustache.py:672)
__global__ void __launch_bounds__(256)
nz_gather_kernel(const unsigned long long* __restrict__ keys, const unsigned* __restrict__ idx, const double* __restrict__ v,
                 long long nnz, int clean, int* __restrict__ xs, double* __restrict__ vs) {
    const double kInf = __longlong_as_double(0x7ff0000000000000LL);
    for (long long o = blockIdx.x * 256LL + threadIdx.x; o < nnz; o += (long long)gridDim.x * 256LL) {
        double val = v[idx[o]];
        if (clean && (isnan(val) || val == kInf || val == -kInf)) val = 0.0;
        xs[o] = (int)(keys[o] & ((1ULL << NZ_POS_BITS) - 1));
        vs[o] = val;
    }
}

// ---------------------------------------------------------------------------------------------------------------
// 2. np.mean / np.std (ddof = 0) of every diagonal's values (mustache.py:638-643; NaN -> mean 0 / std 1 for an empty
// diagonal).  Eight lanes per diagonal: lane q is accumulator r[q] of numpy's unrolled block loop, so a block of 128
// values is eight coalesced 64-byte reads per step; the recursion over halves is uniform across the eight lanes.
// ---------------------------------------------------------------------------------------------------------------
template <class F>
__device__ double nz_pairwise_block8(const double* __restrict__ a, long long n, F f, int q, unsigned gmask) {
    if (n < 8) {
        double res = 0.0;
        for (long long i = 0; i < n; ++i) res = __dadd_rn(res, f(a[i]));
        return res;
    }
    double r = f(a[q]);
    long long i;
    for (i = 8; i < n - (n % 8); i += 8) r = __dadd_rn(r, f(a[i + q]));
    // ((r0 + r1) + (r2 + r3)) + ((r4 + r5) + (r6 + r7)): additions commute, so the butterfly gives every lane that value
    r = __dadd_rn(r, __shfl_xor_sync(gmask, r, 1));
    r = __dadd_rn(r, __shfl_xor_sync(gmask, r, 2));
    r = __dadd_rn(r, __shfl_xor_sync(gmask, r, 4));
    for (; i < n; ++i) r = __dadd_rn(r, f(a[i]));
    return r;
}

template <class F>
__device__ double nz_pairwise8(const double* __restrict__ a, long long n, F f, int q, unsigned gmask) {
    long long st_a[40], st_n[40];
    double st_left[40];
    int st_phase[40];
    int sp = 0;
    st_a[0] = 0; st_n[0] = n; st_phase[0] = 0; st_left[0] = 0.0;
    double ret = 0.0;
    while (sp >= 0) {
        const long long a0 = st_a[sp], nn = st_n[sp];
        if (nn <= 128) {
            ret = nz_pairwise_block8(a + a0, nn, f, q, gmask);
            --sp;
            continue;
        }
        long long n2 = nn / 2;
        n2 -= n2 % 8;
        if (st_phase[sp] == 0) {
            st_phase[sp] = 1;
            ++sp;
            st_a[sp] = a0; st_n[sp] = n2; st_phase[sp] = 0;
        } else if (st_phase[sp] == 1) {
            st_left[sp] = ret;
            st_phase[sp] = 2;
            ++sp;
            st_a[sp] = a0 + n2; st_n[sp] = nn - n2; st_phase[sp] = 0;
        } else {
            ret = __dadd_rn(st_left[sp], ret);
            --sp;
        }
    }
    return ret;
}

__global__ void __launch_bounds__(256)
nz_stats_kernel(const double* __restrict__ vs, const long long* __restrict__ seg, int ndiag, double* __restrict__ mean,
                double* __restrict__ sd) {
    const int q = threadIdx.x & 7;
    const unsigned gmask = 0xffu << (threadIdx.x & 24);           // the eight lanes of this diagonal
    const int d = (blockIdx.x * 256 + threadIdx.x) >> 3;
    if (d >= ndiag) return;                                       // whole groups leave together
    const long long a0 = seg[d], m = seg[d + 1] - a0;
    if (m <= 0) {
        if (q == 0) { mean[d] = 0.0; sd[d] = 1.0; }
        return;
    }
    const double mu = __ddiv_rn(nz_pairwise8(vs + a0, m, NzIdentity(), q, gmask), (double)m);
    NzSqDev sq;
    sq.mean = mu;
    const double var = __ddiv_rn(nz_pairwise8(vs + a0, m, sq, q, gmask), (double)m);
    const double s = __dsqrt_rn(var);
    if (q == 0) {
        mean[d] = isnan(mu) ? 0.0 : mu;
        sd[d] = isnan(s) ? 1.0 : s;
    }
}

// ---------------------------------------------------------------------------------------------------------------
// 3. windowed z-score of every contact (mustache.py:645-668), one warp per contact
// ---------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ double nz_warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = __dadd_rn(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}

__global__ void __launch_bounds__(256)
nz_window_kernel(const unsigned long long* __restrict__ keys, const int* __restrict__ xs, const double* __restrict__ vs,
                 const unsigned* __restrict__ idx, const long long* __restrict__ seg, long long m, long long n, int w,
                 const double* __restrict__ mean, const double* __restrict__ sd, const double* __restrict__ weight,
                 double* __restrict__ out) {
    const double kInf = __longlong_as_double(0x7ff0000000000000LL);
    const int lane = threadIdx.x & 31;
    const long long nwarps = (long long)gridDim.x * 8;
    for (long long o = blockIdx.x * 8LL + (threadIdx.x >> 5); o < m; o += nwarps) {
        const int d = (int)(keys[o] >> NZ_POS_BITS), x = xs[o];
        const long long len = n - d;
        // np.convolve(line, ones(w), 'same')[x] = sum of line[lo..hi], lo = max(0, x + off - w + 1), hi = min(len-1, x + off)
        const long long off = (w - 1) / 2;
        long long lo = x + off - w + 1, hi = x + off;
        if (lo < 0) lo = 0;
        if (hi > len - 1) hi = len - 1;
        // contacts of this diagonal with position in [lo, hi]: positions are sorted inside the diagonal's segment
        long long a = seg[d], b = seg[d + 1];
        {
            long long l = a, h = b;
            while (l < h) { const long long mid = (l + h) >> 1; if (xs[mid] < lo) l = mid + 1; else h = mid; }
            a = l;
            h = b;
            while (l < h) { const long long mid = (l + h) >> 1; if (xs[mid] <= hi) l = mid + 1; else h = mid; }
            b = l;
        }
        double cnt = 0.0, s = 0.0, s2 = 0.0;
        for (long long i = a + lane; i < b; i += 32) {
            const double t = __dadd_rn(vs[i], 0.001);                                       // mustache.py:635
            if (t != 0.0) cnt = __dadd_rn(cnt, 1.0);
            s = __dadd_rn(s, t);
            s2 = __dadd_rn(s2, __dmul_rn(t, t));
        }
        cnt = nz_warp_sum(cnt);
        s = nz_warp_sum(s);
        s2 = nz_warp_sum(s2);
        if (lane == 0) {
            const double g_mean = mean[d], g_sd = sd[d];
            const double g_var = __dmul_rn(g_sd, g_sd);
            double var = __ddiv_rn(__dsub_rn(s2, __ddiv_rn(__dmul_rn(s, s), cnt)), __dsub_rn(cnt, 1.0));   // mustache.py:650
            if (isnan(var) || var == kInf || var == -kInf) var = g_var;
            double mu = __ddiv_rn(s, cnt);
            if (cnt < 30.0) {                                                                       // mustache.py:657-658
                mu = g_mean;
                var = g_var;
            }
            if (isnan(mu) || mu == kInf || mu == -kInf) mu = g_mean;
            const double lsd = __dsqrt_rn(var);
            double val = __ddiv_rn(__dsub_rn(__dadd_rn(vs[o], 0.001), mu), lsd);
            if (isnan(val) || val == kInf || val == -kInf) val = 0.0;
            out[idx[o]] = __dmul_rn(val, weight[d]);                                                // mustache.py:667
        }
    }
}

// global branch (mustache.py:669-685): contacts on diagonals < dlim get (v - mean_d) / std_d; everything was cleaned of
// NaN / Inf first (`vs` holds the cleaned values, the statistics were taken over them).
__global__ void __launch_bounds__(256)
nz_global_kernel(const unsigned long long* __restrict__ keys, const unsigned* __restrict__ idx, const double* __restrict__ vs,
                 long long nnz, int dlim, const double* __restrict__ mean, const double* __restrict__ sd, double* __restrict__ out) {
    const double kInf = __longlong_as_double(0x7ff0000000000000LL);
    for (long long o = blockIdx.x * 256LL + threadIdx.x; o < nnz; o += (long long)gridDim.x * 256LL) {
        const int d = (int)(keys[o] >> NZ_POS_BITS);
        double val = vs[o];
        if (d < dlim) {
            val = __ddiv_rn(__dsub_rn(val, mean[d]), sd[d]);
            if (isnan(val) || val == kInf || val == -kInf) val = 0.0;
        }
        out[idx[o]] = val;
    }
}
