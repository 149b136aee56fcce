// Per-diagonal z-score normalisation of a chromosome's contact list on the device.
//
// Replaces mustache.py:622-686 (normalize_sparse), the producer of the values that get scattered into the tiles:
//   windowed branch (mustache.py:628-668): per diagonal d, a dense line of (v + 0.001), three box filters of
//     2 Mb / resolution bins (np.convolve(..., ones(w), mode='same') of the line, its square and its non-zero
//     indicator), local mean / variance with the global fallback where a window holds < 30 contacts, z-score, weight
//     1 + log30(1 + mean_d);
//   global branch (mustache.py:669-685): (v - mean_d) / std_d per diagonal.
// The reference spends O(n * w) per diagonal on the three convolutions; only the windows centred on contacts are ever
// read back (mustache.py:662-663), so the device evaluates exactly those: one thread per contact sums its window of the
// dense line.  np.mean / np.std are reproduced with numpy's pairwise summation (bit-exact); the window sums run left to
// right, which differs from the BLAS dot product behind np.convolve in the last bits (its order depends on the CPU the
// reference runs on), so normalised values agree to ~1e-13 relative, not bit for bit.  Every arithmetic step is an
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

template <class F>
__device__ double nz_pairwise_block(const double* __restrict__ a, long long n, F f) {
    if (n < 8) {
        double res = 0.0;
        for (long long i = 0; i < n; ++i) res = __dadd_rn(res, f(a[i]));
        return res;
    }
    double r[8];
#pragma unroll
    for (int q = 0; q < 8; ++q) r[q] = f(a[q]);
    long long i;
    for (i = 8; i < n - (n % 8); i += 8) {
#pragma unroll
        for (int q = 0; q < 8; ++q) r[q] = __dadd_rn(r[q], f(a[i + q]));
    }
    double res = __dadd_rn(__dadd_rn(__dadd_rn(r[0], r[1]), __dadd_rn(r[2], r[3])),
                           __dadd_rn(__dadd_rn(r[4], r[5]), __dadd_rn(r[6], r[7])));
    for (; i < n; ++i) res = __dadd_rn(res, f(a[i]));
    return res;
}

template <class F>
__device__ double nz_pairwise(const double* __restrict__ a, long long n, F f) {
    // explicit stack instead of recursion: frame = (start, length, phase, left result)
    long long st_a[40], st_n[40];
    double st_left[40];
    int st_phase[40];
    int sp = 0;
    st_a[0] = 0; st_n[0] = n; st_phase[0] = 0; st_left[0] = 0.0;
    double ret = 0.0;
    while (sp >= 0) {
        const long long a0 = st_a[sp], nn = st_n[sp];
        if (nn <= 128) {
            ret = nz_pairwise_block(a + a0, nn, f);
            --sp;
            continue;
        }
        long long n2 = nn / 2;
        n2 -= n2 % 8;
        if (st_phase[sp] == 0) {                 // descend into the left half
            st_phase[sp] = 1;
            ++sp;
            st_a[sp] = a0; st_n[sp] = n2; st_phase[sp] = 0;
        } else if (st_phase[sp] == 1) {          // left half done: keep it, descend into the right half
            st_left[sp] = ret;
            st_phase[sp] = 2;
            ++sp;
            st_a[sp] = a0 + n2; st_n[sp] = nn - n2; st_phase[sp] = 0;
        } else {                                 // both halves done
            ret = __dadd_rn(st_left[sp], ret);
            --sp;
        }
    }
    return ret;
}

// np.mean / np.std (ddof = 0) of every diagonal's values, in the order the contacts appear in the caller's arrays
// (mustache.py:638-643; NaN -> mean 0 / std 1 for an empty diagonal).  One thread per diagonal.
__global__ void __launch_bounds__(64)
nz_stats_kernel(const double* __restrict__ vs, const long long* __restrict__ seg, int ndiag, double* __restrict__ mean,
                double* __restrict__ sd) {
    const int d = blockIdx.x * 64 + threadIdx.x;
    if (d >= ndiag) return;
    const long long a0 = seg[d], m = seg[d + 1] - a0;
    if (m <= 0) {
        mean[d] = 0.0;
        sd[d] = 1.0;
        return;
    }
    const double mu = __ddiv_rn(nz_pairwise(vs + a0, m, NzIdentity()), (double)m);
    NzSqDev sq;
    sq.mean = mu;
    const double var = __ddiv_rn(nz_pairwise(vs + a0, m, sq), (double)m);
    const double s = __dsqrt_rn(var);
    mean[d] = isnan(mu) ? 0.0 : mu;
    sd[d] = isnan(s) ? 1.0 : s;
}

// dense lines: line[d][x] = v + 0.001 (mustache.py:634-635); line d starts at d * n
__global__ void __launch_bounds__(256)
nz_fill_kernel(const int* __restrict__ xs, const int* __restrict__ ds, const double* __restrict__ vs, long long nnz,
               long long n, double* __restrict__ lines) {
    for (long long e = blockIdx.x * 256LL + threadIdx.x; e < nnz; e += (long long)gridDim.x * 256LL)
        lines[(long long)ds[e] * n + xs[e]] = __dadd_rn(vs[e], 0.001);
}

// windowed z-score of every contact (mustache.py:645-668)
__global__ void __launch_bounds__(256)
nz_window_kernel(const int* __restrict__ xs, const int* __restrict__ ds, const long long* __restrict__ perm, long long nnz,
                 long long n, int w, const double* __restrict__ lines, const double* __restrict__ mean,
                 const double* __restrict__ sd, const double* __restrict__ weight, double* __restrict__ out) {
    const double kInf = __longlong_as_double(0x7ff0000000000000LL);
    for (long long e = blockIdx.x * 256LL + threadIdx.x; e < nnz; e += (long long)gridDim.x * 256LL) {
        const int d = ds[e], x = xs[e];
        const double* line = lines + (long long)d * n;
        const long long len = n - d;
        // np.convolve(line, ones(w), 'same')[x] = sum of line[lo..hi], lo = max(0, x + off - w + 1), hi = min(len-1, x + off)
        const long long off = (w - 1) / 2;
        long long lo = x + off - w + 1, hi = x + off;
        if (lo < 0) lo = 0;
        if (hi > len - 1) hi = len - 1;
        double cnt = 0.0, s = 0.0, s2 = 0.0;
        for (long long i = lo; i <= hi; ++i) {
            const double t = line[i];
            if (t != 0.0) cnt = __dadd_rn(cnt, 1.0);
            s = __dadd_rn(s, t);
            s2 = __dadd_rn(s2, __dmul_rn(t, t));
        }
        const double g_mean = mean[d], g_sd = sd[d];
        const double g_var = __dmul_rn(g_sd, g_sd);
        double var = __ddiv_rn(__dsub_rn(s2, __ddiv_rn(__dmul_rn(s, s), cnt)), __dsub_rn(cnt, 1.0));   // mustache.py:650
        if (isnan(var) || var == kInf || var == -kInf) var = g_var;
        double mu = __ddiv_rn(s, cnt);
        if (cnt < 30.0) {                                                                           // mustache.py:657-658
            mu = g_mean;
            var = g_var;
        }
        if (isnan(mu) || mu == kInf || mu == -kInf) mu = g_mean;
        const double lsd = __dsqrt_rn(var);
        double val = __ddiv_rn(__dsub_rn(line[x], mu), lsd);
        if (isnan(val) || val == kInf || val == -kInf) val = 0.0;
        out[perm[e]] = __dmul_rn(val, weight[d]);                                                    // mustache.py:667
    }
}

// global branch (mustache.py:669-685): contacts on diagonals < dlim get (v - mean_d) / std_d, everything is cleaned of
// NaN / Inf first.  `vs` already holds the cleaned values (the statistics were taken over them).
__global__ void __launch_bounds__(256)
nz_global_kernel(const int* __restrict__ ds, const long long* __restrict__ perm, const double* __restrict__ vs, long long nnz,
                 int dlim, const double* __restrict__ mean, const double* __restrict__ sd, double* __restrict__ out) {
    const double kInf = __longlong_as_double(0x7ff0000000000000LL);
    for (long long e = blockIdx.x * 256LL + threadIdx.x; e < nnz; e += (long long)gridDim.x * 256LL) {
        const int d = ds[e];
        double val = vs[e];
        if (d < dlim) {
            val = __ddiv_rn(__dsub_rn(val, mean[d]), sd[d]);
            if (isnan(val) || val == kInf || val == -kInf) val = 0.0;
        }
        out[perm[e]] = val;
    }
}
