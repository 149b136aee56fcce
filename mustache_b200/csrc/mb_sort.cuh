// Stable LSD radix sort on the device (segmented), hand-written for the two places of the path that need an order:
//   * Benjamini-Hochberg per block (mustache.py:778): the found p-values of every block sorted ascending
//     (64-bit keys = bit patterns of the positive doubles, 8 passes of 8 bits, one segment per block);
//   * normalize_sparse (mustache.py:632-633): the contacts grouped by diagonal in input order (one segment, keys = |y - x|).
// One pass = four short kernels: per-tile digit histograms, an exclusive scan over the tiles per digit, one over the
// digits, and a stable scatter.  Stability inside a tile comes from warp-ordered ranking: every warp owns a contiguous slice of the tile,
// walks it 32 keys at a time, and ranks equal digits with __match_any_sync in lane order.
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

constexpr int RS_BITS = 8;
constexpr int RS_RADIX = 1 << RS_BITS;
constexpr int RS_THREADS = 256;
constexpr int RS_WARPS = RS_THREADS / 32;
constexpr int RS_ITEMS = 16;                               // keys per thread
constexpr int RS_TILE = RS_THREADS * RS_ITEMS;             // 4096 keys per CTA

// Segment s holds keys [seg_off[s], seg_off[s] + seg_len(s)) of the arrays; seg_len comes from `counts` (clamped to cap)
// when counts != nullptr (the per-block record counters), else from seg_off[s + 1] - seg_off[s].
struct RsSegments {
    const long long* seg_off;              // element offset of every segment (nseg + 1 entries when counts == nullptr)
    const unsigned long long* counts;      // optional per-segment length
    long long cap;                         // clamp for counts
    long long stride;                      // when seg_off == nullptr: segment s starts at s * stride
};

__device__ __forceinline__ void rs_segment(const RsSegments& sg, int s, long long& off, long long& len) {
    off = sg.seg_off ? sg.seg_off[s] : (long long)s * sg.stride;
    if (sg.counts) {
        unsigned long long c = sg.counts[s];
        len = (long long)(c > (unsigned long long)sg.cap ? (unsigned long long)sg.cap : c);
    } else {
        len = sg.seg_off[s + 1] - off;
    }
}

// hist[(seg * RS_RADIX + digit) * ntiles + tile]
__global__ void __launch_bounds__(RS_THREADS)
rs_hist_kernel(const unsigned long long* __restrict__ keys, RsSegments sg, int shift, int ntiles, unsigned* __restrict__ hist) {
    __shared__ unsigned h[RS_RADIX];
    const int s = blockIdx.y, tile = blockIdx.x;
    long long off, len;
    rs_segment(sg, s, off, len);
    h[threadIdx.x] = 0;
    __syncthreads();
    const long long t0 = (long long)tile * RS_TILE;
    for (int it = 0; it < RS_ITEMS; ++it) {
        const long long i = t0 + it * RS_THREADS + threadIdx.x;
        if (i < len) atomicAdd(&h[(unsigned)(keys[off + i] >> shift) & (RS_RADIX - 1)], 1u);
    }
    __syncthreads();
    hist[((size_t)s * RS_RADIX + threadIdx.x) * ntiles + tile] = h[threadIdx.x];
}

// Exclusive scan of hist over (digit-major, tile-minor) per segment, in two parallel steps:
//   rs_scan_tiles_kernel  grid (RS_RADIX, nseg): one CTA per digit scans that digit's counts over the tiles in place and
//                         leaves the digit's total in tot[seg * RS_RADIX + digit];
//   rs_scan_digits_kernel grid (nseg): exclusive scan of the RS_RADIX totals in place -> first position of every digit.
// The scatter adds the two.
__global__ void __launch_bounds__(256)
rs_scan_tiles_kernel(unsigned* __restrict__ hist, int ntiles, unsigned* __restrict__ tot) {
    __shared__ unsigned warp_tot[8];
    __shared__ unsigned carry_s;
    const int dg = blockIdx.x, s = blockIdx.y;
    unsigned* h = hist + ((size_t)s * RS_RADIX + dg) * ntiles;
    if (threadIdx.x == 0) carry_s = 0;
    __syncthreads();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int base = 0; base < ntiles; base += 256) {
        const int i = base + threadIdx.x;
        const unsigned v = i < ntiles ? h[i] : 0u;
        unsigned x = v;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const unsigned y = __shfl_up_sync(0xffffffffu, x, o);
            if (lane >= o) x += y;
        }
        if (lane == 31) warp_tot[warp] = x;
        __syncthreads();
        unsigned before = carry_s + (x - v);
        for (int w = 0; w < warp; ++w) before += warp_tot[w];
        if (i < ntiles) h[i] = before;
        __syncthreads();
        if (threadIdx.x == 255) carry_s = before + v;
        __syncthreads();
    }
    if (threadIdx.x == 0) tot[(size_t)s * RS_RADIX + dg] = carry_s;
}

__global__ void __launch_bounds__(RS_RADIX)
rs_scan_digits_kernel(unsigned* __restrict__ tot) {
    __shared__ unsigned warp_tot[RS_RADIX / 32];
    unsigned* t = tot + (size_t)blockIdx.x * RS_RADIX;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const unsigned v = t[threadIdx.x];
    unsigned x = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const unsigned y = __shfl_up_sync(0xffffffffu, x, o);
        if (lane >= o) x += y;
    }
    if (lane == 31) warp_tot[warp] = x;
    __syncthreads();
    unsigned before = x - v;
    for (int w = 0; w < warp; ++w) before += warp_tot[w];
    t[threadIdx.x] = before;
}

// stable scatter of one pass: keys (and a 32-bit payload) from `in` to `out` inside their segment
__global__ void __launch_bounds__(RS_THREADS)
rs_scatter_kernel(const unsigned long long* __restrict__ keys_in, const unsigned* __restrict__ vals_in,
                  unsigned long long* __restrict__ keys_out, unsigned* __restrict__ vals_out, RsSegments sg, int shift,
                  int ntiles, const unsigned* __restrict__ hist, const unsigned* __restrict__ tot) {
    __shared__ unsigned wh[RS_WARPS][RS_RADIX];            // per-warp digit counts, then the warp's base inside the tile
    const int s = blockIdx.y, tile = blockIdx.x;
    long long off, len;
    rs_segment(sg, s, off, len);
    const long long t0 = (long long)tile * RS_TILE;
    if (t0 >= len) return;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int d = threadIdx.x; d < RS_WARPS * RS_RADIX; d += RS_THREADS) (&wh[0][0])[d] = 0;
    __syncthreads();
    // warp w owns keys [t0 + w * 32 * ITEMS, +32 * ITEMS), walked 32 at a time in order
    const long long w0 = t0 + (long long)warp * 32 * RS_ITEMS;
    unsigned long long key[RS_ITEMS];
    unsigned loc[RS_ITEMS];                                 // rank among the equal digits of the warp's slice so far
    const unsigned lt = (1u << lane) - 1u;
#pragma unroll
    for (int it = 0; it < RS_ITEMS; ++it) {
        const long long i = w0 + it * 32 + lane;
        const bool live = i < len;
        key[it] = live ? keys_in[off + i] : ~0ULL;
        const unsigned dg = (unsigned)(key[it] >> shift) & (RS_RADIX - 1);
        const unsigned act = __ballot_sync(0xffffffffu, live);
        unsigned peers = __match_any_sync(0xffffffffu, live ? dg : RS_RADIX + lane) & act;
        const unsigned base = wh[warp][dg];
        loc[it] = base + __popc(peers & lt);
        __syncwarp();
        if (live && (peers & lt) == 0) wh[warp][dg] = base + __popc(peers);      // lowest lane of each digit group
        __syncwarp();
    }
    __syncthreads();
    // exclusive scan over the warps per digit + the tile's global offset of that digit
    {
        const int dg = threadIdx.x;                         // RS_THREADS == RS_RADIX
        unsigned run = tot[(size_t)s * RS_RADIX + dg] + hist[((size_t)s * RS_RADIX + dg) * ntiles + tile];
#pragma unroll
        for (int w = 0; w < RS_WARPS; ++w) {
            const unsigned c = wh[w][dg];
            wh[w][dg] = run;
            run += c;
        }
    }
    __syncthreads();
#pragma unroll
    for (int it = 0; it < RS_ITEMS; ++it) {
        const long long i = w0 + it * 32 + lane;
        if (i < len) {
            const unsigned dg = (unsigned)(key[it] >> shift) & (RS_RADIX - 1);
            const long long dst = off + wh[warp][dg] + loc[it];
            keys_out[dst] = key[it];
            if (vals_in != nullptr) vals_out[dst] = vals_in[off + i];
        }
    }
}
