// Ceiling of the convolution inner loops at the occupancy the real kernels run with (2 CTAs x 256 threads per SM):
// the loops of kh_kernel (conv_slide, sliding register windows) and kv_kernel (kv_group, shared pair sums) on a
// shared-memory tile, no HBM traffic, no epilogue.  Prints FP64 warp-instruction rates to compare with the pipe roof
// measured by tools/fp64_peak.cu.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I../mustache_b200/csrc -o conv_peak conv_peak.cu
#include <cstdio>
#include <cstring>
#include <vector>
#include "mb_kernels.cuh"

template <int K>
__device__ __forceinline__ void conv_pipe(const double* __restrict__ ctr, const int stride, const int R,
                                          const double* __restrict__ tp, double (&acc)[K]) {
    double pl[K], pr[K], t[K], pd[K];
    const double w0 = tp[0];
#pragma unroll
    for (int k = 0; k < K; ++k) { acc[k] = __dmul_rn(ctr[k * stride], w0); t[k] = 0.0; pd[k] = 0.0; }
    double wprev = 0.0;
    const int u0 = (K - (R & (K - 1))) & (K - 1);
    int j = R + u0;
#pragma unroll
    for (int p = 0; p < K; ++p) {
        pl[p] = ctr[(p - j + (p < u0 ? K : 0)) * stride];
        pr[p] = ctr[(p + j - (p > K - 1 - u0 ? K : 0)) * stride];
    }
#define PTAP(u)                                                                                                \
    {                                                                                                          \
        _Pragma("unroll") for (int k = 0; k < K; ++k) acc[k] = __dadd_rn(acc[k], pd[k]);                       \
        _Pragma("unroll") for (int k = 0; k < K; ++k) pd[k] = __dmul_rn(t[k], wprev);                          \
        _Pragma("unroll") for (int k = 0; k < K; ++k) t[k] = __dadd_rn(pl[(k + u) % K], pr[(k - u + K) % K]);  \
        wprev = tp[j - u];                                                                                     \
        pl[u % K] = ctr[(u + K - j) * stride];                                                                 \
        pr[(K - 1 - u) % K] = ctr[(j - u - 1) * stride];                                                       \
    }
    switch (u0) {
        case 0: do { PTAP(0)
        case 1: PTAP(1)
        case 2: PTAP(2)
        case 3: PTAP(3)
        case 4: PTAP(4)
        case 5: PTAP(5)
        case 6: PTAP(6)
        case 7: PTAP(7)
                j -= K; } while (j > 0);
    }
#undef PTAP
#pragma unroll
    for (int k = 0; k < K; ++k) acc[k] = __dadd_rn(acc[k], pd[k]);
#pragma unroll
    for (int k = 0; k < K; ++k) acc[k] = __dadd_rn(acc[k], __dmul_rn(t[k], wprev));
}

struct Taps { double w[8][64]; int R[8]; };

template <int MODE, int CTAS>
__global__ void __launch_bounds__(256, CTAS) k(const __grid_constant__ Taps tp, const __grid_constant__ KvPlan plan, double* out,
                                                int reps, int nr, int pad_smem) {
    extern __shared__ double sm[];
    for (int e = threadIdx.x; e < pad_smem; e += 256) sm[e] = 1.0 + 1e-6 * e;
    __syncthreads();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    double tot = 0.0;
    if (MODE == 0) {            // kh layout: lane = row (pitch 178 + stagger), 8 consecutive columns per thread
        double acc[8];
        for (int r = 0; r < reps; ++r)
            for (int s = 0; s < nr; ++s) {
                conv_slide<8, 1, false>(sm + lane * 178 + ((lane >> 3) & 1) + warp * 8 + 56, tp.R[s], tp.w[s], acc);
                tot += ((acc[0] + acc[1]) + (acc[2] + acc[3])) + ((acc[4] + acc[5]) + (acc[6] + acc[7]));
            }
    } else if (MODE == 2) {     // same, software-pipelined taps
        double acc[8];
        for (int r = 0; r < reps; ++r)
            for (int s = 0; s < nr; ++s) {
                conv_pipe<8>(sm + lane * 178 + ((lane >> 3) & 1) + warp * 8 + 56, 1, tp.R[s], tp.w[s], acc);
                tot += ((acc[0] + acc[1]) + (acc[2] + acc[3])) + ((acc[4] + acc[5]) + (acc[6] + acc[7]));
            }
    } else if (MODE == 1) {     // kv layout: lane = column, KV_K consecutive rows per thread
        double* sink = sm + 6144;      // [8 warps][40 rows][32]
        for (int r = 0; r < reps; ++r)
            for (int gi = 0; gi < plan.n_groups; ++gi) {
                const KvGroup& gr = plan.grp[gi];
                const double* ctr = sm + (warp * KV_K + 60) * KV_TW + lane;
                double* vrow = sink + warp * (KV_K * 32) + lane;
                const unsigned vm = pad_smem ? 0xfu : 0x7u;
                switch (gr.n) {
                    case 1: kv_group<1, false>(ctr, plan.tapsT + gr.tap_off, gr.rmax, gr.step, vrow, 8 * KV_K * 32, 32, vm); break;
                    case 2: kv_group<2, false>(ctr, plan.tapsT + gr.tap_off, gr.rmax, gr.step, vrow, 8 * KV_K * 32, 32, vm); break;
                    case 3: kv_group<3, false>(ctr, plan.tapsT + gr.tap_off, gr.rmax, gr.step, vrow, 8 * KV_K * 32, 32, vm); break;
                    case 4: kv_group<4, false>(ctr, plan.tapsT + gr.tap_off, gr.rmax, gr.step, vrow, 8 * KV_K * 32, 32, vm); break;
                    default: kv_group<5, false>(ctr, plan.tapsT + gr.tap_off, gr.rmax, gr.step, vrow, 8 * KV_K * 32, 32, vm); break;
                }
            }
        tot = sink[threadIdx.x];
    }
    if (tot == 123.456) out[0] = tot;
}

template <int MODE, int CTAS>
void run(const char* name, const Taps& tp, const KvPlan& plan, int nr, double fp64_per_thread_rep, size_t smem) {
    double* d;
    cudaMalloc(&d, 8);
    int sms = 148;
    cudaFuncSetAttribute(k<MODE, CTAS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    const int reps = 40, blocks = sms * CTAS * 4;
    printf("");
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    k<MODE, CTAS><<<blocks, 256, smem>>>(tp, plan, d, 2, nr, (int)(smem / 8));
    cudaError_t st = cudaDeviceSynchronize();
    if (st != cudaSuccess) { printf("{\"op\": \"%s\", \"error\": \"%s\"}\n", name, cudaGetErrorString(st)); return; }
    float best = 1e30f;
    for (int r = 0; r < 3; ++r) {
        cudaEventRecord(e0);
        k<MODE, CTAS><<<blocks, 256, smem>>>(tp, plan, d, reps, nr, (int)(smem / 8));
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        float ms;
        cudaEventElapsedTime(&ms, e0, e1);
        if (ms < best) best = ms;
    }
    const double ops = (double)blocks * 256 * reps * fp64_per_thread_rep;
    printf("{\"op\": \"%s\", \"ctas_per_sm\": %d, \"fp64_instr_per_s\": %.4e, \"frac_of_1.849e13\": %.3f, \"ms\": %.3f}\n", name, CTAS,
           ops / (best * 1e-3), ops / (best * 1e-3) / 1.849e13, best);
    cudaFree(d);
}

int main() {
    Taps tp;
    memset(&tp, 0, sizeof(tp));
    KvPlan plan;
    memset(&plan, 0, sizeof(plan));
    for (int s = 0; s < 8; ++s)
        for (int j = 0; j < 64; ++j) tp.w[s][j] = 1.0 / (1 + j + s);
    const size_t smem = 100 * 1024;
    {   // kh loop, large radii only / small radii only / the 4-octave mix
        int Rbig[8] = {55, 52, 48, 45, 42, 39, 36, 34};
        memcpy(tp.R, Rbig, sizeof(Rbig));
        double f = 0; for (int s = 0; s < 8; ++s) f += 8.0 * (3 * tp.R[s] + 1);
        run<0, 2>("kh conv_slide R=34..55", tp, plan, 8, f, smem);
        run<2, 2>("kh conv_pipe R=34..55", tp, plan, 8, f, smem);
        run<0, 3>("kh conv_slide R=34..55", tp, plan, 8, f, 70 * 1024);
        run<2, 3>("kh conv_pipe R=34..55", tp, plan, 8, f, 70 * 1024);
        run<0, 1>("kh conv_slide R=34..55", tp, plan, 8, f, 200 * 1024);
        int Rsm[8] = {4, 4, 5, 5, 6, 6, 7, 7};
        memcpy(tp.R, Rsm, sizeof(Rsm));
        f = 0; for (int s = 0; s < 8; ++s) f += 8.0 * (3 * tp.R[s] + 1);
        run<0, 2>("kh conv_slide R=4..7", tp, plan, 8, f, smem);
        run<2, 2>("kh conv_pipe R=4..7", tp, plan, 8, f, smem);
        run<2, 3>("kh conv_pipe R=4..7", tp, plan, 8, f, 70 * 1024);
        int Rmid[8] = {8, 9, 10, 12, 14, 16, 19, 23};
        memcpy(tp.R, Rmid, sizeof(Rmid));
        f = 0; for (int s = 0; s < 8; ++s) f += 8.0 * (3 * tp.R[s] + 1);
        run<0, 2>("kh conv_slide R=8..23", tp, plan, 8, f, smem);
        run<2, 2>("kh conv_pipe R=8..23", tp, plan, 8, f, smem);
        run<2, 3>("kh conv_pipe R=8..23", tp, plan, 8, f, 70 * 1024);
    }
    {   // kv loop: groups of 4 at rmax 55 / 12 / 4
        const int rm[3] = {55, 12, 4};
        for (int t = 0; t < 3; ++t)
            for (int n = 2; n <= 5; n += 1) {
                memset(&plan, 0, sizeof(plan));
                plan.n_groups = 4;
                for (int gi = 0; gi < 4; ++gi) {
                    plan.grp[gi].n = n;
                    plan.grp[gi].rmax = rm[t];
                    plan.grp[gi].tap_off = 0;
                    for (int s = 0; s < n; ++s) plan.grp[gi].step[s] = s;
                }
                for (int j = 0; j < (rm[t] + 1) * n; ++j) plan.tapsT[j] = 1.0 / (1 + j);
                char name[64];
                snprintf(name, sizeof(name), "kv_group<%d> rmax=%d", n, rm[t]);
                const double f = 4.0 * KV_K * (rm[t] * (2 * n + 1) + n);
                run<1, 2>(name, tp, plan, 0, f, smem);
            }
    }
    return 0;
}
