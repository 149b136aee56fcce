// FP64 pipe micro-benchmark for the second roofline (MEASURED_PEAKS.json has no fp64 entry).
// Measures sustained DADD / DMUL / DFMA / "DADD+DMUL+DADD tap" warp-instruction rates on all SMs.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o fp64_peak fp64_peak.cu && ./fp64_peak
#include <cstdio>
#include <cuda_runtime.h>

template <int MODE>
__global__ void __launch_bounds__(256) k(double* out, double a, double b, int iters) {
    double x[16];
    unsigned y[8];
#pragma unroll
    for (int i = 0; i < 16; ++i) x[i] = a + threadIdx.x * 1e-9 + i;
#pragma unroll
    for (int i = 0; i < 8; ++i) y[i] = threadIdx.x * 2654435761u + i;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 16; ++i) {
            if (MODE == 0) x[i] = __dadd_rn(x[i], b);
            if (MODE == 1) x[i] = __dmul_rn(x[i], b);
            if (MODE == 2) x[i] = __fma_rn(x[i], b, a);
            if (MODE == 3) x[i] = __dadd_rn(x[i], __dmul_rn(__dadd_rn(x[(i + 1) & 15], x[(i + 5) & 15]), b));   // one folded tap
            if (MODE == 4) x[i] = (x[i] > x[(i + 1) & 15]) ? x[i] : x[(i + 1) & 15];
            if (MODE >= 5) {                 // issue-slot model: 16 DADD + MIX integer instructions per trip
                x[i] = __dadd_rn(x[i], b);
                constexpr int MIX = (MODE == 5) ? 4 : (MODE == 6) ? 8 : 16;
                if (i < MIX) y[i & 7] = (y[i & 7] ^ y[(i + 1) & 7]) + y[(i + 3) & 7];      // LOP3 + IADD
            }
        }
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < 16; ++i) s += x[i];
    if (MODE >= 5) {
#pragma unroll
        for (int i = 0; i < 8; ++i) s += y[i];
    }
    if (s == 123.456) out[0] = s;
}

template <int MODE>
void run(const char* name, int ops_per_elem) {
    double* d;
    cudaMalloc(&d, 8);
    int dev, sms;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    const int iters = 4096, blocks = sms * 8;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    k<MODE><<<blocks, 256>>>(d, 1.0, 1.0000001, 64);
    cudaDeviceSynchronize();
    float best = 1e30f;
    for (int r = 0; r < 5; ++r) {
        cudaEventRecord(e0);
        k<MODE><<<blocks, 256>>>(d, 1.0, 1.0000001, iters);
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        float ms;
        cudaEventElapsedTime(&ms, e0, e1);
        if (ms < best) best = ms;
    }
    const double ops = (double)blocks * 256 * iters * 16 * ops_per_elem;
    printf("{\"op\": \"%s\", \"fp64_instr_per_s\": %.4e, \"ms\": %.3f, \"sms\": %d}\n", name, ops / (best * 1e-3), best, sms);
    cudaFree(d);
}

int main() {
    run<0>("dadd", 1);
    run<1>("dmul", 1);
    run<2>("dfma", 1);
    run<3>("tap(dadd,dmul,dadd)", 3);
    run<4>("dsetp+sel(max)", 1);
    run<5>("16 dadd + 8 int instr per trip (fp64 rate)", 1);
    run<6>("16 dadd + 16 int instr per trip (fp64 rate)", 1);
    run<7>("16 dadd + 32 int instr per trip (fp64 rate)", 1);
    return 0;
}
