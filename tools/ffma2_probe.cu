// ffma2_probe.cu -- issue rate of the packed FP32 instructions of sm_100 (FFMA2 / FADD2 / FMUL2 = fma|add|mul.rn.f32x2)
// against scalar FFMA, alone and mixed with other work (instrumentation for DESIGN.md section 4.4).
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o /tmp/ffma2_probe tools/ffma2_probe.cu && /tmp/ffma2_probe
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e_)); exit(1); } } while (0)

__device__ __forceinline__ unsigned long long fma2(unsigned long long a, unsigned long long b, unsigned long long c) {
    unsigned long long r;
    asm volatile("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c));
    return r;
}
__device__ __forceinline__ unsigned long long pack(float x, float y) {
    unsigned long long r;
    asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(x), "f"(y));
    return r;
}

// MODE 0: 8 scalar FFMA chains; 1: 8 FFMA2 chains (16 FMAs per round); 2: 4 FFMA2 + 4 scalar FFMA; 3: 8 FFMA2 + 8 MUFU-free integer ops
template <int MODE>
__global__ void __launch_bounds__(256) probe(float *out, int iters, float a, float b) {
    float s[8];
    unsigned long long p[8];
    const unsigned long long pa = pack(a, a), pb = pack(b, b);
    for (int k = 0; k < 8; k++) { s[k] = threadIdx.x + k; p[k] = pack(s[k], s[k] + 0.5f); }
    int acc = threadIdx.x;
    for (int i = 0; i < iters; i++) {
#pragma unroll
        for (int r = 0; r < 8; r++) {
            if (MODE == 0) {
#pragma unroll
                for (int k = 0; k < 8; k++) s[k] = fmaf(s[k], a, b);
            } else if (MODE == 1) {
#pragma unroll
                for (int k = 0; k < 8; k++) p[k] = fma2(p[k], pa, pb);
            } else if (MODE == 2) {
#pragma unroll
                for (int k = 0; k < 4; k++) { p[k] = fma2(p[k], pa, pb); s[k] = fmaf(s[k], a, b); }
            } else {
#pragma unroll
                for (int k = 0; k < 8; k++) { p[k] = fma2(p[k], pa, pb); acc = (acc ^ (acc << 1)) + k; }
            }
        }
    }
    float t = 0;
    for (int k = 0; k < 8; k++) t += s[k] + __uint_as_float((unsigned)p[k]) + __uint_as_float((unsigned)(p[k] >> 32));
    if (t == 12345.678f || acc == 0x7fffffff) out[0] = t;
}

int main() {
    int sms = 148, khz = 1965000;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, 0);
    float *out; CK(cudaMalloc(&out, 256));
    cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    const int blocks = sms * 8, threads = 256, iters = 4096;
    const char *names[4] = {"8 x FFMA", "8 x FFMA2", "4 x FFMA2 + 4 x FFMA", "8 x FFMA2 + 16 integer ops"};
    const double fmas_per_round[4] = {8, 16, 12, 16}, inst_per_round[4] = {8, 8, 8, 24};
    for (int m = 0; m < 4; m++) {
        float best = 1e30f;
        for (int rep = 0; rep < 3; rep++) {
            CK(cudaEventRecord(e0));
            if (m == 0) probe<0><<<blocks, threads>>>(out, iters, 1.0000001f, 1e-9f);
            if (m == 1) probe<1><<<blocks, threads>>>(out, iters, 1.0000001f, 1e-9f);
            if (m == 2) probe<2><<<blocks, threads>>>(out, iters, 1.0000001f, 1e-9f);
            if (m == 3) probe<3><<<blocks, threads>>>(out, iters, 1.0000001f, 1e-9f);
            CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
            float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
            if (ms < best) best = ms;
        }
        const double rounds = (double)blocks * threads * iters * 8;
        const double tf = rounds * fmas_per_round[m] * 2 / (best * 1e-3) / 1e12;
        const double ipc = rounds / 32 * inst_per_round[m] / (best * 1e-3) / ((double)sms * khz * 1e3);   // warp instructions per clock per SM
        printf("%-28s %7.3f ms  %6.1f TFLOP/s  %.2f warp instructions / clk / SM\n", names[m], best, tf, ipc);
    }
    return 0;
}
