// Issue-rate probe for the packed FP32 instructions of sm_100a (FFMA2 / FADD2 / FMUL2, PTX *.f32x2) next to their scalar
// forms, alone and interleaved with integer ALU work.  Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o
// /tmp/f32x2_probe tools/probes/f32x2_probe.cu ; run on the GPU box.  Prints warp-instructions per clock per SM.
#include <cuda_runtime.h>
#include <cstdio>
typedef unsigned long long u64;
__device__ __forceinline__ u64 pk(float lo, float hi) { u64 r; asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi)); return r; }
__device__ __forceinline__ void upk(u64 v, float& lo, float& hi) { asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v)); }
__device__ __forceinline__ u64 fma2(u64 a, u64 b, u64 c) { u64 d; asm volatile("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c)); return d; }
__device__ __forceinline__ u64 add2(u64 a, u64 b) { u64 d; asm volatile("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b)); return d; }
__device__ __forceinline__ float fma1(float a, float b, float c) { float d; asm volatile("fma.rn.f32 %0, %1, %2, %3;" : "=f"(d) : "f"(a), "f"(b), "f"(c)); return d; }
__device__ __forceinline__ float add1(float a, float b) { float d; asm volatile("add.rn.f32 %0, %1, %2;" : "=f"(d) : "f"(a), "f"(b)); return d; }
__device__ __forceinline__ unsigned iadd(unsigned a, unsigned b) { unsigned d; asm volatile("add.u32 %0, %1, %2;" : "=r"(d) : "r"(a), "r"(b)); return d; }
__device__ __forceinline__ float fmin3(float a, float b, float c) { return fminf(a, fminf(b, c)); }

#define CH 8
template <int MODE>
__global__ void __launch_bounds__(256) probe(int iters, float seed, float* out) {
    float s = seed + threadIdx.x * 1e-6f;
    float a[CH];
    u64 A[CH];
    unsigned n[CH];
#pragma unroll
    for (int k = 0; k < CH; ++k) { a[k] = s + k; A[k] = pk(s + k, s - k); n[k] = threadIdx.x + k; }
    const float m = 0.999f, c = 1e-3f;
    const u64 M = pk(m, m), C = pk(c, c);
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int k = 0; k < CH; ++k) {
            if (MODE == 0) a[k] = fma1(a[k], m, c);                                  // FFMA
            if (MODE == 1) A[k] = fma2(A[k], M, C);                                  // FFMA2
            if (MODE == 2) { a[k] = fma1(a[k], m, c); n[k] = iadd(n[k], 0x9e3779b9u + k); }   // FFMA + IADD
            if (MODE == 3) { A[k] = fma2(A[k], M, C); n[k] = iadd(n[k], 0x9e3779b9u + k); }   // FFMA2 + IADD
            if (MODE == 4) a[k] = add1(a[k], c);                                     // FADD
            if (MODE == 5) A[k] = add2(A[k], C);                                     // FADD2
            if (MODE == 6) { A[k] = fma2(A[k], M, C); a[k] = fma1(a[k], m, c); }      // FFMA2 + FFMA
            if (MODE == 7) { A[k] = fma2(A[k], M, C); a[k] = fminf(a[k], m + k); }    // FFMA2 + FMNMX
            if (MODE == 8) { a[k] = fma1(a[k], m, c); a[(k + 1) % CH] = fminf(a[(k + 1) % CH], 1e30f - k); }  // FFMA + FMNMX
            if (MODE == 9) n[k] = iadd(n[k], 0x9e3779b9u + k);                       // IADD only
        }
    }
    float r = 0.f;
#pragma unroll
    for (int k = 0; k < CH; ++k) { float lo, hi; upk(A[k], lo, hi); r += a[k] + lo + hi + (float)n[k]; }
    if (r == 123.456f) out[0] = r;
}

template <int MODE>
static void run(const char* name, int per_iter_instr, int flops_per_iter) {
    int dev_sms = 0; cudaDeviceGetAttribute(&dev_sms, cudaDevAttrMultiProcessorCount, 0);
    int clk_khz = 0; cudaDeviceGetAttribute(&clk_khz, cudaDevAttrClockRate, 0);
    float* out; cudaMalloc(&out, 4);
    const int iters = 20000, blocks = dev_sms * 8, threads = 256;
    probe<MODE><<<blocks, threads>>>(iters / 10, 1.f, out);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    cudaEventRecord(e0);
    probe<MODE><<<blocks, threads>>>(iters, 1.f, out);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    double winst = (double)blocks * (threads / 32) * iters * CH * per_iter_instr;
    double clocks = ms * 1e-3 * clk_khz * 1e3;
    printf("%-14s %8.3f ms  %6.2f warp-inst/clk/SM (at %d MHz)  %7.2f TFLOP/s\n", name, ms, winst / clocks / dev_sms,
           clk_khz / 1000, (double)blocks * threads * iters * CH * flops_per_iter / (ms * 1e-3) / 1e12);
    cudaFree(out);
}

int main() {
    run<0>("FFMA", 1, 2);
    run<1>("FFMA2", 1, 4);
    run<2>("FFMA+IADD", 2, 2);
    run<3>("FFMA2+IADD", 2, 4);
    run<4>("FADD", 1, 1);
    run<5>("FADD2", 1, 2);
    run<6>("FFMA2+FFMA", 2, 6);
    run<7>("FFMA2+FMNMX", 2, 4);
    run<8>("FFMA+FMNMX", 2, 2);
    run<9>("IADD", 1, 0);
    return 0;
}
