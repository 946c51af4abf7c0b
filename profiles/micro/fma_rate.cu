// Micro-benchmark: FP32 FMA issue rate on B200 for the operand forms the blur kernel can use.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o fma_rate fma_rate.cu && ./fma_rate
#include <cstdio>
#include <cuda_runtime.h>
typedef unsigned long long f32x2;
__device__ __forceinline__ f32x2 pack2(float lo, float hi) { f32x2 r; asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi)); return r; }
__device__ __forceinline__ void unpack2(f32x2 v, float& lo, float& hi) { asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v)); }
__device__ __forceinline__ f32x2 fma2(f32x2 a, f32x2 b, f32x2 c) { f32x2 d; asm volatile("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c)); return d; }
struct W { float w[32]; };
constexpr int ITERS = 4096;
// 8 independent chains, weights in registers (3-register FFMA)
__global__ void k_ffma_reg(float* out, float a, float b) {
    float acc[8]; for (int k = 0; k < 8; k++) acc[k] = threadIdx.x * 0.001f + k;
    float w0 = a, w1 = b;
    for (int it = 0; it < ITERS; it++) {
#pragma unroll
        for (int k = 0; k < 8; k++) acc[k] = fmaf(w0, acc[k], w1);
#pragma unroll
        for (int k = 0; k < 8; k++) acc[k] = fmaf(w1, acc[k], w0);
    }
    float s = 0; for (int k = 0; k < 8; k++) s += acc[k];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
// weights from the constant bank (kernel parameter), like the blur kernel: FFMA R, R, c[][], R
__global__ void k_ffma_const(float* out, const __grid_constant__ W w, float x0) {
    float acc[8]; for (int k = 0; k < 8; k++) acc[k] = 0.f;
    float v[8]; for (int k = 0; k < 8; k++) v[k] = x0 + threadIdx.x + k;
    for (int it = 0; it < ITERS / 2; it++) {
#pragma unroll
        for (int i = 0; i < 4; i++) {
#pragma unroll
            for (int k = 0; k < 8; k++) acc[k] = fmaf(w.w[i], v[k], acc[k]);
        }
    }
    float s = 0; for (int k = 0; k < 8; k++) s += acc[k];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
// packed FFMA2, 8 independent chains of pairs, scalar-broadcast weight
__global__ void k_ffma2(float* out, const __grid_constant__ W w, float x0) {
    f32x2 acc[8]; for (int k = 0; k < 8; k++) acc[k] = pack2(0.f, 0.f);
    f32x2 v[8]; for (int k = 0; k < 8; k++) v[k] = pack2(x0 + threadIdx.x + k, x0 - k);
    for (int it = 0; it < ITERS / 2; it++) {
#pragma unroll
        for (int i = 0; i < 4; i++) {
            const f32x2 ww = pack2(w.w[i], w.w[i]);
#pragma unroll
            for (int k = 0; k < 8; k++) acc[k] = fma2(ww, v[k], acc[k]);
        }
    }
    float s = 0; for (int k = 0; k < 8; k++) { float a, b; unpack2(acc[k], a, b); s += a + b; }
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
template <class F> float timeit(F f) {
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    f(); cudaDeviceSynchronize();
    cudaEventRecord(a); for (int i = 0; i < 5; i++) f(); cudaEventRecord(b); cudaEventSynchronize(b);
    float ms; cudaEventElapsedTime(&ms, a, b); return ms / 5;
}
int main() {
    cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
    const int sms = p.multiProcessorCount, blocks = sms * 8, threads = 256;
    float* out; cudaMalloc(&out, blocks * threads * sizeof(float));
    W w; for (int i = 0; i < 32; i++) w.w[i] = 0.5f + i * 0.01f;
    const double clk = p.clockRate * 1e3;
    auto report = [&](const char* name, float ms, double fmasPerThread) {
        const double total = fmasPerThread * blocks * threads;
        printf("%-28s %8.3f ms  %7.2f TFMA/s  %6.1f FMA/clk/SM (at %.0f MHz nominal)\n", name, ms,
               total / ms / 1e9, total / (ms * 1e-3) / clk / sms, clk / 1e6);
    };
    report("FFMA 3-register", timeit([&] { k_ffma_reg<<<blocks, threads>>>(out, 0.999f, 0.001f); }), 16.0 * ITERS);
    report("FFMA const-bank operand", timeit([&] { k_ffma_const<<<blocks, threads>>>(out, w, 1.f); }), 32.0 * ITERS / 2);
    report("FFMA2 (f32x2) broadcast w", timeit([&] { k_ffma2<<<blocks, threads>>>(out, w, 1.f); }), 2 * 32.0 * ITERS / 2);
    printf("err=%s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
