// Issue-rate micro-benchmark: warp-instructions per clock per SM of MUFU.EX2, F2FP (cvt.rn.f16x2.f32), FFMA, FFMA2 and
// of MUFU + F2FP interleaved (do they share the XU pipe?).   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o xu_rate xu_rate.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

template <int MODE>
__global__ void __launch_bounds__(256) k(float* out, long long* clk, int reps) {
    float a[8], b[8];
    unsigned h[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) { a[i] = threadIdx.x * 1e-3f + i; b[i] = -a[i]; h[i] = 0; }
    __syncthreads();
    const long long t0 = clock64();
    for (int r = 0; r < reps; ++r) {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            if (MODE == 0) asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(a[i]));
            if (MODE == 1) asm volatile("cvt.rn.f16x2.f32 %0, %1, %2;" : "=r"(h[i]) : "f"(a[i]), "f"(b[i]));
            if (MODE == 2) asm volatile("fma.rn.ftz.f32 %0, %0, %1, %1;" : "+f"(a[i]) : "f"(b[i]));
            if (MODE == 3) {
                asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(a[i]));
                asm volatile("cvt.rn.f16x2.f32 %0, %1, %2;" : "=r"(h[i]) : "f"(b[i]), "f"(b[i]));
            }
            if (MODE == 4) {
                unsigned long long x, y;
                asm volatile("mov.b64 %0, {%1, %2};" : "=l"(x) : "f"(a[i]), "f"(b[i]));
                asm volatile("fma.rn.ftz.f32x2 %0, %1, %1, %1;" : "=l"(y) : "l"(x));
                asm volatile("mov.b64 {%0, %1}, %2;" : "=f"(a[i]), "=f"(b[i]) : "l"(y));
            }
            if (MODE == 6) asm volatile("{.reg .b16 lo, hi; mov.b32 {lo, hi}, %0; ex2.approx.f16 lo, lo; mov.b32 %0, {lo, hi};}" : "+r"(h[i]));
            if (MODE == 7) asm volatile("ex2.approx.f16x2 %0, %0;" : "+r"(h[i]));
            if (MODE == 5) asm volatile("max.ftz.f32 %0, %0, %1, %2;" : "+f"(a[i]) : "f"(b[i]), "f"(b[(i + 1) & 7]));
        }
    }
    const long long t1 = clock64();
    float s = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) s += a[i] + b[i] + __uint_as_float(h[i]);
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0) clk[blockIdx.x] = t1 - t0;
}

template <int MODE>
void run(const char* name, int per_iter) {
    float* out; long long* clk;
    cudaMalloc(&out, 148 * 256 * 4); cudaMalloc(&clk, 148 * 8);
    const int reps = 4000;
    k<MODE><<<148, 256>>>(out, clk, 10);
    k<MODE><<<148, 256>>>(out, clk, reps);
    cudaError_t e = cudaDeviceSynchronize();
    long long h[148];
    cudaMemcpy(h, clk, sizeof(h), cudaMemcpyDeviceToHost);
    const double winst = 8.0 * reps * 8 * per_iter;        // warps x reps x unroll x instr
    printf("%-34s %s  %.3f warp-instr/clk/SM  (= %.1f lanes/clk/SM, %.2f clk per warp-instr per scheduler)\n", name,
           cudaGetErrorString(e), winst / h[0], 32.0 * winst / h[0], 4.0 * h[0] / winst);
    cudaFree(out); cudaFree(clk);
}

int main() {
    run<0>("MUFU.EX2", 1);
    run<1>("F2FP (cvt.rn.f16x2.f32)", 1);
    run<2>("FFMA", 1);
    run<3>("MUFU.EX2 + F2FP interleaved", 2);
    run<4>("FFMA2 (fma.f32x2)", 1);
    run<5>("FMNMX3 (3-input max)", 1);
    run<6>("MUFU.EX2.F16 (scalar half)", 1);
    run<7>("ex2.approx.f16x2 (2 x MUFU.EX2.F16 + PRMT per instruction)", 1);
    return 0;
}
