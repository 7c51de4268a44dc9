// TMEM read bandwidth micro-benchmark: how many bytes per clock can tcgen05.ld (32x32b.x32) deliver per SM?
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tmem_bw tmem_bw.cu ; run: ./tmem_bw
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t* r) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
          "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
          "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr));
}

template <int WARPS>
__global__ void __launch_bounds__(WARPS * 32) k(uint32_t* out, long long* clk, int reps) {
    __shared__ uint32_t slot;
    const int warp = threadIdx.x >> 5;
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"((uint32_t)__cvta_generic_to_shared(&slot)));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    asm volatile("tcgen05.fence::before_thread_sync;");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;");
    const uint32_t base = slot + ((uint32_t)((warp & 3) * 32) << 16) + (warp >> 2) * 128;
    uint32_t acc = 0;
    __syncthreads();
    const long long t0 = clock64();
    for (int i = 0; i < reps; ++i) {
        uint32_t v[128];
        tmem_ld32(base, v);
        tmem_ld32(base + 32, v + 32);
        tmem_ld32(base + 64, v + 64);
        tmem_ld32(base + 96, v + 96);
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
        for (int j = 0; j < 128; ++j) acc ^= v[j];
    }
    const long long t1 = clock64();
    __syncthreads();
    if (threadIdx.x == 0) clk[blockIdx.x] = t1 - t0;
    out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(slot));
}

template <int WARPS>
void run(const char* name) {
    uint32_t* out;
    long long* clk;
    cudaMalloc(&out, 148 * 512 * 4);
    cudaMalloc(&clk, 148 * 8);
    const int reps = 2000;
    k<WARPS><<<148, WARPS * 32>>>(out, clk, 10);
    k<WARPS><<<148, WARPS * 32>>>(out, clk, reps);
    cudaError_t e = cudaDeviceSynchronize();
    long long h[148];
    cudaMemcpy(h, clk, sizeof(h), cudaMemcpyDeviceToHost);
    const double bytes = (double)WARPS * 32 * 128 * 4 * reps;            // per SM
    printf("%-28s %s  %lld clk for %d x (%d warps x 16 KB)  -> %.1f B/clk/SM, %.0f clk per 64 KB tile\n", name,
           cudaGetErrorString(e), h[0], reps, WARPS, bytes / h[0], 65536.0 * h[0] / bytes);
    cudaFree(out);
    cudaFree(clk);
}

int main() {
    run<1>("1 warp  (1 SMSP)");
    run<4>("4 warps (1 per SMSP)");
    run<8>("8 warps (2 per SMSP)");
    return 0;
}
