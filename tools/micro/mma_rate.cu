// tcgen05.mma issue-to-retire cost for the shapes the kernels use: cycles per MMA (K = 16) for dependent chains on one
// accumulator, N = 64 / 128 / 256, A from shared memory (SS) or tensor memory (TS), and for two interleaved accumulators.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I../../star_b200/csrc -o mma_rate mma_rate.cu
#include <cstdio>
#include "common.cuh"
using namespace star;

template <int N, int TS, int TWO_ACC>
__global__ void __launch_bounds__(128) k(long long* clk, int reps, int chain) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    __shared__ uint64_t bar;
    __shared__ uint32_t slot;
    for (int i = threadIdx.x; i < (16384 + 32768) / 4; i += 128) reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u;   // fp16 1.0
    if (threadIdx.x == 0) { mbar_init(&bar, 1); fence_barrier_init(); }
    if (threadIdx.x < 32) tmem_alloc<512>(&slot);
    fence_proxy_async_smem();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = slot;
    if (threadIdx.x == 0) {
        constexpr uint32_t idesc = umma_idesc_f16(128, N, 0, 0);
        const uint64_t da = umma_desc_sw128(smem_u32(smem), 16, 1024);
        const uint64_t db = umma_desc_sw128(smem_u32(smem) + 16384, 16, 1024);
        uint32_t ph = 0;
        const long long t0 = clock64();
        for (int r = 0; r < reps; ++r) {
            for (int c = 0; c < chain; ++c) {
                const uint32_t acc = tmem + ((TWO_ACC && (c & 1)) ? 256 : 0);
                const uint64_t ko = (uint64_t)((c & 3) * 2);
                if (TS) umma_f16_ts(acc, tmem + 448, db + ko, idesc, 1u);
                else umma_f16_ss(acc, da + ko, db + ko, idesc, 1u);
            }
            umma_commit(&bar);
            mbar_wait(&bar, ph);
            ph ^= 1;
        }
        const long long t1 = clock64();
        clk[blockIdx.x] = t1 - t0;
    }
    __syncthreads();
    if (threadIdx.x < 32) { tc_fence_after(); tmem_dealloc<512>(tmem); }
}

template <int N, int TS, int TWO_ACC>
void run(const char* name, int chain) {
    long long* clk;
    cudaMalloc(&clk, 148 * 8);
    auto kern = k<N, TS, TWO_ACC>;
    cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
    const int reps = 2000;
    kern<<<148, 128, 64 * 1024>>>(clk, 10, chain);
    kern<<<148, 128, 64 * 1024>>>(clk, reps, chain);
    cudaError_t e = cudaDeviceSynchronize();
    long long h[148];
    cudaMemcpy(h, clk, sizeof(h), cudaMemcpyDeviceToHost);
    const double per = (double)h[0] / reps;
    printf("%-44s chain %3d  %s  %8.1f clk per batch = %6.1f clk per MMA  (floor %d)  -> %.0f%% of the tensor peak\n", name, chain,
           cudaGetErrorString(e), per, per / chain, N / 2, 100.0 * (N / 2) * chain / per);
    cudaFree(clk);
}

int main() {
    for (int chain : {4, 8, 32}) {
        run<128, 0, 0>("M128 N128 K16 SS, one accumulator", chain);
        run<128, 0, 1>("M128 N128 K16 SS, two accumulators", chain);
        run<64, 0, 0>("M128 N64  K16 SS, one accumulator", chain);
        run<64, 1, 0>("M128 N64  K16 TS (A in TMEM)", chain);
        run<256, 0, 0>("M128 N256 K16 SS, one accumulator", chain);
    }
    return 0;
}
