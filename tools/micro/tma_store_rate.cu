// TMA store / load throughput vs box row width.  The tapgemm2 epilogue moves output (and residual) tiles as [128 rows x 32
// fp16 columns] boxes (64-byte rows, SWIZZLE_64B).  Hypothesis under test (round 2): the short-K GEMMs (K = 320) run at ~7000 clk
// per 128-row tile regardless of BN because every 64-byte row of a box is one TMA request -- wider boxes (128-byte rows,
// SWIZZLE_128B) should move the same bytes in half the requests.
// Each CTA (148, one per SM) issues `reps` stores (or loads) of a [128 x W] box of a row-major [rows x ld] fp16 matrix,
// walking down the rows like the persistent GEMM does; reported: clk per box, B/clk/SM and aggregate GB/s.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I../../star_b200/csrc -o tma_store_rate tma_store_rate.cu -lcuda
#include <cstdio>
#include <cstdlib>
#include <cuda.h>
#include <cudaTypedefs.h>
#include "common.cuh"
using namespace star;

static PFN_cuTensorMapEncodeTiled_v12000 g_encode;

static CUtensorMap make_map(void* base, int rank5, long long rows, long long ld, int box_cols, CUtensorMapSwizzle swz) {
    CUtensorMap m;
    if (!rank5) {
        cuuint64_t gdim[2] = {(cuuint64_t)ld, (cuuint64_t)rows};
        cuuint64_t gstr[1] = {(cuuint64_t)ld * 2};
        cuuint32_t box[2] = {(cuuint32_t)box_cols, 128}, es[2] = {1, 1};
        CUresult r = g_encode(&m, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, base, gdim, gstr, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, swz,
                              CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS) { printf("encode failed %d\n", (int)r); exit(1); }
    } else {                  // the GEMM's (N, n1, n2, n3, n4) view with a [W x 128 x 1 x 1 x 1] box
        cuuint64_t gdim[5] = {(cuuint64_t)ld, (cuuint64_t)rows, 1, 1, 1};
        cuuint64_t gstr[4] = {(cuuint64_t)ld * 2, (cuuint64_t)ld * 2 * rows, (cuuint64_t)ld * 2 * rows, (cuuint64_t)ld * 2 * rows};
        cuuint32_t box[5] = {(cuuint32_t)box_cols, 128, 1, 1, 1}, es[5] = {1, 1, 1, 1, 1};
        CUresult r = g_encode(&m, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 5, base, gdim, gstr, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, swz,
                              CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS) { printf("encode5 failed %d\n", (int)r); exit(1); }
    }
    return m;
}

STAR_DEVINL void st2d(const CUtensorMap* m, const void* src, int c0, int c1) {
    asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];" ::"l"(reinterpret_cast<uint64_t>(m)),
                 "r"(smem_u32(src)), "r"(c0), "r"(c1) : "memory");
}
STAR_DEVINL void st5d(const CUtensorMap* m, const void* src, int c0, int c1) {
    asm volatile("cp.async.bulk.tensor.5d.global.shared::cta.bulk_group [%0, {%2, %3, %4, %4, %4}], [%1];" ::"l"(reinterpret_cast<uint64_t>(m)),
                 "r"(smem_u32(src)), "r"(c0), "r"(c1), "r"(0) : "memory");
}

// MODE 0: stores, wait_group.read after every `per_group` boxes (the epilogue's pattern); MODE 1: loads through a 2-slot ring
template <int RANK5, int MODE>
__global__ void __launch_bounds__(128) k(const __grid_constant__ CUtensorMap m, long long* clk, int reps, int boxes_per_tile, int box_cols,
                                         int tiles_total) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    __shared__ uint64_t bar;
    for (int i = threadIdx.x; i < 65536 / 4; i += 128) reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u;
    if (threadIdx.x == 0) { mbar_init(&bar, 1); fence_barrier_init(); }
    fence_proxy_async_smem();
    __syncthreads();
    if (threadIdx.x == 0) {
        const int box_bytes = 128 * box_cols * 2;
        uint32_t ph = 0;
        const long long t0 = clock64();
        int tile = blockIdx.x;
        for (int r = 0; r < reps; ++r) {
            if (MODE == 0) {
                for (int b = 0; b < boxes_per_tile; ++b) {
                    if (RANK5) st5d(&m, smem + b * box_bytes, b * box_cols, tile * 128);
                    else st2d(&m, smem + b * box_bytes, b * box_cols, tile * 128);
                }
                asm volatile("cp.async.bulk.commit_group;" ::: "memory");
                asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
            } else {
                mbar_expect_tx(&bar, (uint32_t)(boxes_per_tile * box_bytes));
                for (int b = 0; b < boxes_per_tile; ++b) {
                    if (RANK5) tma_load_5d(smem + b * box_bytes, &m, &bar, b * box_cols, tile * 128, 0, 0, 0);
                    else tma_load_2d(smem + b * box_bytes, &m, &bar, b * box_cols, tile * 128);
                }
                mbar_wait(&bar, ph);
                ph ^= 1;
            }
            tile += gridDim.x;
            if (tile >= tiles_total) tile = blockIdx.x;
        }
        if (MODE == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
        clk[blockIdx.x] = clock64() - t0;
    }
    __syncthreads();
}

template <int RANK5, int MODE>
void run(const char* name, void* buf, long long rows, int ld, int box_cols, CUtensorMapSwizzle swz) {
    const int boxes = ld / box_cols;
    CUtensorMap m = make_map(buf, RANK5, rows, ld, box_cols, swz);
    long long* clk;
    cudaMalloc(&clk, 148 * 8);
    auto kern = k<RANK5, MODE>;
    cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024);
    const int tiles = (int)(rows / 128), reps = 4000;
    kern<<<148, 128, 96 * 1024>>>(m, clk, 50, boxes, box_cols, tiles);
    cudaEvent_t a, b;
    cudaEventCreate(&a); cudaEventCreate(&b);
    cudaEventRecord(a);
    kern<<<148, 128, 96 * 1024>>>(m, clk, reps, boxes, box_cols, tiles);
    cudaEventRecord(b);
    cudaError_t e = cudaDeviceSynchronize();
    float ms = 0;
    cudaEventElapsedTime(&ms, a, b);
    long long h[148];
    cudaMemcpy(h, clk, sizeof(h), cudaMemcpyDeviceToHost);
    double avg = 0;
    for (int i = 0; i < 148; ++i) avg += (double)h[i] / 148;
    const double per_tile = avg / reps, bytes_tile = 128.0 * ld * 2;
    printf("%-62s %s  %8.1f clk / 128x%d tile (%d boxes)  %6.1f B/clk/SM  %8.1f GB/s aggregate\n", name, cudaGetErrorString(e), per_tile, ld,
           boxes, bytes_tile / per_tile, 148.0 * reps * bytes_tile / (ms * 1e6));
    cudaFree(clk);
}

int main() {
    void* fn = nullptr;
    cudaDriverEntryPointQueryResult qr;
    cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qr);
    g_encode = (PFN_cuTensorMapEncodeTiled_v12000)fn;
    const long long rows = 843264;
    void* buf;
    cudaMalloc(&buf, rows * 1024 * 2);
    cudaMemset(buf, 0, rows * 1024 * 2);
    puts("# stores: one commit + wait_group.read per tile (the GEMM epilogue's pattern); matrix = 843264 rows");
    run<0, 0>("store 2-D  ld=256  box 32 cols (64 B rows, SWIZZLE_64B)", buf, rows, 256, 32, CU_TENSOR_MAP_SWIZZLE_64B);
    run<1, 0>("store 5-D  ld=256  box 32 cols (64 B rows, SWIZZLE_64B)", buf, rows, 256, 32, CU_TENSOR_MAP_SWIZZLE_64B);
    run<0, 0>("store 2-D  ld=256  box 64 cols (128 B rows, SWIZZLE_128B)", buf, rows, 256, 64, CU_TENSOR_MAP_SWIZZLE_128B);
    run<1, 0>("store 5-D  ld=256  box 64 cols (128 B rows, SWIZZLE_128B)", buf, rows, 256, 64, CU_TENSOR_MAP_SWIZZLE_128B);
    run<0, 0>("store 2-D  ld=256  box 128 cols (256 B rows, no swizzle)", buf, rows, 256, 128, CU_TENSOR_MAP_SWIZZLE_NONE);
    run<0, 0>("store 2-D  ld=256  box 256 cols (512 B rows, no swizzle)", buf, rows, 256, 256, CU_TENSOR_MAP_SWIZZLE_NONE);
    run<1, 0>("store 5-D  ld=320  box 32 cols (64 B rows, SWIZZLE_64B)", buf, rows, 320, 32, CU_TENSOR_MAP_SWIZZLE_64B);
    run<1, 0>("store 5-D  ld=320  box 64 cols (128 B rows, SWIZZLE_128B)", buf, rows, 320, 64, CU_TENSOR_MAP_SWIZZLE_128B);
    run<1, 0>("store 5-D  ld=960  box 32 cols, 8 boxes of a 256-col tile", buf, rows, 256, 32, CU_TENSOR_MAP_SWIZZLE_64B);
    puts("# loads (one mbarrier round trip per tile)");
    run<1, 1>("load  5-D  ld=256  box 32 cols (64 B rows, SWIZZLE_64B)", buf, rows, 256, 32, CU_TENSOR_MAP_SWIZZLE_64B);
    run<1, 1>("load  5-D  ld=256  box 64 cols (128 B rows, SWIZZLE_128B)", buf, rows, 256, 64, CU_TENSOR_MAP_SWIZZLE_128B);
    run<0, 1>("load  2-D  ld=256  box 64 cols (128 B rows, SWIZZLE_128B)", buf, rows, 256, 64, CU_TENSOR_MAP_SWIZZLE_128B);
    return 0;
}
