// PROTOTYPE (not part of libstar_sm100.so): CTA-pair GEMM for the narrow layers.  Protocol validated on a B200 at the end of
// round 1 (profiles/r01_micro_gemm_2cta_prototype.log: PASS at M = 4096 / K = 512 and at the conv shape M = 105 408 /
// K = 2880; 413 TFLOP/s as a one-tile-per-cluster kernel with a scalar-store epilogue -- the persistent version with the
// tapgemm2 epilogue is next round's work).
//
// Why: the 320-channel convolutions / linears (N = 320 = 2 x 160) cannot use the 128x256 tiles that took the wide layers to
// 1.3-1.46 PFLOP/s; per 128x160 tile and k-block a CTA pulls A 16 KB + B 20 KB through L2 and its
// own shared memory.  With tcgen05 cta_group::2 a PAIR of CTAs (two SMs of one TPC) computes a 256x160 tile: each CTA
// stages its own 128 rows of A and only HALF of the B tile (80 weight rows), the leader CTA issues one M = 256 MMA that
// reads both halves; per CTA and k-block that is 16 + 10 = 26 KB (-28 %), and half the B reads from shared memory.
//
// This file is a self-contained test bench for that kernel (plain C = A W^T, fp16 in / fp32 accumulate / fp16 out):
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I../../star_b200/csrc -o gemm_2cta gemm_2cta.cu -lcuda
//   ./gemm_2cta [M K]         -> max |err| against a straightforward CUDA reference, then TFLOP/s
// Protocol (one 256 x BN tile per cluster, 4-stage ring; to be made persistent once it is green):
//   warp 0 (both CTAs)  TMA producer: own A tile [128 x 64] and own B half [BN/2 x 64] per k-block.  Both CTAs' loads
//                       complete_tx on the LEADER's full barrier (address with the peer bit cleared, as CUTLASS'
//                       SM100_TMA_2SM_LOAD does); the leader's producer arms it with the pair's total byte count.
//   warp 1 (leader)     one thread issues tcgen05.mma.cta_group::2 (M = 256, N = BN) and commits with
//                       .multicast::cluster to the `empty` barrier of BOTH CTAs, finally to both `acc_full` barriers.
//   warps 2-5           epilogue: every CTA reads its own 128 TMEM lanes and stores its own rows.
// Confirmed by that run: (1) tcgen05.alloc.cta_group::2 issued by warp 1 of both CTAs; (2) the B descriptor of an N = BN MMA
// addresses BN/2 rows in each CTA; (3) the peer-bit mask 0xFEFFFFFF on the barrier address routes the peer's TMA completion
// to the leader's barrier; (4) tcgen05.commit ... .multicast::cluster with mask 3 arrives in both CTAs.
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <cuda.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <cudaTypedefs.h>
#include "common.cuh"
using namespace star;

constexpr int BM = 128, BK = 64, BN = 160, STAGES = 4;
constexpr int A_BYTES = BM * BK * 2, BH_BYTES = (BN / 2) * BK * 2, STAGE_BYTES = A_BYTES + BH_BYTES;   // 16 KB + 10 KB
constexpr uint32_t PEER_MASK = 0xFEFFFFFFu;

__device__ __forceinline__ uint32_t cluster_rank() {
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// 2-CTA TMA load: data lands in THIS CTA's shared memory, completion bytes are credited to the leader's barrier
__device__ __forceinline__ void tma_load_2d_2sm(void* dst, const CUtensorMap* m, uint64_t* bar, int c0, int c1) {
    asm volatile(
        "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
        ::"r"(smem_u32(dst)), "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar) & PEER_MASK), "r"(c0), "r"(c1)
        : "memory");
}
__device__ __forceinline__ void umma_f16_ss_2sm(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t idesc, uint32_t acc) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}\n"
        ::"r"(tmem_d), "l"(da), "l"(db), "r"(idesc), "r"(acc)
        : "memory");
}
// arrive on the barrier at this shared-memory offset in BOTH CTAs of the pair once the issued MMAs have retired
__device__ __forceinline__ void umma_commit_2sm(uint64_t* bar) {
    const uint16_t mask = 3;
    asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
                 ::"r"(smem_u32(bar)), "h"(mask)
                 : "memory");
}

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(192, 1)
gemm_2cta_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_w, __half* __restrict__ C,
                 int M, int N, int K) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + STAGES * STAGE_BYTES);
    uint64_t* empty_bar = full_bar + STAGES;
    uint64_t* acc_full = empty_bar + STAGES;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_full + 1);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t rank = cluster_rank();
    const bool leader = rank == 0;
    const int pair = blockIdx.x >> 1;                    // clusters are consecutive block pairs along x
    const int n_tiles = (N + BN - 1) / BN;
    const int n_tile = pair % n_tiles, m_pair = pair / n_tiles;
    const int m0 = (m_pair * 2 + (int)rank) * BM;        // this CTA's rows
    const int n0 = n_tile * BN;
    const int k_iters = (K + BK - 1) / BK;

    if (warp == 1) {
        if (lane == 0) {
            for (int s = 0; s < STAGES; ++s) {
                mbar_init(&full_bar[s], 1);              // leader's copy is the one in use: armed by the leader's producer
                mbar_init(&empty_bar[s], 1);             // one multicast commit per phase
            }
            mbar_init(acc_full, 1);
            fence_barrier_init();
        }
        __syncwarp();
        asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], 256;" ::"r"(smem_u32(tmem_slot)) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    cluster_sync_all();                                  // barriers of both CTAs initialised before any remote arrive
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        if (lane == 0) {
            int s = 0;
            uint32_t ph = 0;
            for (int kc = 0; kc < k_iters; ++kc) {
                mbar_wait(&empty_bar[s], ph ^ 1);
                uint8_t* sa = smem + s * STAGE_BYTES;
                if (leader) mbar_expect_tx(&full_bar[s], 2u * STAGE_BYTES);
                tma_load_2d_2sm(sa, &tmap_a, &full_bar[s], kc * BK, m0);
                tma_load_2d_2sm(sa + A_BYTES, &tmap_w, &full_bar[s], kc * BK, n0 + (int)rank * (BN / 2));
                if (++s == STAGES) { s = 0; ph ^= 1; }
            }
        }
    } else if (warp == 1) {
        if (leader && lane == 0) {
            constexpr uint32_t idesc = umma_idesc_f16(2 * BM, BN, 0, 0);
            const uint64_t da0 = umma_desc_sw128(smem_u32(smem), 16, 1024);
            const uint64_t db0 = umma_desc_sw128(smem_u32(smem) + A_BYTES, 16, 1024);
            constexpr uint64_t STAGE_INC = (uint64_t)(STAGE_BYTES >> 4);
            int s = 0;
            uint32_t ph = 0;
            for (int kc = 0; kc < k_iters; ++kc) {
                mbar_wait(&full_bar[s], ph);
                tc_fence_after();
                const uint64_t da = da0 + STAGE_INC * (uint64_t)s, db = db0 + STAGE_INC * (uint64_t)s;
                umma_f16_ss_2sm(tmem_base, da, db, idesc, kc > 0 ? 1u : 0u);
                umma_f16_ss_2sm(tmem_base, da + 2, db + 2, idesc, 1u);
                umma_f16_ss_2sm(tmem_base, da + 4, db + 4, idesc, 1u);
                umma_f16_ss_2sm(tmem_base, da + 6, db + 6, idesc, 1u);
                umma_commit_2sm(&empty_bar[s]);
                if (++s == STAGES) { s = 0; ph ^= 1; }
            }
            umma_commit_2sm(acc_full);
        }
    } else {
        const int q = warp & 3;                          // TMEM lane quadrant this warp may read
        const int r = q * 32 + lane;
        mbar_wait(acc_full, 0);
        tc_fence_after();
        const uint32_t t_row = tmem_base + ((uint32_t)(q * 32) << 16);
        const long long row = (long long)m0 + r;
        for (int c0 = 0; c0 < BN; c0 += 32) {
            uint32_t v[32];
            tmem_ld32(t_row + c0, v);
            tmem_ld_wait();
            if (row < M) {
                for (int j = 0; j < 32; ++j)
                    if (n0 + c0 + j < N) C[row * N + n0 + c0 + j] = __float2half_rn(__uint_as_float(v[j]));
            }
        }
        tc_fence_before();
    }
    cluster_sync_all();                                  // the peer's shared memory must outlive the leader's last MMA
    if (warp == 1) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, 256;" ::"r"(tmem_base) : "memory");
    }
}

// ---------------------------------------------------------------------------------------------------------------
// Persistent variant (NOT yet run on hardware): 74 clusters loop over the 256 x BN pair-tiles, two TMEM accumulators so that
// the epilogue of tile i overlaps the MMAs of tile i+1.  New protocol element: the leader's MMA thread may only overwrite
// an accumulator once the epilogue warps of BOTH CTAs have drained it -> acc_empty lives in the leader and the peer's
// epilogue threads arrive on it remotely (shared::cluster address with the peer bit cleared).
__device__ __forceinline__ void mbar_arrive_leader(uint64_t* bar) {
    asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(smem_u32(bar) & PEER_MASK) : "memory");
}

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(192, 1)
gemm_2cta_persistent_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_w,
                            __half* __restrict__ C, int M, int N, int K) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + STAGES * STAGE_BYTES);
    uint64_t* empty_bar = full_bar + STAGES;
    uint64_t* acc_full = empty_bar + STAGES;       // 2
    uint64_t* acc_empty = acc_full + 2;            // 2 (leader's copies are the ones in use)
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_empty + 2);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t rank = cluster_rank();
    const bool leader = rank == 0;
    const int cluster_id = blockIdx.x >> 1, n_clusters = gridDim.x >> 1;
    const int n_tiles = (N + BN - 1) / BN, m_pairs = (M + 2 * BM - 1) / (2 * BM);
    const int total = n_tiles * m_pairs;
    const int k_iters = (K + BK - 1) / BK;

    if (warp == 1) {
        if (lane == 0) {
            for (int s = 0; s < STAGES; ++s) {
                mbar_init(&full_bar[s], 1);
                mbar_init(&empty_bar[s], 1);
            }
            for (int b = 0; b < 2; ++b) {
                mbar_init(&acc_full[b], 1);
                mbar_init(&acc_empty[b], 2 * 128);   // 4 epilogue warps of each CTA
            }
            fence_barrier_init();
        }
        __syncwarp();
        asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(tmem_slot)) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    cluster_sync_all();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        if (lane == 0) {
            int s = 0;
            uint32_t ph = 0;
            for (int t = cluster_id; t < total; t += n_clusters) {
                const int n0 = (t % n_tiles) * BN;
                const int m0 = ((t / n_tiles) * 2 + (int)rank) * BM;
                for (int kc = 0; kc < k_iters; ++kc) {
                    mbar_wait(&empty_bar[s], ph ^ 1);
                    uint8_t* sa = smem + s * STAGE_BYTES;
                    if (leader) mbar_expect_tx(&full_bar[s], 2u * STAGE_BYTES);
                    tma_load_2d_2sm(sa, &tmap_a, &full_bar[s], kc * BK, m0);
                    tma_load_2d_2sm(sa + A_BYTES, &tmap_w, &full_bar[s], kc * BK, n0 + (int)rank * (BN / 2));
                    if (++s == STAGES) { s = 0; ph ^= 1; }
                }
            }
        }
    } else if (warp == 1) {
        if (leader && lane == 0) {
            constexpr uint32_t idesc = umma_idesc_f16(2 * BM, BN, 0, 0);
            const uint64_t da0 = umma_desc_sw128(smem_u32(smem), 16, 1024);
            const uint64_t db0 = umma_desc_sw128(smem_u32(smem) + A_BYTES, 16, 1024);
            constexpr uint64_t STAGE_INC = (uint64_t)(STAGE_BYTES >> 4);
            int s = 0, local = 0;
            uint32_t ph = 0;
            for (int t = cluster_id; t < total; t += n_clusters, ++local) {
                const int buf = local & 1;
                mbar_wait(&acc_empty[buf], ((local >> 1) & 1) ^ 1);
                tc_fence_after();
                const uint32_t acc = tmem_base + buf * 256;
                for (int kc = 0; kc < k_iters; ++kc) {
                    mbar_wait(&full_bar[s], ph);
                    tc_fence_after();
                    const uint64_t da = da0 + STAGE_INC * (uint64_t)s, db = db0 + STAGE_INC * (uint64_t)s;
                    umma_f16_ss_2sm(acc, da, db, idesc, kc > 0 ? 1u : 0u);
                    umma_f16_ss_2sm(acc, da + 2, db + 2, idesc, 1u);
                    umma_f16_ss_2sm(acc, da + 4, db + 4, idesc, 1u);
                    umma_f16_ss_2sm(acc, da + 6, db + 6, idesc, 1u);
                    umma_commit_2sm(&empty_bar[s]);
                    if (++s == STAGES) { s = 0; ph ^= 1; }
                }
                umma_commit_2sm(&acc_full[buf]);
            }
        }
    } else {
        const int q = warp & 3;
        const int r = q * 32 + lane;
        int local = 0;
        for (int t = cluster_id; t < total; t += n_clusters, ++local) {
            const int buf = local & 1;
            const int n0 = (t % n_tiles) * BN;
            const long long row = (long long)((t / n_tiles) * 2 + (int)rank) * BM + r;
            mbar_wait(&acc_full[buf], (local >> 1) & 1);
            tc_fence_after();
            const uint32_t t_row = tmem_base + buf * 256 + ((uint32_t)(q * 32) << 16);
            for (int c0 = 0; c0 < BN; c0 += 32) {
                uint32_t v[32];
                tmem_ld32(t_row + c0, v);
                tmem_ld_wait();
                if (c0 + 32 >= BN) {                       // accumulator drained by this thread
                    tc_fence_before();
                    mbar_arrive_leader(&acc_empty[buf]);
                }
                if (row < M && n0 + c0 + 32 <= N) {
                    uint4* dst = reinterpret_cast<uint4*>(C + row * N + n0 + c0);
#pragma unroll
                    for (int u = 0; u < 4; ++u) {
                        uint4 o;
                        o.x = pack_half2(__uint_as_float(v[u * 8 + 0]), __uint_as_float(v[u * 8 + 1]));
                        o.y = pack_half2(__uint_as_float(v[u * 8 + 2]), __uint_as_float(v[u * 8 + 3]));
                        o.z = pack_half2(__uint_as_float(v[u * 8 + 4]), __uint_as_float(v[u * 8 + 5]));
                        o.w = pack_half2(__uint_as_float(v[u * 8 + 6]), __uint_as_float(v[u * 8 + 7]));
                        dst[u] = o;
                    }
                }
            }
        }
        tc_fence_before();
    }
    cluster_sync_all();
    if (warp == 1) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, 512;" ::"r"(tmem_base) : "memory");
    }
}

__global__ void ref_kernel(const __half* A, const __half* W, float* C, int M, int N, int K) {
    const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (i >= (long long)M * N) return;
    const int m = (int)(i / N), n = (int)(i % N);
    float acc = 0.f;
    for (int k = 0; k < K; ++k) acc += __half2float(A[(long long)m * K + k]) * __half2float(W[(long long)n * K + k]);
    C[i] = acc;
}

static PFN_cuTensorMapEncodeTiled_v12000 g_encode;
static CUtensorMap make_map(const void* base, unsigned long long inner, unsigned long long rows, unsigned box_rows) {
    CUtensorMap m;
    cuuint64_t dims[2] = {inner, rows}, strides[1] = {inner * 2};
    cuuint32_t box[2] = {(cuuint32_t)BK, box_rows}, es[2] = {1, 1};
    CUresult r = g_encode(&m, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, const_cast<void*>(base), dims, strides, box, es,
                          CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                          CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { printf("cuTensorMapEncodeTiled failed: %d\n", (int)r); exit(1); }
    return m;
}

int main(int argc, char** argv) {
    const int M = argc > 2 ? atoi(argv[1]) : 26352 * 4, K = argc > 2 ? atoi(argv[2]) : 2880, N = 320;
    void* fn = nullptr;
    cudaDriverEntryPointQueryResult qr;
    cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qr);
    g_encode = reinterpret_cast<PFN_cuTensorMapEncodeTiled_v12000>(fn);
    std::vector<__half> ha((size_t)M * K), hw((size_t)N * K);
    srand(1);
    for (auto& x : ha) x = __float2half((rand() % 2001 - 1000) * 1e-3f);
    for (auto& x : hw) x = __float2half((rand() % 2001 - 1000) * 1e-3f / 32.f);
    __half *A, *W, *C;
    float* R;
    cudaMalloc(&A, ha.size() * 2); cudaMalloc(&W, hw.size() * 2); cudaMalloc(&C, (size_t)M * N * 2); cudaMalloc(&R, (size_t)M * N * 4);
    cudaMemcpy(A, ha.data(), ha.size() * 2, cudaMemcpyHostToDevice);
    cudaMemcpy(W, hw.data(), hw.size() * 2, cudaMemcpyHostToDevice);
    cudaMemset(C, 0, (size_t)M * N * 2);
    const CUtensorMap ta = make_map(A, K, M, BM), tw = make_map(W, K, N, BN / 2);
    const int m_pairs = (M + 2 * BM - 1) / (2 * BM), n_tiles = (N + BN - 1) / BN;
    const size_t smem = STAGES * STAGE_BYTES + 256 + 1024;       // barriers live in the 256 bytes behind the ring
    cudaFuncSetAttribute(gemm_2cta_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    const bool persistent = argc > 3 && atoi(argv[3]) != 0;          // ./gemm_2cta M K 1 -> persistent variant
    cudaFuncSetAttribute(gemm_2cta_persistent_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    cudaDeviceProp prop;
    cudaGetDeviceProperties(&prop, 0);
    const dim3 grid(persistent ? (unsigned)(prop.multiProcessorCount & ~1) : (unsigned)(2 * m_pairs * n_tiles));
    auto launch = [&]() {
        if (persistent) gemm_2cta_persistent_kernel<<<grid, 192, smem>>>(ta, tw, C, M, N, K);
        else gemm_2cta_kernel<<<grid, 192, smem>>>(ta, tw, C, M, N, K);
    };
    launch();
    cudaError_t e = cudaDeviceSynchronize();
    printf("gemm_2cta%s M=%d N=%d K=%d grid=%u: %s\n", persistent ? " (persistent)" : "", M, N, K, grid.x, cudaGetErrorString(e));
    if (e != cudaSuccess) return 1;
    ref_kernel<<<(unsigned)(((long long)M * N + 255) / 256), 256>>>(A, W, R, M, N, K);
    cudaDeviceSynchronize();
    std::vector<__half> hc((size_t)M * N);
    std::vector<float> hr((size_t)M * N);
    cudaMemcpy(hc.data(), C, hc.size() * 2, cudaMemcpyDeviceToHost);
    cudaMemcpy(hr.data(), R, hr.size() * 4, cudaMemcpyDeviceToHost);
    double max_err = 0, max_ref = 0;
    for (size_t i = 0; i < hc.size(); ++i) {
        max_err = fmax(max_err, fabs((double)__half2float(hc[i]) - hr[i]));
        max_ref = fmax(max_ref, fabs((double)hr[i]));
    }
    printf("max |err| %.4g (max |ref| %.4g) -> %s\n", max_err, max_ref, max_err <= 2e-3 * max_ref + 1e-3 ? "PASS" : "FAIL");
    cudaEvent_t a, b;
    cudaEventCreate(&a); cudaEventCreate(&b);
    cudaEventRecord(a);
    for (int i = 0; i < 5; ++i) launch();
    cudaEventRecord(b);
    cudaEventSynchronize(b);
    float ms;
    cudaEventElapsedTime(&ms, a, b);
    printf("%.3f ms per launch = %.0f TFLOP/s\n", ms / 5, 2.0 * M * N * K / (ms / 5) / 1e9);
    return 0;
}
