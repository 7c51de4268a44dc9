// star_b200 / csrc / attn.cuh
// Flash-style attention forward for head_dim 64, fp16 in / fp32 softmax + accumulate:
//     O = softmax(Q K^T * scale) V          (no bias; optional causal mask for the text tower)
// replaces xformers.ops.memory_efficient_attention at unet_v2v.py:179/:184 for the
// spatial self-attention (N = H*W up to 26 352) and the text cross-attention (Nk = 77).
//
// One CTA = 128 query rows of one (batch, head).  Warp roles (192 threads):
//   warp 0     TMA producer: Q once, then K/V tiles (128 keys) through a 3-stage ring
//   warp 1     TMEM owner + single-thread tcgen05.mma issuer:
//                S_j  = Q K_j^T   -> TMEM S[j&1]   (128x128 fp32, double buffered)
//                Op_j = P_j V_j   -> TMEM O[j&1]   (128x64 fp32)
//   warps 2-5  softmax: thread r owns query row r (TMEM lane r): row max / exp2 / sum with
//              no cross-thread traffic, P_j written as fp16 into the SWIZZLE_128B K-major smem
//              layout the PV MMA reads; running output kept in registers and rescaled by
//              exp2(m_old - m_new) when the partial product is folded in.
// S_{j+1} is issued before P_j is consumed, so the QK^T of the next tile overlaps the softmax
// of the current one.
#pragma once
#include "common.cuh"

namespace star {

constexpr int AT_BQ = 128;
constexpr int AT_BKV = 128;
constexpr int AT_D = 64;
constexpr int AT_KV_STAGES = 3;
constexpr int AT_THREADS = 192;

struct AttnParams {
    int Nq, Nk;
    int causal;             // 1: query row q attends keys 0..q only (text tower); honoured by attn_fwd_kernel, not by attn4
    int kv_batch_div;       // kv batch index = q batch index / kv_batch_div (text context shared by the frames of a clip)
    float scale_log2;       // softmax scale * log2(e)
    __half* out;
    long long ldo;          // elements between consecutive rows of O
};

// SINGLE = the whole key range is one KV tile (text cross-attention, Nk = 77): no rings, one S / P / O buffer,
// 256 TMEM columns and 81 KB of shared memory, so that TWO CTAs are resident per SM -- such a CTA is a strictly
// serial load -> MMA -> softmax -> MMA -> store chain and only co-resident CTAs overlap it (the general
// configuration holds one CTA per SM: 1.34 ms for the 26 352-token level, 12 % of the HBM rate).
template <bool SINGLE>
struct AttnSmemT {
    static constexpr int STAGES = SINGLE ? 1 : AT_KV_STAGES;
    static constexpr int NBUF = SINGLE ? 1 : 2;
    static constexpr int Q_BYTES = AT_BQ * AT_D * 2;        // 16 KB
    static constexpr int KV_BYTES = AT_BKV * AT_D * 2;      // 16 KB each
    static constexpr int P_BYTES = AT_BQ * AT_BKV * 2;      // 32 KB
    static constexpr int OFF_Q = 0;
    static constexpr int OFF_K = OFF_Q + Q_BYTES;
    static constexpr int OFF_V = OFF_K + STAGES * KV_BYTES;
    static constexpr int OFF_P = OFF_V + STAGES * KV_BYTES;
    static constexpr int OFF_BAR = OFF_P + NBUF * P_BYTES;
    static constexpr int TOTAL = OFF_BAR + 256 + 1024;
};
using AttnSmem = AttnSmemT<false>;

template <bool SINGLE>
__global__ void __launch_bounds__(AT_THREADS, SINGLE ? 2 : 1)
attn_fwd_kernel(const __grid_constant__ CUtensorMap tmap_q, const __grid_constant__ CUtensorMap tmap_k,
                const __grid_constant__ CUtensorMap tmap_v, const __grid_constant__ AttnParams p) {
    using AttnSmem = AttnSmemT<SINGLE>;
    constexpr int AT_KV_STAGES = AttnSmem::STAGES;           // shadows the namespace constant
    constexpr uint32_t TMEM_COLS = SINGLE ? 256 : 512;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + AttnSmem::OFF_BAR);
    uint64_t* q_full = bars;                 // 1
    uint64_t* kv_full = bars + 1;            // 3
    uint64_t* kv_empty = bars + 4;           // 3
    uint64_t* s_full = bars + 7;             // 2
    uint64_t* s_empty = bars + 9;            // 2
    uint64_t* p_full = bars + 11;            // 2
    uint64_t* p_empty = bars + 13;           // 2
    uint64_t* o_full = bars + 15;            // 2
    uint64_t* o_empty = bars + 17;           // 2
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 19);

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    const int q0 = blockIdx.x * AT_BQ;
    const int head = blockIdx.y;
    const int batch = blockIdx.z;
    const int kv_batch = batch / p.kv_batch_div;
    const int nt = (p.Nk + AT_BKV - 1) / AT_BKV;

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&tmap_q);
        tma_prefetch_desc(&tmap_k);
        tma_prefetch_desc(&tmap_v);
    }
    if (warp == 1) {
        if (lane == 0) {
            mbar_init(q_full, 1);
            for (int s = 0; s < AT_KV_STAGES; ++s) {
                mbar_init(&kv_full[s], 1);
                mbar_init(&kv_empty[s], 1);
            }
            for (int b = 0; b < 2; ++b) {
                mbar_init(&s_full[b], 1);
                mbar_init(&s_empty[b], 128);
                mbar_init(&p_full[b], 128);
                mbar_init(&p_empty[b], 1);
                mbar_init(&o_full[b], 1);
                mbar_init(&o_empty[b], 128);
            }
            fence_barrier_init();
        }
        __syncwarp();
        tmem_alloc<TMEM_COLS>(tmem_slot);
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    const uint32_t tmem_s0 = tmem_base;            // S[b] at columns b*128
    const uint32_t tmem_o0 = tmem_base + (SINGLE ? 128 : 256);      // O[b] at columns 256 + b*64 (SINGLE: one O at 128)

    if (warp == 0) {
        if (lane == 0) {
            mbar_expect_tx(q_full, AttnSmem::Q_BYTES);
            tma_load_3d(smem + AttnSmem::OFF_Q, &tmap_q, q_full, head * AT_D, q0, batch);
            for (int j = 0; j < nt; ++j) {
                const int s = j % AT_KV_STAGES;
                const uint32_t ph = (j / AT_KV_STAGES) & 1;
                mbar_wait(&kv_empty[s], ph ^ 1);
                mbar_expect_tx(&kv_full[s], 2 * AttnSmem::KV_BYTES);
                tma_load_3d(smem + AttnSmem::OFF_K + s * AttnSmem::KV_BYTES, &tmap_k, &kv_full[s], head * AT_D,
                            j * AT_BKV, kv_batch);
                tma_load_3d(smem + AttnSmem::OFF_V + s * AttnSmem::KV_BYTES, &tmap_v, &kv_full[s], head * AT_D,
                            j * AT_BKV, kv_batch);
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {
            constexpr uint32_t idesc_s = umma_idesc_f16(AT_BQ, AT_BKV, 0, 0);   // S = Q K^T : both K-major
            constexpr uint32_t idesc_o = umma_idesc_f16(AT_BQ, AT_D, 0, 1);     // O = P V   : V is MN-major
            const uint32_t q_addr = smem_u32(smem + AttnSmem::OFF_Q);
            auto issue_s = [&](int j) {
                const int s = j % AT_KV_STAGES;
                const int b = j & 1;
                mbar_wait(&kv_full[s], (j / AT_KV_STAGES) & 1);
                mbar_wait(&s_empty[b], ((j >> 1) & 1) ^ 1);
                tc_fence_after();
                const uint32_t k_addr = smem_u32(smem + AttnSmem::OFF_K + s * AttnSmem::KV_BYTES);
#pragma unroll
                for (int k = 0; k < AT_D / 16; ++k) {
                    umma_f16_ss(tmem_s0 + b * 128, umma_desc_sw128(q_addr + k * 32, 16, 1024),
                                umma_desc_sw128(k_addr + k * 32, 16, 1024), idesc_s, k > 0 ? 1u : 0u);
                }
                umma_commit(&s_full[b]);
            };
            mbar_wait(q_full, 0);
            issue_s(0);
            for (int j = 0; j < nt; ++j) {
                if (j + 1 < nt) issue_s(j + 1);
                const int s = j % AT_KV_STAGES;
                const int b = j & 1;
                mbar_wait(&p_full[b], (j >> 1) & 1);
                mbar_wait(&o_empty[b], ((j >> 1) & 1) ^ 1);
                tc_fence_after();
                const uint32_t p_addr = smem_u32(smem + AttnSmem::OFF_P + b * AttnSmem::P_BYTES);
                const uint32_t v_addr = smem_u32(smem + AttnSmem::OFF_V + s * AttnSmem::KV_BYTES);
#pragma unroll
                for (int k = 0; k < AT_BKV / 16; ++k) {
                    const uint64_t da = umma_desc_sw128(p_addr + (k >> 2) * 16384 + (k & 3) * 32, 16, 1024);
                    const uint64_t db = umma_desc_sw128(v_addr + k * 2048, 8192, 1024);
                    umma_f16_ss(tmem_o0 + b * 64, da, db, idesc_o, k > 0 ? 1u : 0u);
                }
                umma_commit(&o_full[b]);
                umma_commit(&kv_empty[s]);
                umma_commit(&p_empty[b]);
            }
        }
    } else {
        const int quad = warp & 3;
        const int r = quad * 32 + lane;
        const uint32_t lane_off = (uint32_t)(quad * 32) << 16;
        float m_run = -INFINITY, l_run = 0.f, alpha_prev = 0.f;
        float o[AT_D];
#pragma unroll
        for (int i = 0; i < AT_D; ++i) o[i] = 0.f;

        auto fold_o = [&](int j) {      // o = o * alpha_j + Opart_j
            const int b = j & 1;
            mbar_wait(&o_full[b], (j >> 1) & 1);
            tc_fence_after();
            uint32_t v[32];
#pragma unroll
            for (int c = 0; c < 2; ++c) {
                tmem_ld32(tmem_o0 + b * 64 + lane_off + c * 32, v);
                tmem_ld_wait();
#pragma unroll
                for (int i = 0; i < 32; ++i) o[c * 32 + i] = o[c * 32 + i] * alpha_prev + __uint_as_float(v[i]);
            }
            tc_fence_before();
            mbar_arrive(&o_empty[b]);
        };

        for (int j = 0; j < nt; ++j) {
            const int b = j & 1;
            const int kbase = j * AT_BKV;
            const int klim = p.causal ? min(p.Nk, q0 + r + 1) : p.Nk;      // keys [0, klim) are visible to this row
            const bool tail = (kbase + AT_BKV > klim);
            mbar_wait(&s_full[b], (j >> 1) & 1);
            tc_fence_after();
            const uint32_t t_s = tmem_s0 + b * 128 + lane_off;
            // pass A: row max
            float mx = -INFINITY;
            uint32_t v[32];
#pragma unroll 1
            for (int c = 0; c < 4; ++c) {
                tmem_ld32(t_s + c * 32, v);
                tmem_ld_wait();
#pragma unroll
                for (int i = 0; i < 32; ++i) {
                    float sv = __uint_as_float(v[i]);
                    if (tail && kbase + c * 32 + i >= klim) sv = -INFINITY;
                    mx = fmaxf(mx, sv);
                }
            }
            const float m_new = fmaxf(m_run, mx * p.scale_log2);
            const float alpha = exp2f(m_run - m_new);
            m_run = m_new;
            // P buffer must have been consumed by PV_{j-2}
            mbar_wait(&p_empty[b], ((j >> 1) & 1) ^ 1);
            uint8_t* p_row = smem + AttnSmem::OFF_P + b * AttnSmem::P_BYTES + r * 128;
            float l_part = 0.f;
            // pass B: probabilities
#pragma unroll 1
            for (int c = 0; c < 4; ++c) {
                tmem_ld32(t_s + c * 32, v);
                tmem_ld_wait();
                uint32_t pk[16];
#pragma unroll
                for (int i = 0; i < 32; i += 2) {
                    float p0 = exp2f(__uint_as_float(v[i]) * p.scale_log2 - m_new);
                    float p1 = exp2f(__uint_as_float(v[i + 1]) * p.scale_log2 - m_new);
                    if (tail) {
                        if (kbase + c * 32 + i >= klim) p0 = 0.f;
                        if (kbase + c * 32 + i + 1 >= klim) p1 = 0.f;
                    }
                    // the row sum uses the fp16-rounded probabilities that the PV MMA consumes
                    __half2 h = __floats2half2_rn(p0, p1);
                    float2 hf = __half22float2(h);
                    l_part += hf.x + hf.y;
                    pk[i >> 1] = *reinterpret_cast<uint32_t*>(&h);
                }
                // columns [c*32, c*32+32) -> 64-col block (c>>1), 16-byte chunks ((c&1)*4 .. +3), swizzled by row
                uint8_t* blk = p_row + (c >> 1) * 16384;
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    const int chunk = ((c & 1) * 4 + u) ^ (r & 7);
                    *reinterpret_cast<uint4*>(blk + chunk * 16) = make_uint4(pk[u * 4], pk[u * 4 + 1], pk[u * 4 + 2], pk[u * 4 + 3]);
                }
            }
            tc_fence_before();
            mbar_arrive(&s_empty[b]);          // S[b] fully read
            fence_proxy_async_smem();          // make P visible to the async (UMMA) proxy
            mbar_arrive(&p_full[b]);
            l_run = l_run * alpha + l_part;
            if (j > 0) fold_o(j - 1);
            alpha_prev = alpha;
        }
        fold_o(nt - 1);
        const int q = q0 + r;
        if (q < p.Nq) {
            const float inv = 1.0f / l_run;
            __half* op = p.out + ((long long)batch * p.Nq + q) * p.ldo + head * AT_D;
#pragma unroll
            for (int u = 0; u < 8; ++u) {
                uint4 w;
                w.x = pack_half2(o[u * 8 + 0] * inv, o[u * 8 + 1] * inv);
                w.y = pack_half2(o[u * 8 + 2] * inv, o[u * 8 + 3] * inv);
                w.z = pack_half2(o[u * 8 + 4] * inv, o[u * 8 + 5] * inv);
                w.w = pack_half2(o[u * 8 + 6] * inv, o[u * 8 + 7] * inv);
                reinterpret_cast<uint4*>(op)[u] = w;
            }
        }
    }
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        tmem_dealloc<TMEM_COLS>(tmem_base);
    }
}

}  // namespace star
