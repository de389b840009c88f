// Flash-attention forward for sm_100a, head dim 128 -- variant with TWO softmax warpgroups per Q tile (16 softmax warps per CTA).
//
// attn_sm100.cuh gives each 128-row Q tile one warpgroup: a thread owns a row and walks all 128 scores of a K/V step.  ncu (round 2,
// profiles/ncu_r02_attn.txt) shows that kernel latency-bound, not throughput-bound: tensor pipe 78 % of cycles with MUFU at 53 %, FMA at
// 43 % and 51 % of the issue slots used -- each scheduler hosts only two softmax warps, whose chains (wait S -> tcgen05.ld -> max ->
// exp2 -> pack -> tcgen05.st -> arrive) rarely overlap, so the MMA warp waits for P.  Here the two 64-key halves of a tile's score row go
// to two DIFFERENT warpgroups that run concurrently: warpgroup (i, h) handles keys [64 h, 64 h + 64) of Q tile i -- 64 scores per
// thread per step, half the latency before each half of P is published, four softmax warps per scheduler to hide each other's stalls.
// The two threads of a row agree on the reference max through a 1-float exchange in shared memory (double-buffered by step parity,
// one named barrier per step that also orders "both halves have read S" before either overwrites it with P); l is combined once at the
// end; each thread rescales / normalises / stores its own 64 columns of O.
//
// Everything else (TMA producer, MMA issue order PV0_j, S0_{j+1}, PV1_j, S1_{j+1}, TMEM map, lazy rescale) is attn_sm100.cuh's.
#pragma once
#include <cuda.h>

#include "attn_sm100.cuh"

namespace b200 {

constexpr int ATT3_THREADS = 128 + 512;             // 4 service warps + 4 softmax warpgroups
constexpr int ATT3_SMEM_BYTES = ATT_SMEM_BYTES + 2 * 2 * 2 * 128 * 4;      // + max exchange [parity][tile][half][row]

template <int POLY>
__global__ void __launch_bounds__(ATT3_THREADS, 1)
attn_fwd_d128_w16_kernel(const __grid_constant__ CUtensorMap tmap_q, const __grid_constant__ CUtensorMap tmap_k,
                         const __grid_constant__ CUtensorMap tmap_v, const AttnParams p) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint8_t* sQ = smem;
    uint8_t* sK = sQ + ATT_QTILES * ATT_TILE_BYTES;
    uint8_t* sV = sK + ATT_KV_STAGES * ATT_TILE_BYTES;
    uint64_t* bars = reinterpret_cast<uint64_t*>(sV + ATT_KV_STAGES * ATT_TILE_BYTES);
    uint64_t* q_full = bars;            // [1]
    uint64_t* k_full = bars + 1;        // [2]
    uint64_t* k_empty = bars + 3;       // [2]
    uint64_t* v_full = bars + 5;        // [2]
    uint64_t* v_empty = bars + 7;       // [2]
    uint64_t* s_full = bars + 9;        // [2]     per Q tile: MMA -> its two softmax warpgroups
    uint64_t* p_full = bars + 11;       // [2][2]  per Q tile and key half: warpgroup (i, h) -> MMA, 4 arrivals (one per warp)
    uint64_t* pv_done = bars + 15;      // [2]
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 17);
    float* xch = reinterpret_cast<float*>(bars + 20);          // [2 parity][2 tiles][2 halves][128 rows]

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    const int q_blk = blockIdx.x;
    const int head = blockIdx.y;
    const int n_kv = (p.Lk + ATT_BN - 1) / ATT_BN;
    const int q_row0 = blockIdx.z * p.Lq, k_row0 = blockIdx.z * p.Lk;

    if (warp == 0 && lane == 0) { prefetch_tmap(&tmap_q); prefetch_tmap(&tmap_k); prefetch_tmap(&tmap_v); }
    if (warp == 1 && lane == 0) {
        mbar_init(q_full, 1);
        for (int i = 0; i < ATT_KV_STAGES; ++i) {
            mbar_init(&k_full[i], 1); mbar_init(&k_empty[i], 1);
            mbar_init(&v_full[i], 1); mbar_init(&v_empty[i], 1);
        }
        for (int i = 0; i < 2; ++i) { mbar_init(&s_full[i], 1); mbar_init(&pv_done[i], 1); }
        for (int i = 0; i < 4; ++i) mbar_init(&p_full[i], 4);
        fence_mbar_init();
    }
    if (warp == 2) tmem_alloc(tmem_slot, 512);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp < 4) {
    setmaxnreg_dec<64>();
    if (warp == 0) {
        // ============================ TMA producer ============================
        if (elect_one()) {
            const int col = head * ATT_D;
            mbar_arrive_expect_tx(q_full, ATT_QTILES * ATT_TILE_BYTES);
            #pragma unroll
            for (int i = 0; i < ATT_QTILES; ++i) {
                const int r0 = q_row0 + (q_blk * ATT_QTILES + i) * ATT_BM;
                tma_load_2d(sQ + i * ATT_TILE_BYTES, &tmap_q, q_full, col, r0);
                tma_load_2d(sQ + i * ATT_TILE_BYTES + ATT_TILE_BYTES / 2, &tmap_q, q_full, col + 64, r0);
            }
            for (int j = 0; j < n_kv; ++j) {
                const int st = j % ATT_KV_STAGES;
                const uint32_t ph = (j / ATT_KV_STAGES) & 1;
                mbar_wait(&k_empty[st], ph ^ 1);
                mbar_arrive_expect_tx(&k_full[st], ATT_TILE_BYTES);
                tma_load_2d(sK + st * ATT_TILE_BYTES, &tmap_k, &k_full[st], col, k_row0 + j * ATT_BN);
                tma_load_2d(sK + st * ATT_TILE_BYTES + ATT_TILE_BYTES / 2, &tmap_k, &k_full[st], col + 64, k_row0 + j * ATT_BN);
                mbar_wait(&v_empty[st], ph ^ 1);
                mbar_arrive_expect_tx(&v_full[st], ATT_TILE_BYTES);
                tma_load_2d(sV + st * ATT_TILE_BYTES, &tmap_v, &v_full[st], col, k_row0 + j * ATT_BN);
                tma_load_2d(sV + st * ATT_TILE_BYTES + ATT_TILE_BYTES / 2, &tmap_v, &v_full[st], col + 64, k_row0 + j * ATT_BN);
            }
        }
        __syncwarp();
    } else if (warp == 1) {
        // ============================ MMA issuer ============================
        if (elect_one()) {
            constexpr uint32_t idesc_s = umma_idesc_bf16(ATT_BM, ATT_BN, /*b_mn_major=*/false);
            constexpr uint32_t idesc_o = umma_idesc_bf16(ATT_BM, ATT_D, /*b_mn_major=*/true);
            auto issue_s = [&](int i, int j) {
                const uint32_t aQ = smem_u32(sQ + i * ATT_TILE_BYTES);
                const uint32_t aK = smem_u32(sK + (j % ATT_KV_STAGES) * ATT_TILE_BYTES);
                #pragma unroll
                for (int kk = 0; kk < ATT_D / 16; ++kk) {
                    const uint32_t off = (kk >> 2) * (ATT_TILE_BYTES / 2) + (kk & 3) * 32;
                    umma_bf16_ss(tmem_base + i * 128, umma_desc_kmajor_sw128(aQ + off), umma_desc_kmajor_sw128(aK + off), idesc_s, kk != 0);
                }
                umma_commit(&s_full[i]);
            };
            auto issue_pv = [&](int i, int j) {
                const uint32_t aV = smem_u32(sV + (j % ATT_KV_STAGES) * ATT_TILE_BYTES);
                #pragma unroll
                for (int kk = 0; kk < ATT_BN / 16; ++kk) {
                    if (kk == 0 || kk == ATT_BN / 32) {
                        mbar_wait(&p_full[i * 2 + (kk != 0)], j & 1);
                        tc_fence_after();
                    }
                    umma_bf16_ts(tmem_base + 256 + i * 128, tmem_base + i * 128 + kk * 8,
                                 umma_desc_mnmajor_sw128(aV + kk * 2048, ATT_TILE_BYTES / 2), idesc_o, (j | kk) != 0);
                }
                umma_commit(&pv_done[i]);
            };
            mbar_wait(q_full, 0);
            mbar_wait(&k_full[0], 0);
            tc_fence_after();
            issue_s(0, 0);
            issue_s(1, 0);
            umma_commit(&k_empty[0]);
            for (int j = 0; j < n_kv; ++j) {
                const int st = j % ATT_KV_STAGES;
                const uint32_t kvph = (j / ATT_KV_STAGES) & 1;
                const bool more = j + 1 < n_kv;
                mbar_wait(&v_full[st], kvph);
                issue_pv(0, j);
                if (more) {
                    mbar_wait(&k_full[(j + 1) % ATT_KV_STAGES], ((j + 1) / ATT_KV_STAGES) & 1);
                    tc_fence_after();
                    issue_s(0, j + 1);
                }
                issue_pv(1, j);
                umma_commit(&v_empty[st]);
                if (more) {
                    issue_s(1, j + 1);
                    umma_commit(&k_empty[(j + 1) % ATT_KV_STAGES]);
                }
            }
        }
        __syncwarp();
    }
    } else {
        // ============================ softmax warpgroup (tile qi, key half kh) ============================
        setmaxnreg_inc<104>();
        const int sw = warp - 4;
        const int qi = sw >> 3;                     // Q tile
        const int kh = (sw >> 2) & 1;               // key half of every K/V step this warpgroup owns
        const int wq = sw & 3;                      // TMEM lane quarter (== warp % 4)
        const int row = wq * 32 + lane;
        const uint32_t lane_off = (uint32_t)(wq * 32) << 16;
        const uint32_t tS = tmem_base + lane_off + qi * 128;
        const uint32_t tO = tmem_base + lane_off + 256 + qi * 128 + kh * 64;      // this thread's 64 columns of O
        float m_used = -INFINITY;
        float l = 0.f;
        for (int j = 0; j < n_kv; ++j) {
            mbar_wait(&s_full[qi], j & 1);
            tc_fence_after();
            const int valid = p.Lk - j * ATT_BN - kh * 64;      // valid keys inside this half (>= 64 except in the last, partial tile)
            uint32_t v[64];
            tmem_ld_32x32b_x32(tS + kh * 64, v);
            tmem_ld_32x32b_x32(tS + kh * 64 + 32, v + 32);
            tmem_ld_wait();
            if (valid < 64) {
                #pragma unroll
                for (int i = 0; i < 64; ++i)
                    if (i >= valid) v[i] = 0xff800000u;
            }
            float mx4[4];
            #pragma unroll
            for (int i = 0; i < 4; ++i) mx4[i] = fmax3(__uint_as_float(v[3 * i]), __uint_as_float(v[3 * i + 1]), __uint_as_float(v[3 * i + 2]));
            #pragma unroll
            for (int i = 12; i < 64; i += 2) mx4[(i >> 1) & 3] = fmax3(mx4[(i >> 1) & 3], __uint_as_float(v[i]), __uint_as_float(v[i + 1]));
            float mx = fmaxf(fmax3(mx4[0], mx4[1], mx4[2]), mx4[3]) * p.scale_log2;
            // row max over both halves; the barrier also means: both warpgroups hold their S values in registers, P may overwrite S
            float* xs = xch + (((j & 1) * 2 + qi) * 2) * 128;
            xs[kh * 128 + row] = mx;
            named_bar_sync(1 + qi, 256);
            mx = fmaxf(mx, xs[(kh ^ 1) * 128 + row]);
            const bool need = mx > m_used + 8.0f;
            if (__any_sync(0xffffffffu, need)) {                 // identical decision in the partner warp (same rows, same mx, same m_used)
                const float m_new = need ? mx : m_used;
                const float alpha = ex2_approx(m_used - m_new);
                if (j > 0) {
                    mbar_wait(&pv_done[qi], (j - 1) & 1);
                    tc_fence_after();
                    #pragma unroll 1
                    for (int c = 0; c < 4; ++c) {                // 16 columns at a time: the 64 score registers stay live across this rare path
                        uint32_t o[16];
                        tmem_ld_32x32b_x16(tO + c * 16, o);
                        tmem_ld_wait();
                        #pragma unroll
                        for (int i = 0; i < 16; ++i) o[i] = __float_as_uint(__uint_as_float(o[i]) * alpha);
                        tmem_st_32x32b_x16(tO + c * 16, o);
                    }
                    tmem_st_wait();
                }
                l *= alpha;
                m_used = m_new;
            }
            const uint64_t sc2 = pack_f32x2(p.scale_log2, p.scale_log2), nm2 = pack_f32x2(-m_used, -m_used);
            uint64_t acc2[4] = {0ull, 0ull, 0ull, 0ull};
            #pragma unroll
            for (int c = 0; c < 32; ++c) {
                const uint64_t x2 = fma_f32x2(pack_f32x2(__uint_as_float(v[2 * c]), __uint_as_float(v[2 * c + 1])), sc2, nm2);
                float e0, e1;
                if (POLY > 0 && c % (POLY > 0 ? POLY : 1) == POLY - 1) {
                    ex2_poly3_x2(x2, e0, e1);
                } else {
                    float x0, x1;
                    unpack_f32x2(x2, x0, x1);
                    e0 = ex2_approx(x0);
                    e1 = ex2_approx(x1);
                }
                acc2[c & 3] = add_f32x2(acc2[c & 3], pack_f32x2(e0, e1));
                v[c] = pack_bf16x2(e0, e1);
            }
            tmem_st_32x32b_x32(tS + kh * 32, v);                 // bf16 P of keys [64 kh, 64 kh + 64): packed columns [32 kh, 32 kh + 32)
            tmem_st_wait();
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&p_full[qi * 2 + kh]);
            float a0, a1;
            unpack_f32x2(add_f32x2(add_f32x2(acc2[0], acc2[1]), add_f32x2(acc2[2], acc2[3])), a0, a1);
            l += a0 + a1;
        }
        // ---- epilogue: l of the row = sum of both halves; each thread normalises and stores its 64 columns of O
        float* xs = xch + ((n_kv & 1) * 2 + qi) * 2 * 128;       // the buffer the last step did not use
        xs[kh * 128 + row] = l;
        named_bar_sync(1 + qi, 256);
        l += xs[(kh ^ 1) * 128 + row];
        mbar_wait(&pv_done[qi], (n_kv - 1) & 1);
        tc_fence_after();
        const float inv_l = 1.0f / l;
        const long long grow = ((long long)q_blk * ATT_QTILES + qi) * ATT_BM + row;
        __nv_bfloat16* orow = p.out + (q_row0 + grow) * p.ldo + head * ATT_D + kh * 64;
        #pragma unroll 1
        for (int c = 0; c < 2; ++c) {
            uint32_t o[32];
            tmem_ld_32x32b_x32(tO + c * 32, o);
            tmem_ld_wait();
            if (grow < p.Lq) {
                #pragma unroll
                for (int i = 0; i < 32; i += 8) {
                    uint4 pk;
                    pk.x = pack_bf16x2(__uint_as_float(o[i]) * inv_l, __uint_as_float(o[i + 1]) * inv_l);
                    pk.y = pack_bf16x2(__uint_as_float(o[i + 2]) * inv_l, __uint_as_float(o[i + 3]) * inv_l);
                    pk.z = pack_bf16x2(__uint_as_float(o[i + 4]) * inv_l, __uint_as_float(o[i + 5]) * inv_l);
                    pk.w = pack_bf16x2(__uint_as_float(o[i + 6]) * inv_l, __uint_as_float(o[i + 7]) * inv_l);
                    *reinterpret_cast<uint4*>(orow + c * 32 + i) = pk;
                }
            }
            __syncwarp();
        }
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 2) {
        tc_fence_after();
        tmem_dealloc(tmem_base, 512);
    }
}

}  // namespace b200
