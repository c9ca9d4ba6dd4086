// Axial multi-head attention BACKWARD on short sequences (S <= 64, head_dim 32) for the bf16 tensor mode.
// Same tiling as the forward (attention_mma.cuh): one CTA (4 warps) = one block of R_pad rows x 4 heads, one warp
// per head, short sequences packed block-diagonally; Q/K/V/dO head slices staged once in XOR-swizzled shared
// memory, all five contractions on mma.sync.m16n8k16 with the probabilities recomputed from the saved qkv:
//   pass 1 (query-major, per 16-query block):  S = Q K^T, dP = dO V^T, P = softmax(S), D = rowsum(P o dP),
//           dS = P o (dP - D) * scale, dQ = dS K;  per-query (m, 1/l, D) kept in shared memory
//   pass 2 (key-major, per 16-key block):      S^T = K Q^T, dP^T = V dO^T recomputed with the roles swapped, so
//           P^T and dS^T come out of the accumulators directly in A-fragment layout (no shared-memory transpose):
//           dV = P^T dO, dK = dS^T Q; the finished key block's K / V rows are overwritten in place by dK / dV.
// HBM traffic = read qkv + dO once, write dqkv once (2.5 KB + 1.5 KB per token at C = 256).
#pragma once
#include "attention_mma.cuh"

namespace tante {

template <int NKB /* R_pad / 8 */, bool DROP = false>
__global__ void __launch_bounds__(128) axial_attention_bwd_mma_kernel(const __nv_bfloat16* __restrict__ qkv,
                                                                      const __nv_bfloat16* __restrict__ dout,
                                                                      __nv_bfloat16* __restrict__ dqkv, int n_groups,
                                                                      int S, int inner_sz, int C, int causal, int G,
                                                                      float scale, float scale_log2e, DropCfg drop, uint32_t site,
                                                                      int n_head) {
    pdl_trigger();
    constexpr int R = NKB * 8;
    extern __shared__ __align__(128) uint8_t attb_smem[];
    __shared__ long long s_tok[R];
    __shared__ int s_gp[R];              // row -> (group << 8) | position, -1 for padding rows (see attention_mma.cuh)
    const int tid = threadIdx.x, warp = tid / 32, lane = tid % 32;
    const int blk = blockIdx.x, hq = blockIdx.y;
    const int Gv = min(G, n_groups - blk * G);
    const int rows_valid = Gv * S;
    // tiles: [mat q,k,v,dO][head 0..3][R][64 B], then dQ staging [head][R][64 B], then stats [head][3][R] f32
    uint8_t* tiles = attb_smem;
    uint8_t* stage = tiles + (size_t)16 * R * 64;
    float* stats = reinterpret_cast<float*>(stage + (size_t)4 * R * 64);

    if (tid < R) {
        long long tok = -1;
        int gp = -1;
        if (tid < rows_valid) {
            const int i = tid / S, p = tid - i * S;
            const long long gid = (long long)blk * G + i;
            const long long outer = gid / inner_sz, inner = gid - outer * inner_sz;
            tok = outer * S * inner_sz + (long long)p * inner_sz + inner;
            gp = (i << 8) | p;
        }
        s_tok[tid] = tok;
        s_gp[tid] = gp;
    }
    __syncthreads();
    // thread = (16-byte chunk of the 256-byte head-group slice, row group), as in the forward kernel
    const int ch = tid & 15, rg = tid >> 4;
    const int hh_l = ch >> 2, part_l = ch & 3;
    const int col_l = (hq * 4 + hh_l) * 32 + part_l * 8;
    {
        const uint32_t sbase = (uint32_t)__cvta_generic_to_shared(tiles);
#pragma unroll
        for (int r = rg; r < R; r += 8) {
            const long long tok = s_tok[r];
            const int nbytes = tok >= 0 ? 16 : 0;
            const __nv_bfloat16* src = qkv + (tok >= 0 ? (size_t)tok * 3 * C + col_l : 0);
            const uint32_t dst = sbase + (uint32_t)(hh_l * R) * 64 + att_off(r, part_l);
#pragma unroll
            for (int mat = 0; mat < 3; ++mat)
                asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst + (uint32_t)(mat * 4 * R * 64)),
                             "l"(src + (tok >= 0 ? mat * C : 0)), "r"(nbytes) : "memory");
            const __nv_bfloat16* srco = dout + (tok >= 0 ? (size_t)tok * C + col_l : 0);
            asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst + (uint32_t)(3 * 4 * R * 64)), "l"(srco),
                         "r"(nbytes) : "memory");
        }
        asm volatile("cp.async.commit_group;" ::: "memory");
        asm volatile("cp.async.wait_group 0;" ::: "memory");
    }
    __syncthreads();

    const uint32_t sQ = (uint32_t)__cvta_generic_to_shared(tiles + (size_t)((0 * 4 + warp) * R) * 64);
    const uint32_t sK = (uint32_t)__cvta_generic_to_shared(tiles + (size_t)((1 * 4 + warp) * R) * 64);
    const uint32_t sV = (uint32_t)__cvta_generic_to_shared(tiles + (size_t)((2 * 4 + warp) * R) * 64);
    const uint32_t sO = (uint32_t)__cvta_generic_to_shared(tiles + (size_t)((3 * 4 + warp) * R) * 64);
    uint8_t* gK = tiles + (size_t)((1 * 4 + warp) * R) * 64;
    uint8_t* gV = tiles + (size_t)((2 * 4 + warp) * R) * 64;
    uint8_t* gS = stage + (size_t)(warp * R) * 64;
    float* st_m = stats + (size_t)warp * 3 * R;
    float* st_il = st_m + R;
    float* st_D = st_il + R;
    const int g = lane >> 2, t = lane & 3;
    const int lrow = (lane & 7) + 8 * ((lane >> 3) & 1);
    const int lchk = lane >> 4;

    // column metadata of a score fragment: the 2 columns (8*blk + 2t + j) this lane holds in every 8-column block
    int cgrp[NKB][2], cpos[NKB][2];
#pragma unroll
    for (int kb = 0; kb < NKB; ++kb)
#pragma unroll
        for (int j = 0; j < 2; ++j) {
            const int gp = s_gp[kb * 8 + 2 * t + j];
            cgrp[kb][j] = gp < 0 ? -1 : (gp >> 8);
            cpos[kb][j] = gp & 255;
        }

    // ================= pass 1: query-major =================
#pragma unroll 1
    for (int qb = 0; qb < R / 16; ++qb) {
        if (qb * 16 >= rows_valid) {
            // padding block: its dQ staging rows are never written back, stats never read for valid columns
            break;
        }
        uint32_t qa[2][4], oa[2][4];
#pragma unroll
        for (int ks = 0; ks < 2; ++ks) {
            ldsm_x4(sQ + att_off(qb * 16 + lrow, ks * 2 + lchk), qa[ks][0], qa[ks][1], qa[ks][2], qa[ks][3]);
            ldsm_x4(sO + att_off(qb * 16 + lrow, ks * 2 + lchk), oa[ks][0], oa[ks][1], oa[ks][2], oa[ks][3]);
        }
        float s[NKB][4], dp[NKB][4];
#pragma unroll
        for (int kb = 0; kb < NKB; ++kb) {
            s[kb][0] = s[kb][1] = s[kb][2] = s[kb][3] = 0.f;
            dp[kb][0] = dp[kb][1] = dp[kb][2] = dp[kb][3] = 0.f;
            uint32_t k0, k1, k2, k3;
            ldsm_x4(sK + att_off(kb * 8 + (lane & 7), lane >> 3), k0, k1, k2, k3);
            mma_bf16_16816(s[kb], qa[0], k0, k1);
            mma_bf16_16816(s[kb], qa[1], k2, k3);
            ldsm_x4(sV + att_off(kb * 8 + (lane & 7), lane >> 3), k0, k1, k2, k3);
            mma_bf16_16816(dp[kb], oa[0], k0, k1);
            mma_bf16_16816(dp[kb], oa[1], k2, k3);
        }
        const int r0 = qb * 16 + g, r1 = r0 + 8;
        const int gp0 = s_gp[r0], gp1 = s_gp[r1];
        const int g0 = gp0 < 0 ? -2 : (gp0 >> 8), p0 = gp0 & 255;
        const int g1 = gp1 < 0 ? -2 : (gp1 >> 8), p1 = gp1 & 255;
        float m0 = -INFINITY, m1 = -INFINITY;
#pragma unroll
        for (int kb = 0; kb < NKB; ++kb)
#pragma unroll
            for (int j = 0; j < 2; ++j) {
                const int kg = cgrp[kb][j], kp = cpos[kb][j];
                const bool ok0 = kg == g0 && (!causal || kp <= p0);
                const bool ok1 = kg == g1 && (!causal || kp <= p1);
                s[kb][j] = ok0 ? s[kb][j] : -INFINITY;
                s[kb][2 + j] = ok1 ? s[kb][2 + j] : -INFINITY;
                m0 = fmaxf(m0, s[kb][j]);
                m1 = fmaxf(m1, s[kb][2 + j]);
            }
        m0 = fmaxf(m0, __shfl_xor_sync(0xffffffffu, m0, 1)); m0 = fmaxf(m0, __shfl_xor_sync(0xffffffffu, m0, 2));
        m1 = fmaxf(m1, __shfl_xor_sync(0xffffffffu, m1, 1)); m1 = fmaxf(m1, __shfl_xor_sync(0xffffffffu, m1, 2));
        if (m0 == -INFINITY) m0 = 0.f;
        if (m1 == -INFINITY) m1 = 0.f;
        float l0 = 0.f, l1 = 0.f;
#pragma unroll
        for (int kb = 0; kb < NKB; ++kb)
#pragma unroll
            for (int j = 0; j < 2; ++j) {
                s[kb][j] = exp2f((s[kb][j] - m0) * scale_log2e);
                s[kb][2 + j] = exp2f((s[kb][2 + j] - m1) * scale_log2e);
                l0 += s[kb][j];
                l1 += s[kb][2 + j];
            }
        l0 += __shfl_xor_sync(0xffffffffu, l0, 1); l0 += __shfl_xor_sync(0xffffffffu, l0, 2);
        l1 += __shfl_xor_sync(0xffffffffu, l1, 1); l1 += __shfl_xor_sync(0xffffffffu, l1, 2);
        const float i0 = l0 > 0.f ? 1.0f / l0 : 0.f, i1 = l1 > 0.f ? 1.0f / l1 : 0.f;
        if (DROP) {
            // O = (P o Z) V with Z = mask / (1 - p): dP = (dO V^T) o Z; everything below uses the masked dP
            const int head = hq * 4 + warp;
            const long long t0 = s_tok[r0], t1 = s_tok[r1];
#pragma unroll
            for (int kb = 0; kb < NKB; ++kb) {
                const uint4 w0 = drop_words(drop, site, drop_attn_grp(t0, n_head, head, cpos[kb][0]));
                const uint4 w1 = drop_words(drop, site, drop_attn_grp(t1, n_head, head, cpos[kb][0]));
#pragma unroll
                for (int j = 0; j < 2; ++j) {
                    dp[kb][j] *= drop_mul(drop, w0, cpos[kb][j] & 7);
                    dp[kb][2 + j] *= drop_mul(drop, w1, cpos[kb][j] & 7);
                }
            }
        }
        float D0 = 0.f, D1 = 0.f;
#pragma unroll
        for (int kb = 0; kb < NKB; ++kb)
#pragma unroll
            for (int j = 0; j < 2; ++j) {
                s[kb][j] *= i0;
                s[kb][2 + j] *= i1;
                D0 = fmaf(s[kb][j], dp[kb][j], D0);
                D1 = fmaf(s[kb][2 + j], dp[kb][2 + j], D1);
            }
        D0 += __shfl_xor_sync(0xffffffffu, D0, 1); D0 += __shfl_xor_sync(0xffffffffu, D0, 2);
        D1 += __shfl_xor_sync(0xffffffffu, D1, 1); D1 += __shfl_xor_sync(0xffffffffu, D1, 2);
        if (t == 0) {
            st_m[r0] = m0; st_il[r0] = i0; st_D[r0] = D0;
            st_m[r1] = m1; st_il[r1] = i1; st_D[r1] = D1;
        }
        // dS = P o (dP - D) * scale  (overwrites s)
#pragma unroll
        for (int kb = 0; kb < NKB; ++kb)
#pragma unroll
            for (int j = 0; j < 2; ++j) {
                s[kb][j] = s[kb][j] * (dp[kb][j] - D0) * scale;
                s[kb][2 + j] = s[kb][2 + j] * (dp[kb][2 + j] - D1) * scale;
            }
        // dQ = dS K
        float o[4][4];
#pragma unroll
        for (int nb = 0; nb < 4; ++nb) o[nb][0] = o[nb][1] = o[nb][2] = o[nb][3] = 0.f;
#pragma unroll
        for (int kk = 0; kk < NKB / 2; ++kk) {
            uint32_t pa[4];
            pa[0] = pack_bf16x2(s[2 * kk][0], s[2 * kk][1]);
            pa[1] = pack_bf16x2(s[2 * kk][2], s[2 * kk][3]);
            pa[2] = pack_bf16x2(s[2 * kk + 1][0], s[2 * kk + 1][1]);
            pa[3] = pack_bf16x2(s[2 * kk + 1][2], s[2 * kk + 1][3]);
#pragma unroll
            for (int j = 0; j < 2; ++j) {
                uint32_t v0, v1, v2, v3;
                ldsm_x4_t(sK + att_off(kk * 16 + lrow, 2 * j + lchk), v0, v1, v2, v3);
                mma_bf16_16816(o[2 * j], pa, v0, v1);
                mma_bf16_16816(o[2 * j + 1], pa, v2, v3);
            }
        }
#pragma unroll
        for (int nb = 0; nb < 4; ++nb) {
            *reinterpret_cast<uint32_t*>(gS + att_off(r0, nb) + t * 4) = pack_bf16x2(o[nb][0], o[nb][1]);
            *reinterpret_cast<uint32_t*>(gS + att_off(r1, nb) + t * 4) = pack_bf16x2(o[nb][2], o[nb][3]);
        }
    }
    __syncwarp();      // stats written by the t == 0 lanes are read by every lane below

    // ================= pass 2: key-major =================
#pragma unroll 1
    for (int kb16 = 0; kb16 < R / 16; ++kb16) {
        if (kb16 * 16 >= rows_valid) break;
        uint32_t ka[2][4], va[2][4];
#pragma unroll
        for (int ks = 0; ks < 2; ++ks) {
            ldsm_x4(sK + att_off(kb16 * 16 + lrow, ks * 2 + lchk), ka[ks][0], ka[ks][1], ka[ks][2], ka[ks][3]);
            ldsm_x4(sV + att_off(kb16 * 16 + lrow, ks * 2 + lchk), va[ks][0], va[ks][1], va[ks][2], va[ks][3]);
        }
        float s[NKB][4], dp[NKB][4];      // S^T and dP^T: rows = keys (g, g+8 of this block), columns = queries
#pragma unroll
        for (int qb = 0; qb < NKB; ++qb) {
            s[qb][0] = s[qb][1] = s[qb][2] = s[qb][3] = 0.f;
            dp[qb][0] = dp[qb][1] = dp[qb][2] = dp[qb][3] = 0.f;
            uint32_t q0, q1, q2, q3;
            ldsm_x4(sQ + att_off(qb * 8 + (lane & 7), lane >> 3), q0, q1, q2, q3);
            mma_bf16_16816(s[qb], ka[0], q0, q1);
            mma_bf16_16816(s[qb], ka[1], q2, q3);
            ldsm_x4(sO + att_off(qb * 8 + (lane & 7), lane >> 3), q0, q1, q2, q3);
            mma_bf16_16816(dp[qb], va[0], q0, q1);
            mma_bf16_16816(dp[qb], va[1], q2, q3);
        }
        const int r0 = kb16 * 16 + g, r1 = r0 + 8;        // key rows
        const int gp0 = s_gp[r0], gp1 = s_gp[r1];
        const int g0 = gp0 < 0 ? -2 : (gp0 >> 8), p0 = gp0 & 255;
        const int g1 = gp1 < 0 ? -2 : (gp1 >> 8), p1 = gp1 & 255;
#pragma unroll
        for (int qb = 0; qb < NKB; ++qb)
#pragma unroll
            for (int j = 0; j < 2; ++j) {
                const int col = qb * 8 + 2 * t + j;                  // query row
                const int qg = cgrp[qb][j], qp = cpos[qb][j];
                const float m = st_m[col], il = st_il[col], D = st_D[col];
                const bool ok0 = qg == g0 && (!causal || p0 <= qp);
                const bool ok1 = qg == g1 && (!causal || p1 <= qp);
                const float pt0 = ok0 ? exp2f((s[qb][j] - m) * scale_log2e) * il : 0.f;
                const float pt1 = ok1 ? exp2f((s[qb][2 + j] - m) * scale_log2e) * il : 0.f;
                float z0 = 1.f, z1 = 1.f;
                if (DROP) {       // Z of (query = this column, key = rows r0 / r1)
                    const long long tq = s_tok[col];
                    const int head = hq * 4 + warp;
                    z0 = drop_mul(drop, drop_words(drop, site, drop_attn_grp(tq, n_head, head, p0)), p0 & 7);
                    z1 = drop_mul(drop, drop_words(drop, site, drop_attn_grp(tq, n_head, head, p1)), p1 & 7);
                }
                s[qb][j] = pt0 * z0;                                              // (P o Z)^T: dV = (P o Z)^T dO
                s[qb][2 + j] = pt1 * z1;
                dp[qb][j] = ok0 ? pt0 * (dp[qb][j] * z0 - D) * scale : 0.f;       // dS^T (stats of padding queries are unset)
                dp[qb][2 + j] = ok1 ? pt1 * (dp[qb][2 + j] * z1 - D) * scale : 0.f;
            }
        float ov[4][4], ok_[4][4];
#pragma unroll
        for (int nb = 0; nb < 4; ++nb) {
            ov[nb][0] = ov[nb][1] = ov[nb][2] = ov[nb][3] = 0.f;
            ok_[nb][0] = ok_[nb][1] = ok_[nb][2] = ok_[nb][3] = 0.f;
        }
#pragma unroll
        for (int kk = 0; kk < NKB / 2; ++kk) {
            uint32_t pa[4], da[4];
            pa[0] = pack_bf16x2(s[2 * kk][0], s[2 * kk][1]);
            pa[1] = pack_bf16x2(s[2 * kk][2], s[2 * kk][3]);
            pa[2] = pack_bf16x2(s[2 * kk + 1][0], s[2 * kk + 1][1]);
            pa[3] = pack_bf16x2(s[2 * kk + 1][2], s[2 * kk + 1][3]);
            da[0] = pack_bf16x2(dp[2 * kk][0], dp[2 * kk][1]);
            da[1] = pack_bf16x2(dp[2 * kk][2], dp[2 * kk][3]);
            da[2] = pack_bf16x2(dp[2 * kk + 1][0], dp[2 * kk + 1][1]);
            da[3] = pack_bf16x2(dp[2 * kk + 1][2], dp[2 * kk + 1][3]);
#pragma unroll
            for (int j = 0; j < 2; ++j) {
                uint32_t v0, v1, v2, v3;
                ldsm_x4_t(sO + att_off(kk * 16 + lrow, 2 * j + lchk), v0, v1, v2, v3);      // dV = P^T dO
                mma_bf16_16816(ov[2 * j], pa, v0, v1);
                mma_bf16_16816(ov[2 * j + 1], pa, v2, v3);
                ldsm_x4_t(sQ + att_off(kk * 16 + lrow, 2 * j + lchk), v0, v1, v2, v3);      // dK = dS^T Q
                mma_bf16_16816(ok_[2 * j], da, v0, v1);
                mma_bf16_16816(ok_[2 * j + 1], da, v2, v3);
            }
        }
        __syncwarp();   // every lane has loaded this key block's K / V fragments before the rows are overwritten
#pragma unroll
        for (int nb = 0; nb < 4; ++nb) {
            *reinterpret_cast<uint32_t*>(gK + att_off(r0, nb) + t * 4) = pack_bf16x2(ok_[nb][0], ok_[nb][1]);
            *reinterpret_cast<uint32_t*>(gK + att_off(r1, nb) + t * 4) = pack_bf16x2(ok_[nb][2], ok_[nb][3]);
            *reinterpret_cast<uint32_t*>(gV + att_off(r0, nb) + t * 4) = pack_bf16x2(ov[nb][0], ov[nb][1]);
            *reinterpret_cast<uint32_t*>(gV + att_off(r1, nb) + t * 4) = pack_bf16x2(ov[nb][2], ov[nb][3]);
        }
    }
    __syncthreads();
    // ---- coalesced write-back of (dq, dk, dv): per row 3 x 4 heads x 64 B (same thread map as the loads) ----
#pragma unroll
    for (int r = rg; r < R; r += 8) {
        const long long tok = s_tok[r];
        if (tok < 0) continue;
        __nv_bfloat16* dst = dqkv + (size_t)tok * 3 * C + col_l;
        const uint32_t so = (uint32_t)(hh_l * R) * 64 + att_off(r, part_l);
        *reinterpret_cast<uint4*>(dst) = *reinterpret_cast<const uint4*>(stage + so);
        *reinterpret_cast<uint4*>(dst + C) = *reinterpret_cast<const uint4*>(tiles + (size_t)(1 * 4 * R) * 64 + so);
        *reinterpret_cast<uint4*>(dst + 2 * C) = *reinterpret_cast<const uint4*>(tiles + (size_t)(2 * 4 * R) * 64 + so);
    }
}

static void attb_set_attrs() {
    static unsigned long long attr = 0;
    if (!attrs_needed(attr)) return;
    cudaFuncSetAttribute(axial_attention_bwd_mma_kernel<4, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
    cudaFuncSetAttribute(axial_attention_bwd_mma_kernel<4, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
    cudaFuncSetAttribute(axial_attention_bwd_mma_kernel<6, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
    cudaFuncSetAttribute(axial_attention_bwd_mma_kernel<6, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
    cudaFuncSetAttribute(axial_attention_bwd_mma_kernel<8, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
    cudaFuncSetAttribute(axial_attention_bwd_mma_kernel<8, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
}

// Host launcher.  Returns false when the configuration is outside this kernel (caller falls back to the SIMT kernel).
static bool launch_attention_bwd_mma(const __nv_bfloat16* qkv, const __nv_bfloat16* dout, __nv_bfloat16* dqkv,
                                     long long n_groups, int S, int inner_sz, int n_head, int C, int head_dim, int causal,
                                     cudaStream_t st, cudaError_t* err, const DropCfg& drop = DropCfg(), uint32_t site = 0) {
    if (head_dim != 32 || n_head % 4 != 0 || S > 64 || S < 1) return false;
    const int G = S <= 16 ? 16 / S : 1;
    const int R = S <= 16 ? 16 : ((S + 15) / 16) * 16;
    const long long blocks = (n_groups + G - 1) / G;
    if (blocks > 0x7fffffffLL) return false;
    dim3 grid((unsigned)blocks, (unsigned)(n_head / 4));
    const size_t smem = (size_t)20 * R * 64 + (size_t)12 * R * 4;
    const float scale = 1.0f / sqrtf((float)head_dim);
    const float sl2 = scale * 1.4426950408889634f;
    attb_set_attrs();
    const bool dr = drop.p > 0.f;
    switch (R) {
        case 16: if (dr) axial_attention_bwd_mma_kernel<2, true><<<grid, 128, smem, st>>>(qkv, dout, dqkv, (int)n_groups, S, inner_sz, C, causal, G, scale, sl2, drop, site, n_head); else axial_attention_bwd_mma_kernel<2, false><<<grid, 128, smem, st>>>(qkv, dout, dqkv, (int)n_groups, S, inner_sz, C, causal, G, scale, sl2, drop, site, n_head); break;
        case 32: if (dr) axial_attention_bwd_mma_kernel<4, true><<<grid, 128, smem, st>>>(qkv, dout, dqkv, (int)n_groups, S, inner_sz, C, causal, G, scale, sl2, drop, site, n_head); else axial_attention_bwd_mma_kernel<4, false><<<grid, 128, smem, st>>>(qkv, dout, dqkv, (int)n_groups, S, inner_sz, C, causal, G, scale, sl2, drop, site, n_head); break;
        case 48: if (dr) axial_attention_bwd_mma_kernel<6, true><<<grid, 128, smem, st>>>(qkv, dout, dqkv, (int)n_groups, S, inner_sz, C, causal, G, scale, sl2, drop, site, n_head); else axial_attention_bwd_mma_kernel<6, false><<<grid, 128, smem, st>>>(qkv, dout, dqkv, (int)n_groups, S, inner_sz, C, causal, G, scale, sl2, drop, site, n_head); break;
        default: if (dr) axial_attention_bwd_mma_kernel<8, true><<<grid, 128, smem, st>>>(qkv, dout, dqkv, (int)n_groups, S, inner_sz, C, causal, G, scale, sl2, drop, site, n_head); else axial_attention_bwd_mma_kernel<8, false><<<grid, 128, smem, st>>>(qkv, dout, dqkv, (int)n_groups, S, inner_sz, C, causal, G, scale, sl2, drop, site, n_head); break;
    }
    *err = cudaGetLastError();
    return true;
}

}  // namespace tante
