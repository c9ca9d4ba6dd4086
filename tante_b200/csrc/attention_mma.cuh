// Axial multi-head attention on short sequences (S <= 64, head_dim 32) for the bf16 tensor mode.
//
// Sequences are addressed in place in the (B,T,Hp,Wp) token order (no rearrange copies, reference
// attn_backbone.py:149-162): token(gid, p) = outer*S*inner_sz + p*inner_sz + inner with
// outer = gid / inner_sz, inner = gid % inner_sz  (T: inner_sz = L, causal; H: inner_sz = Wp; W: inner_sz = 1).
//
// One CTA (4 warps) = one block of R_pad rows x 4 heads (one warp per head).  Short sequences are
// packed: G = 16 / S sequences share a 16-row block and attention is masked to stay inside a sequence
// (block-diagonal), so the causal T axis (S = 4) runs on the same tensor-core path as H and W.
// Q/K/V head slices are staged once in XOR-swizzled shared memory with coalesced 16-byte loads;
// QK^T and PV run on mma.sync.m16n8k16 (bf16 in, fp32 accumulate; the tiles are far too small for
// tcgen05), the softmax lives in registers (whole score row fits: S <= 64), O is written back
// through shared memory with coalesced 16-byte stores.  HBM traffic = read qkv once + write out once.
#pragma once
#include "common.cuh"

namespace tante {

__device__ __forceinline__ void ldsm_x4(uint32_t addr, uint32_t& r0, uint32_t& r1, uint32_t& r2, uint32_t& r3) {
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0, %1, %2, %3}, [%4];"
                 : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3) : "r"(addr));
}
__device__ __forceinline__ void ldsm_x4_t(uint32_t addr, uint32_t& r0, uint32_t& r1, uint32_t& r2, uint32_t& r3) {
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0, %1, %2, %3}, [%4];"
                 : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3) : "r"(addr));
}
__device__ __forceinline__ void mma_bf16_16816(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
    asm volatile(
        "mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0, %1, %2, %3}, {%4, %5, %6, %7}, {%8, %9}, "
        "{%0, %1, %2, %3};"
        : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
        : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
__device__ __forceinline__ uint32_t pack_bf16x2(float lo, float hi) {
    __nv_bfloat162 h = __floats2bfloat162_rn(lo, hi);
    return *reinterpret_cast<uint32_t*>(&h);
}
// tile row = 64 B (32 bf16) = four 16-byte chunks, chunk index XOR-swizzled by (row >> 1) & 3
__device__ __forceinline__ uint32_t att_off(int r, int chunk) { return (uint32_t)(r * 64 + ((chunk ^ ((r >> 1) & 3)) << 4)); }

template <int NKB /* R_pad / 8 */, bool DROP = false>
__global__ void __launch_bounds__(128) axial_attention_mma_kernel(const __nv_bfloat16* __restrict__ qkv,
                                                                  __nv_bfloat16* __restrict__ out, int n_groups, int S,
                                                                  int inner_sz, int C, int causal, int G,
                                                                  float scale_log2e, DropCfg drop, uint32_t site, int n_head) {
    pdl_trigger();
    constexpr int R = NKB * 8;
    extern __shared__ __align__(128) uint8_t att_smem[];
    __shared__ long long s_tok[R];
    __shared__ int s_gp[R];              // row -> (group << 8) | position, -1 for padding rows
    // tiles: [mat q,k,v][head 0..3][R rows][64 B]
    const int tid = threadIdx.x, warp = tid / 32, lane = tid % 32;
    const int blk = blockIdx.x, hq = blockIdx.y;
    const int Gv = min(G, n_groups - blk * G);           // sequences actually present in this block
    const int rows_valid = Gv * S;

    // (the kernel is issue-bound: every division below is done once per ROW by one thread, the per-thread loops further
    //  down only add constants to what these tables hold)
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
    // ---- stage Q/K/V slices of 4 heads: per row 3 x 256 contiguous bytes ----
    // thread = (16-byte chunk ch of the 256-byte head-group slice, row group): 16 consecutive threads fetch one slice;
    // (cp.async: all 16-byte requests of a thread are in flight at once, no register staging; src-size 0 zero-fills)
    const int ch = tid & 15, rg = tid >> 4;              // ch = head(2 bits) | part(2 bits); 8 row groups
    const int hh_l = ch >> 2, part_l = ch & 3;
    const int col_l = (hq * 4 + hh_l) * 32 + part_l * 8; // element offset of this thread's chunk inside a C-wide row
    {
        const uint32_t sbase = (uint32_t)__cvta_generic_to_shared(att_smem);
#pragma unroll
        for (int r = rg; r < R; r += 8) {
            const long long tok = s_tok[r];
            const __nv_bfloat16* src = qkv + (tok >= 0 ? (size_t)tok * 3 * C + col_l : 0);
            const uint32_t dst = sbase + (uint32_t)(hh_l * R) * 64 + att_off(r, part_l);
            const int nbytes = tok >= 0 ? 16 : 0;
#pragma unroll
            for (int mat = 0; mat < 3; ++mat)
                asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst + (uint32_t)(mat * 4 * R * 64)),
                             "l"(src + (tok >= 0 ? mat * C : 0)), "r"(nbytes) : "memory");
        }
        asm volatile("cp.async.commit_group;" ::: "memory");
        asm volatile("cp.async.wait_group 0;" ::: "memory");
    }
    __syncthreads();

    const uint32_t sQ = (uint32_t)__cvta_generic_to_shared(att_smem + (size_t)((0 * 4 + warp) * R) * 64);
    const uint32_t sK = (uint32_t)__cvta_generic_to_shared(att_smem + (size_t)((1 * 4 + warp) * R) * 64);
    const uint32_t sV = (uint32_t)__cvta_generic_to_shared(att_smem + (size_t)((2 * 4 + warp) * R) * 64);
    uint8_t* gQ = att_smem + (size_t)((0 * 4 + warp) * R) * 64;
    const int g = lane >> 2, t = lane & 3;
    const int lrow = (lane & 7) + 8 * ((lane >> 3) & 1);   // ldmatrix row supplied by this lane (A / V pattern)
    const int lchk = lane >> 4;                            // 0/1: which 16-byte chunk of the pair

    // per-thread key metadata (the keys this lane sees in every score fragment): group id and position
    int kgrp[NKB][2], kpos[NKB][2];
#pragma unroll
    for (int kb = 0; kb < NKB; ++kb)
#pragma unroll
        for (int j = 0; j < 2; ++j) {
            const int gp = s_gp[kb * 8 + 2 * t + j];
            kgrp[kb][j] = gp < 0 ? -1 : (gp >> 8);
            kpos[kb][j] = gp & 255;
        }

    const bool nomask = !causal && G == 1 && rows_valid == R;
#pragma unroll 1
    for (int qb = 0; qb < R / 16; ++qb) {
        if (qb * 16 >= rows_valid) break;
        uint32_t qa[2][4];
#pragma unroll
        for (int ks = 0; ks < 2; ++ks)
            ldsm_x4(sQ + att_off(qb * 16 + lrow, ks * 2 + lchk), qa[ks][0], qa[ks][1], qa[ks][2], qa[ks][3]);
        float s[NKB][4];
#pragma unroll
        for (int kb = 0; kb < NKB; ++kb) {
            s[kb][0] = s[kb][1] = s[kb][2] = s[kb][3] = 0.f;
            uint32_t k0, k1, k2, k3;
            ldsm_x4(sK + att_off(kb * 8 + (lane & 7), lane >> 3), k0, k1, k2, k3);
            mma_bf16_16816(s[kb], qa[0], k0, k1);
            mma_bf16_16816(s[kb], qa[1], k2, k3);
        }
        // ---- mask + softmax over the key axis (rows g and g+8 of this 16-row block) ----
        const int r0 = qb * 16 + g, r1 = r0 + 8;
        // padding query rows get group -2: no key matches, the row stays fully masked (as before: zero output)
        const int gp0 = s_gp[r0], gp1 = s_gp[r1];
        const int g0 = gp0 < 0 ? -2 : (gp0 >> 8), p0 = gp0 & 255;
        const int g1 = gp1 < 0 ? -2 : (gp1 >> 8), p1 = gp1 & 255;
        float m0 = -INFINITY, m1 = -INFINITY;
        if (nomask) {       // one full, non-causal sequence fills the block: every key is visible to every query
#pragma unroll
            for (int kb = 0; kb < NKB; ++kb) {
#pragma unroll
                for (int j = 0; j < 2; ++j) {
                    m0 = fmaxf(m0, s[kb][j]);
                    m1 = fmaxf(m1, s[kb][2 + j]);
                }
            }
        } else {
#pragma unroll
            for (int kb = 0; kb < NKB; ++kb) {
#pragma unroll
                for (int j = 0; j < 2; ++j) {
                    const int kg = kgrp[kb][j], kp = kpos[kb][j];
                    const bool ok0 = kg == g0 && (!causal || kp <= p0);
                    const bool ok1 = kg == g1 && (!causal || kp <= p1);
                    s[kb][j] = ok0 ? s[kb][j] : -INFINITY;
                    s[kb][2 + j] = ok1 ? s[kb][2 + j] : -INFINITY;
                    m0 = fmaxf(m0, s[kb][j]);
                    m1 = fmaxf(m1, s[kb][2 + j]);
                }
            }
        }
        m0 = fmaxf(m0, __shfl_xor_sync(0xffffffffu, m0, 1)); m0 = fmaxf(m0, __shfl_xor_sync(0xffffffffu, m0, 2));
        m1 = fmaxf(m1, __shfl_xor_sync(0xffffffffu, m1, 1)); m1 = fmaxf(m1, __shfl_xor_sync(0xffffffffu, m1, 2));
        if (m0 == -INFINITY) m0 = 0.f;      // fully masked (padding) row
        if (m1 == -INFINITY) m1 = 0.f;
        float l0 = 0.f, l1 = 0.f;
#pragma unroll
        for (int kb = 0; kb < NKB; ++kb) {
#pragma unroll
            for (int j = 0; j < 2; ++j) {
                s[kb][j] = exp2f((s[kb][j] - m0) * scale_log2e);
                s[kb][2 + j] = exp2f((s[kb][2 + j] - m1) * scale_log2e);
                l0 += s[kb][j];
                l1 += s[kb][2 + j];
            }
        }
        l0 += __shfl_xor_sync(0xffffffffu, l0, 1); l0 += __shfl_xor_sync(0xffffffffu, l0, 2);
        l1 += __shfl_xor_sync(0xffffffffu, l1, 1); l1 += __shfl_xor_sync(0xffffffffu, l1, 2);
        if (DROP) {
            // training: dropout on the (normalised) probabilities -- the normaliser above saw every key.  The two keys this
            // lane holds in a block are lanes (kpos & 7), (kpos & 7) + 1 of one Philox group of (query token, head)
            const int head = hq * 4 + warp;
            const long long t0 = s_tok[r0], t1 = s_tok[r1];
#pragma unroll
            for (int kb = 0; kb < NKB; ++kb) {
                const uint4 w0 = drop_words(drop, site, drop_attn_grp(t0, n_head, head, kpos[kb][0]));
                const uint4 w1 = drop_words(drop, site, drop_attn_grp(t1, n_head, head, kpos[kb][0]));
#pragma unroll
                for (int j = 0; j < 2; ++j) {
                    s[kb][j] *= drop_mul(drop, w0, kpos[kb][j] & 7);
                    s[kb][2 + j] *= drop_mul(drop, w1, kpos[kb][j] & 7);
                }
            }
        }
        // ---- O = P V ----
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
                ldsm_x4_t(sV + att_off(kk * 16 + lrow, 2 * j + lchk), v0, v1, v2, v3);
                mma_bf16_16816(o[2 * j], pa, v0, v1);
                mma_bf16_16816(o[2 * j + 1], pa, v2, v3);
            }
        }
        const float i0 = l0 > 0.f ? 1.0f / l0 : 0.f, i1 = l1 > 0.f ? 1.0f / l1 : 0.f;
        __syncwarp();   // all lanes finished reading this q-block's Q rows before they are overwritten with O
#pragma unroll
        for (int nb = 0; nb < 4; ++nb) {
            *reinterpret_cast<uint32_t*>(gQ + att_off(r0, nb) + t * 4) = pack_bf16x2(o[nb][0] * i0, o[nb][1] * i0);
            *reinterpret_cast<uint32_t*>(gQ + att_off(r1, nb) + t * 4) = pack_bf16x2(o[nb][2] * i1, o[nb][3] * i1);
        }
    }
    __syncthreads();
    // ---- coalesced write-back: per row 4 heads x 64 B = 256 contiguous bytes (same thread map as the loads) ----
#pragma unroll
    for (int r = rg; r < R; r += 8) {
        const long long tok = s_tok[r];
        if (tok < 0) continue;
        const uint4 v = *reinterpret_cast<const uint4*>(att_smem + (size_t)(hh_l * R) * 64 + att_off(r, part_l));
        *reinterpret_cast<uint4*>(out + (size_t)tok * C + col_l) = v;
    }
}

static void att_set_attrs() {
    static unsigned long long attr = 0;
    if (!attrs_needed(attr)) return;
    cudaFuncSetAttribute(axial_attention_mma_kernel<8, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
    cudaFuncSetAttribute(axial_attention_mma_kernel<6, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
    cudaFuncSetAttribute(axial_attention_mma_kernel<8, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
    cudaFuncSetAttribute(axial_attention_mma_kernel<6, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
}

// Host launcher.  Returns false when the configuration is outside this kernel (caller falls back).
static bool launch_attention_mma(const __nv_bfloat16* qkv, __nv_bfloat16* out, long long n_groups, int S, int inner_sz,
                                 int n_head, int C, int head_dim, int causal, cudaStream_t st, cudaError_t* err,
                                 const DropCfg& drop = DropCfg(), uint32_t site = 0) {
    if (head_dim != 32 || n_head % 4 != 0 || S > 64 || S < 1) return false;
    const int G = S <= 16 ? 16 / S : 1;
    const int R = S <= 16 ? 16 : ((S + 15) / 16) * 16;
    const long long blocks = (n_groups + G - 1) / G;
    if (blocks > 0x7fffffffLL) return false;
    dim3 grid((unsigned)blocks, (unsigned)(n_head / 4));
    const size_t smem = (size_t)3 * 4 * R * 64;
    const float sl2 = (1.0f / sqrtf((float)head_dim)) * 1.4426950408889634f;
    att_set_attrs();
    const bool dr = drop.p > 0.f;
    switch (R) {
        case 16: if (dr) axial_attention_mma_kernel<2, true><<<grid, 128, smem, st>>>(qkv, out, (int)n_groups, S, inner_sz, C, causal, G, sl2, drop, site, n_head); else axial_attention_mma_kernel<2, false><<<grid, 128, smem, st>>>(qkv, out, (int)n_groups, S, inner_sz, C, causal, G, sl2, drop, site, n_head); break;
        case 32: if (dr) axial_attention_mma_kernel<4, true><<<grid, 128, smem, st>>>(qkv, out, (int)n_groups, S, inner_sz, C, causal, G, sl2, drop, site, n_head); else axial_attention_mma_kernel<4, false><<<grid, 128, smem, st>>>(qkv, out, (int)n_groups, S, inner_sz, C, causal, G, sl2, drop, site, n_head); break;
        case 48: if (dr) axial_attention_mma_kernel<6, true><<<grid, 128, smem, st>>>(qkv, out, (int)n_groups, S, inner_sz, C, causal, G, sl2, drop, site, n_head); else axial_attention_mma_kernel<6, false><<<grid, 128, smem, st>>>(qkv, out, (int)n_groups, S, inner_sz, C, causal, G, sl2, drop, site, n_head); break;
        default: if (dr) axial_attention_mma_kernel<8, true><<<grid, 128, smem, st>>>(qkv, out, (int)n_groups, S, inner_sz, C, causal, G, sl2, drop, site, n_head); else axial_attention_mma_kernel<8, false><<<grid, 128, smem, st>>>(qkv, out, (int)n_groups, S, inner_sz, C, causal, G, sl2, drop, site, n_head); break;
    }
    *err = cudaGetLastError();
    return true;
}

}  // namespace tante
