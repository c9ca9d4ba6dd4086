// Attention backward for sequences LONGER than 64 tokens: the composite axes L = (h w), Y = (t h), A = (t h w) of Attn_Backbone
// (reference models/attn_backbone.py:164-182; autograd of nn.MultiheadAttention's core, :74-80) and plain axes of 65 .. 96
// tokens.  The short-sequence kernels keep a whole sequence's score matrix in shared memory; here keys and queries stream
// through 64 x 64 tiles and the probabilities are RECOMPUTED from the saved qkv (nothing of size S x S is ever stored):
//
//   1. attn_long_stats_kernel  per query row: LSE_i = log sum_j exp(s_ij) and delta_i = sum_j P_ij dP_ij  (online over key tiles)
//   2. attn_long_dq_kernel     per 64-query tile, over key tiles:   dS = P o (dP - delta) * scale,  dQ += dS K
//   3. attn_long_dkv_kernel    per 64-key tile, over query tiles:   dV += (P o Z)^T dO,  dK += dS^T Q
//
// with s = scale * Q K^T (+ causal mask), P = exp(s - LSE), dP = (dO V^T) o Z, Z = the dropout multipliers the forward drew
// (dropout.cuh; Z = 1 without dropout).  No atomics: every element of dqkv is written exactly once, results are deterministic.
// Same in-place token addressing as the forward kernels: token(pos) = (outer * S + pos) * inner_sz + inner, packed qkv rows
// of 3C.  fp32 arithmetic on FFMA for both activation types (the exact mode's path, and the tensor mode's path for these
// inference-first axes); 256 threads, thread (ty, tx) owns the 4 x 4 sub-tile rows ty + 16 a, columns tx + 16 b.
#pragma once
#include "common.cuh"
#include "dropout.cuh"

namespace tante {

constexpr int kAlbTile = 64;
constexpr int kAlbSP = kAlbTile + 1;      // pitch of the 64 x 64 score tiles

template <typename TA, int HD>
__device__ __forceinline__ void alb_load_tile(float* s, const TA* base, size_t row_stride, int pos0, int S) {
    constexpr int P = HD + 1;
    for (int i = threadIdx.x; i < kAlbTile * (HD / 4); i += blockDim.x) {
        const int r = i / (HD / 4), d4 = (i % (HD / 4)) * 4;
        const int pos = pos0 + r;
        float v[4] = {0.f, 0.f, 0.f, 0.f};
        if (pos < S) Vec4<TA>::load(base + (size_t)pos * row_stride + d4, v);
#pragma unroll
        for (int j = 0; j < 4; ++j) s[r * P + d4 + j] = v[j];
    }
}

// acc[a][b] = sum_d A[ty + 16 a][d] * B[tx + 16 b][d]
template <int HD>
__device__ __forceinline__ void alb_dots(const float* sA, const float* sB, int ty, int tx, float (&acc)[4][4]) {
    constexpr int P = HD + 1;
#pragma unroll
    for (int a = 0; a < 4; ++a)
#pragma unroll
        for (int b = 0; b < 4; ++b) acc[a][b] = 0.f;
#pragma unroll 8
    for (int d = 0; d < HD; ++d) {
        float av[4], bv[4];
#pragma unroll
        for (int a = 0; a < 4; ++a) av[a] = sA[(ty + 16 * a) * P + d];
#pragma unroll
        for (int b = 0; b < 4; ++b) bv[b] = sB[(tx + 16 * b) * P + d];
#pragma unroll
        for (int a = 0; a < 4; ++a)
#pragma unroll
            for (int b = 0; b < 4; ++b) acc[a][b] = fmaf(av[a], bv[b], acc[a][b]);
    }
}

__device__ __forceinline__ float alb_half_max(float v) {      // over the 16 lanes that share ty
#pragma unroll
    for (int o = 1; o < 16; o <<= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}
__device__ __forceinline__ float alb_half_sum(float v) {
#pragma unroll
    for (int o = 1; o < 16; o <<= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

struct AlbGeom {
    int S, inner_sz, n_head, C, causal;
    float scale;
};

// blockIdx.x = sequence * n_head + head, blockIdx.y = 64-row tile
__device__ __forceinline__ void alb_locate(const AlbGeom& g, int& head, size_t& tok0) {
    head = blockIdx.x % g.n_head;
    const long long seq = blockIdx.x / g.n_head;
    const long long outer = seq / g.inner_sz, inner = seq % g.inner_sz;
    tok0 = (size_t)outer * g.S * g.inner_sz + (size_t)inner;
}

template <typename TA, int HD>
__global__ void __launch_bounds__(256) attn_long_stats_kernel(const TA* __restrict__ qkv, const TA* __restrict__ dout,
                                                              float* __restrict__ lse, float* __restrict__ delta, AlbGeom g,
                                                              DropCfg drop, uint32_t site) {
    extern __shared__ float alb_smem[];
    constexpr int P = HD + 1;
    float* sQ = alb_smem;
    float* sO = sQ + kAlbTile * P;
    float* sK = sO + kAlbTile * P;
    float* sV = sK + kAlbTile * P;
    int head; size_t tok0;
    alb_locate(g, head, tok0);
    const int S = g.S, C = g.C, ld = 3 * C;
    const int q0 = blockIdx.y * kAlbTile;
    const int ty = threadIdx.x >> 4, tx = threadIdx.x & 15;
    const TA* qb = qkv + tok0 * ld + head * HD;
    alb_load_tile<TA, HD>(sQ, qb, (size_t)g.inner_sz * ld, q0, S);
    alb_load_tile<TA, HD>(sO, dout + tok0 * C + head * HD, (size_t)g.inner_sz * C, q0, S);
    float m[4], l[4], acc[4];
#pragma unroll
    for (int a = 0; a < 4; ++a) { m[a] = -INFINITY; l[a] = 0.f; acc[a] = 0.f; }
    const int nkb = (S + kAlbTile - 1) / kAlbTile;
    for (int kb = 0; kb < nkb; ++kb) {
        __syncthreads();
        alb_load_tile<TA, HD>(sK, qb + C, (size_t)g.inner_sz * ld, kb * kAlbTile, S);
        alb_load_tile<TA, HD>(sV, qb + 2 * C, (size_t)g.inner_sz * ld, kb * kAlbTile, S);
        __syncthreads();
        float s[4][4], dp[4][4];
        alb_dots<HD>(sQ, sK, ty, tx, s);
        alb_dots<HD>(sO, sV, ty, tx, dp);
#pragma unroll
        for (int a = 0; a < 4; ++a) {
            const int qpos = q0 + ty + 16 * a;
            float mx = -INFINITY;
#pragma unroll
            for (int b = 0; b < 4; ++b) {
                const int kpos = kb * kAlbTile + tx + 16 * b;
                const bool ok = kpos < S && !(g.causal && kpos > qpos);
                s[a][b] = ok ? s[a][b] * g.scale : -INFINITY;
                if (drop.p > 0.f && ok && qpos < S) {
                    const long long tokq = (long long)tok0 + (long long)qpos * g.inner_sz;
                    const uint4 w = drop_words(drop, site, drop_attn_grp(tokq, g.n_head, head, kpos));
                    dp[a][b] *= drop_mul(drop, w, kpos & 7);
                }
                mx = fmaxf(mx, s[a][b]);
            }
            mx = alb_half_max(mx);
            const float mn = fmaxf(m[a], mx);
            const float ms = mn == -INFINITY ? 0.f : mn;      // a row with no visible key yet: everything below evaluates to 0
            const float corr = expf(m[a] - ms);
            float ps = 0.f, ds = 0.f;
#pragma unroll
            for (int b = 0; b < 4; ++b) {
                const float p = expf(s[a][b] - ms);      // exp(-inf) = 0 for masked keys
                ps += p;
                ds = fmaf(p, dp[a][b], ds);
            }
            ps = alb_half_sum(ps);
            ds = alb_half_sum(ds);
            l[a] = l[a] * corr + ps;
            acc[a] = acc[a] * corr + ds;
            m[a] = mn;
        }
    }
    if (tx == 0) {
#pragma unroll
        for (int a = 0; a < 4; ++a) {
            const int qpos = q0 + ty + 16 * a;
            if (qpos < S) {
                const size_t idx = (size_t)blockIdx.x * S + qpos;
                lse[idx] = m[a] + logf(l[a]);
                delta[idx] = acc[a] / l[a];
            }
        }
    }
}

// P, P o Z and dS of one 64 x 64 tile from the two dot-product tiles; rows = queries, columns = keys
template <bool kWantPz>
__device__ __forceinline__ void alb_tile_grads(const AlbGeom& g, const DropCfg& drop, uint32_t site, size_t tok0, int head,
                                               int q0, int k0, int ty, int tx, const float (&lse_r)[4], const float (&del_r)[4],
                                               float (&s)[4][4], float (&dp)[4][4], float* sD, float* sPz) {
#pragma unroll
    for (int a = 0; a < 4; ++a) {
        const int qpos = q0 + ty + 16 * a;
#pragma unroll
        for (int b = 0; b < 4; ++b) {
            const int kpos = k0 + tx + 16 * b;
            const bool ok = qpos < g.S && kpos < g.S && !(g.causal && kpos > qpos);
            float p = ok ? expf(s[a][b] * g.scale - lse_r[a]) : 0.f;
            float z = 1.f;
            if (drop.p > 0.f && ok) {
                const long long tokq = (long long)tok0 + (long long)qpos * g.inner_sz;
                const uint4 w = drop_words(drop, site, drop_attn_grp(tokq, g.n_head, head, kpos));
                z = drop_mul(drop, w, kpos & 7);
            }
            const float ds = p * (dp[a][b] * z - del_r[a]) * g.scale;
            sD[(ty + 16 * a) * kAlbSP + tx + 16 * b] = ds;
            if (kWantPz) sPz[(ty + 16 * a) * kAlbSP + tx + 16 * b] = p * z;
        }
    }
}

template <typename TA, int HD>
__global__ void __launch_bounds__(256) attn_long_dq_kernel(const TA* __restrict__ qkv, const TA* __restrict__ dout,
                                                           TA* __restrict__ dqkv, const float* __restrict__ lse,
                                                           const float* __restrict__ delta, AlbGeom g, DropCfg drop, uint32_t site) {
    extern __shared__ float alb_smem[];
    constexpr int P = HD + 1, DG = HD / 4;
    float* sQ = alb_smem;
    float* sO = sQ + kAlbTile * P;
    float* sK = sO + kAlbTile * P;
    float* sV = sK + kAlbTile * P;
    float* sD = sV + kAlbTile * P;      // [64][65]
    int head; size_t tok0;
    alb_locate(g, head, tok0);
    const int S = g.S, C = g.C, ld = 3 * C;
    const int q0 = blockIdx.y * kAlbTile;
    const int ty = threadIdx.x >> 4, tx = threadIdx.x & 15;
    const TA* qb = qkv + tok0 * ld + head * HD;
    alb_load_tile<TA, HD>(sQ, qb, (size_t)g.inner_sz * ld, q0, S);
    alb_load_tile<TA, HD>(sO, dout + tok0 * C + head * HD, (size_t)g.inner_sz * C, q0, S);
    float lse_r[4], del_r[4];
#pragma unroll
    for (int a = 0; a < 4; ++a) {
        const int qpos = q0 + ty + 16 * a;
        const size_t idx = (size_t)blockIdx.x * S + (qpos < S ? qpos : 0);
        lse_r[a] = lse[idx];
        del_r[a] = delta[idx];
    }
    const int r = threadIdx.x >> 2, dg = threadIdx.x & 3;      // output element group: row r, features dg * DG .. + DG
    float dq[DG];
#pragma unroll
    for (int e = 0; e < DG; ++e) dq[e] = 0.f;
    const int nkb = (S + kAlbTile - 1) / kAlbTile;
    for (int kb = 0; kb < nkb; ++kb) {
        __syncthreads();
        alb_load_tile<TA, HD>(sK, qb + C, (size_t)g.inner_sz * ld, kb * kAlbTile, S);
        alb_load_tile<TA, HD>(sV, qb + 2 * C, (size_t)g.inner_sz * ld, kb * kAlbTile, S);
        __syncthreads();
        float s[4][4], dp[4][4];
        alb_dots<HD>(sQ, sK, ty, tx, s);
        alb_dots<HD>(sO, sV, ty, tx, dp);
        alb_tile_grads<false>(g, drop, site, tok0, head, q0, kb * kAlbTile, ty, tx, lse_r, del_r, s, dp, sD, nullptr);
        __syncthreads();
#pragma unroll 4
        for (int j = 0; j < kAlbTile; ++j) {
            const float dsv = sD[r * kAlbSP + j];
#pragma unroll
            for (int e = 0; e < DG; ++e) dq[e] = fmaf(dsv, sK[j * P + dg * DG + e], dq[e]);
        }
    }
    const int qpos = q0 + r;
    if (qpos < S) {
        TA* dst = dqkv + (tok0 + (size_t)qpos * g.inner_sz) * ld + head * HD + dg * DG;
#pragma unroll
        for (int e = 0; e < DG; e += 4) {
            const float v[4] = {dq[e], dq[e + 1], dq[e + 2], dq[e + 3]};
            Vec4<TA>::store(dst + e, v);
        }
    }
}

template <typename TA, int HD>
__global__ void __launch_bounds__(256) attn_long_dkv_kernel(const TA* __restrict__ qkv, const TA* __restrict__ dout,
                                                            TA* __restrict__ dqkv, const float* __restrict__ lse,
                                                            const float* __restrict__ delta, AlbGeom g, DropCfg drop, uint32_t site) {
    extern __shared__ float alb_smem[];
    constexpr int P = HD + 1, DG = HD / 4;
    float* sQ = alb_smem;
    float* sO = sQ + kAlbTile * P;
    float* sK = sO + kAlbTile * P;
    float* sV = sK + kAlbTile * P;
    float* sD = sV + kAlbTile * P;          // dS  [64 queries][65]
    float* sPz = sD + kAlbTile * kAlbSP;    // P o Z
    int head; size_t tok0;
    alb_locate(g, head, tok0);
    const int S = g.S, C = g.C, ld = 3 * C;
    const int k0 = blockIdx.y * kAlbTile;
    const int ty = threadIdx.x >> 4, tx = threadIdx.x & 15;
    const TA* qb = qkv + tok0 * ld + head * HD;
    alb_load_tile<TA, HD>(sK, qb + C, (size_t)g.inner_sz * ld, k0, S);
    alb_load_tile<TA, HD>(sV, qb + 2 * C, (size_t)g.inner_sz * ld, k0, S);
    const int c = threadIdx.x >> 2, dg = threadIdx.x & 3;      // output element group: key c, features dg * DG .. + DG
    float dk[DG], dv[DG];
#pragma unroll
    for (int e = 0; e < DG; ++e) { dk[e] = 0.f; dv[e] = 0.f; }
    const int nqb = (S + kAlbTile - 1) / kAlbTile;
    for (int qt = 0; qt < nqb; ++qt) {
        const int q0 = qt * kAlbTile;
        if (g.causal && q0 + kAlbTile - 1 < k0) continue;      // every query of the tile precedes every key: all masked
        __syncthreads();
        alb_load_tile<TA, HD>(sQ, qb, (size_t)g.inner_sz * ld, q0, S);
        alb_load_tile<TA, HD>(sO, dout + tok0 * C + head * HD, (size_t)g.inner_sz * C, q0, S);
        float lse_r[4], del_r[4];
#pragma unroll
        for (int a = 0; a < 4; ++a) {
            const int qpos = q0 + ty + 16 * a;
            const size_t idx = (size_t)blockIdx.x * S + (qpos < S ? qpos : 0);
            lse_r[a] = lse[idx];
            del_r[a] = delta[idx];
        }
        __syncthreads();
        float s[4][4], dp[4][4];
        alb_dots<HD>(sQ, sK, ty, tx, s);
        alb_dots<HD>(sO, sV, ty, tx, dp);
        alb_tile_grads<true>(g, drop, site, tok0, head, q0, k0, ty, tx, lse_r, del_r, s, dp, sD, sPz);
        __syncthreads();
#pragma unroll 4
        for (int i = 0; i < kAlbTile; ++i) {
            const float dsv = sD[i * kAlbSP + c], pz = sPz[i * kAlbSP + c];
#pragma unroll
            for (int e = 0; e < DG; ++e) {
                dk[e] = fmaf(dsv, sQ[i * P + dg * DG + e], dk[e]);
                dv[e] = fmaf(pz, sO[i * P + dg * DG + e], dv[e]);
            }
        }
    }
    const int kpos = k0 + c;
    if (kpos < S) {
        TA* dst = dqkv + (tok0 + (size_t)kpos * g.inner_sz) * ld + C + head * HD + dg * DG;
#pragma unroll
        for (int e = 0; e < DG; e += 4) {
            const float kv[4] = {dk[e], dk[e + 1], dk[e + 2], dk[e + 3]};
            const float vv[4] = {dv[e], dv[e + 1], dv[e + 2], dv[e + 3]};
            Vec4<TA>::store(dst + e, kv);
            Vec4<TA>::store(dst + C + e, vv);
        }
    }
}

// Host launcher.  stats: 2 * n_seq * n_head * S floats of scratch (LSE, delta).  Returns false when the shape is not covered
// (head_dim other than 16 / 32 / 64, grids beyond the launch limits).
template <typename TA>
static bool launch_attention_long_bwd(const TA* qkv, const TA* dout, TA* dqkv, float* stats, long long n_seq, int S, int inner_sz,
                                      int n_head, int C, int head_dim, int causal, cudaStream_t st, cudaError_t* err,
                                      const DropCfg& drop = DropCfg(), uint32_t site = 0) {
    const long long gx = n_seq * n_head;
    const int tiles = (S + kAlbTile - 1) / kAlbTile;
    if (gx > 0x7fffffffLL || tiles > 65535 || (head_dim != 16 && head_dim != 32 && head_dim != 64)) return false;
    AlbGeom g{S, inner_sz, n_head, C, causal, 1.0f / sqrtf((float)head_dim)};
    float* lse = stats;
    float* delta = stats + (size_t)gx * S;
    const dim3 grid((unsigned)gx, (unsigned)tiles);
    cudaError_t e = cudaSuccess;
#define TANTE_ALB(HDv)                                                                                                            \
    do {                                                                                                                          \
        const size_t tile = (size_t)4 * kAlbTile * (HDv + 1) * sizeof(float), sc = (size_t)kAlbTile * kAlbSP * sizeof(float);    \
        e = cudaFuncSetAttribute(attn_long_stats_kernel<TA, HDv>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)tile);       \
        if (e == cudaSuccess) e = cudaFuncSetAttribute(attn_long_dq_kernel<TA, HDv>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(tile + sc)); \
        if (e == cudaSuccess) e = cudaFuncSetAttribute(attn_long_dkv_kernel<TA, HDv>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(tile + 2 * sc)); \
        if (e != cudaSuccess) break;                                                                                              \
        attn_long_stats_kernel<TA, HDv><<<grid, 256, tile, st>>>(qkv, dout, lse, delta, g, drop, site);                           \
        attn_long_dq_kernel<TA, HDv><<<grid, 256, tile + sc, st>>>(qkv, dout, dqkv, lse, delta, g, drop, site);                   \
        attn_long_dkv_kernel<TA, HDv><<<grid, 256, tile + 2 * sc, st>>>(qkv, dout, dqkv, lse, delta, g, drop, site);              \
        e = cudaGetLastError();                                                                                                   \
    } while (0)
    if (head_dim == 32) TANTE_ALB(32); else if (head_dim == 64) TANTE_ALB(64); else TANTE_ALB(16);
#undef TANTE_ALB
    *err = e;
    return true;
}

}  // namespace tante
