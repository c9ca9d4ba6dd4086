// Residual axis MLP ("propagator", attn_backbone.py:111-119,140-146) along a SHORT axis (S <= 8, e.g. the T axis of
// tante.yaml with in_T = 4), forward and backward.  y = x + W2 gelu(W1 x + b1) + b2 along the axis.
// With S this small the op is a per-column S x S mat-vec pair: one thread owns 4 adjacent columns (128-bit accesses)
// and all S axis positions in registers, so the kernels stream the latent exactly once (read x [+ dy], write y / dx)
// with every warp access covering whole 512-byte runs.  Element (outer, p, col) lives at (outer*S + p)*IC + col.
// TM = float: accurate erf/exp (exact mode); TM = bf16: the approximate versions of the tensor mode.
#pragma once
#include "common.cuh"

namespace tante {

template <typename TM, int SMAX>
__global__ void __launch_bounds__(256) propagator_small_kernel(const float* xin, float* x, int S, long long IC, long long n_outer,
                                                               const float* __restrict__ W1, const float* __restrict__ b1,
                                                               const float* __restrict__ W2, const float* __restrict__ b2) {
    __shared__ float sw1[SMAX * SMAX], sw2[SMAX * SMAX], sb1[SMAX], sb2[SMAX];
    for (int i = threadIdx.x; i < SMAX * SMAX; i += blockDim.x) {
        const int a = i / SMAX, b = i % SMAX;
        const bool ok = a < S && b < S;
        sw1[i] = ok ? W1[a * S + b] : 0.f;
        sw2[i] = ok ? W2[a * S + b] : 0.f;
    }
    for (int i = threadIdx.x; i < SMAX; i += blockDim.x) { sb1[i] = i < S ? b1[i] : 0.f; sb2[i] = i < S ? b2[i] : 0.f; }
    __syncthreads();
    const long long ic4 = IC / 4;
    const long long total = n_outer * ic4;
    for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (long long)gridDim.x * blockDim.x) {
        const long long outer = idx / ic4, c4 = idx % ic4;
        const size_t base = (size_t)outer * S * IC + (size_t)c4 * 4;
        float4 v[SMAX];
#pragma unroll
        for (int p = 0; p < SMAX; ++p) v[p] = p < S ? *reinterpret_cast<const float4*>(xin + base + (size_t)p * IC) : make_float4(0.f, 0.f, 0.f, 0.f);
        float4 hd[SMAX];
#pragma unroll
        for (int j = 0; j < SMAX; ++j) {
            float4 a = make_float4(sb1[j], sb1[j], sb1[j], sb1[j]);
#pragma unroll
            for (int i = 0; i < SMAX; ++i) {
                const float w = sw1[j * SMAX + i];
                a.x = fmaf(w, v[i].x, a.x); a.y = fmaf(w, v[i].y, a.y); a.z = fmaf(w, v[i].z, a.z); a.w = fmaf(w, v[i].w, a.w);
            }
            hd[j] = make_float4(ActMath<TM>::gelu_erf_f(a.x), ActMath<TM>::gelu_erf_f(a.y), ActMath<TM>::gelu_erf_f(a.z),
                                ActMath<TM>::gelu_erf_f(a.w));
        }
#pragma unroll
        for (int p = 0; p < SMAX; ++p) {
            if (p >= S) continue;
            float4 a = make_float4(sb2[p], sb2[p], sb2[p], sb2[p]);
#pragma unroll
            for (int j = 0; j < SMAX; ++j) {
                const float w = sw2[p * SMAX + j];
                a.x = fmaf(w, hd[j].x, a.x); a.y = fmaf(w, hd[j].y, a.y); a.z = fmaf(w, hd[j].z, a.z); a.w = fmaf(w, hd[j].w, a.w);
            }
            *reinterpret_cast<float4*>(x + base + (size_t)p * IC) = make_float4(v[p].x + a.x, v[p].y + a.y, v[p].z + a.z, v[p].w + a.w);
        }
    }
}

// In place on the gradient stream: dx = dy + W1^T (gelu'(pre) o (W2^T dy)); weight / bias gradients are accumulated in
// registers over the grid-stride loop, reduced by warp shuffles + shared-memory atomics and flushed with one global
// atomic per entry per CTA.
template <typename TM, int SMAX>
__global__ void __launch_bounds__(256) propagator_small_bwd_kernel(const float* __restrict__ xin, float* __restrict__ dy, int S,
                                                                   long long IC, long long n_outer,
                                                                   const float* __restrict__ W1, const float* __restrict__ b1,
                                                                   const float* __restrict__ W2, float* __restrict__ gW1,
                                                                   float* __restrict__ gb1, float* __restrict__ gW2,
                                                                   float* __restrict__ gb2) {
    __shared__ float sw1[SMAX * SMAX], sw2[SMAX * SMAX], sb1[SMAX];
    __shared__ float rg1[SMAX * SMAX], rg2[SMAX * SMAX], rb1[SMAX], rb2[SMAX];
    for (int i = threadIdx.x; i < SMAX * SMAX; i += blockDim.x) {
        const int a = i / SMAX, b = i % SMAX;
        const bool ok = a < S && b < S;
        sw1[i] = ok ? W1[a * S + b] : 0.f;
        sw2[i] = ok ? W2[a * S + b] : 0.f;
        rg1[i] = rg2[i] = 0.f;
    }
    for (int i = threadIdx.x; i < SMAX; i += blockDim.x) { sb1[i] = i < S ? b1[i] : 0.f; rb1[i] = rb2[i] = 0.f; }
    __syncthreads();
    float aw1[SMAX][SMAX], aw2[SMAX][SMAX], ab1[SMAX], ab2[SMAX];
#pragma unroll
    for (int j = 0; j < SMAX; ++j) {
        ab1[j] = ab2[j] = 0.f;
#pragma unroll
        for (int i = 0; i < SMAX; ++i) aw1[j][i] = aw2[j][i] = 0.f;
    }
    const long long ic4 = IC / 4;
    const long long total = n_outer * ic4;
    for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (long long)gridDim.x * blockDim.x) {
        const long long outer = idx / ic4, c4 = idx % ic4;
        const size_t base = (size_t)outer * S * IC + (size_t)c4 * 4;
        float xv[SMAX][4], dv[SMAX][4];
#pragma unroll
        for (int p = 0; p < SMAX; ++p) {
            float4 a = make_float4(0.f, 0.f, 0.f, 0.f), b = a;
            if (p < S) {
                a = *reinterpret_cast<const float4*>(xin + base + (size_t)p * IC);
                b = *reinterpret_cast<const float4*>(dy + base + (size_t)p * IC);
            }
            xv[p][0] = a.x; xv[p][1] = a.y; xv[p][2] = a.z; xv[p][3] = a.w;
            dv[p][0] = b.x; dv[p][1] = b.y; dv[p][2] = b.z; dv[p][3] = b.w;
        }
        float dp[SMAX][4];      // dpre
#pragma unroll
        for (int j = 0; j < SMAX; ++j) {
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                float pre = sb1[j], dh = 0.f;
#pragma unroll
                for (int i = 0; i < SMAX; ++i) {
                    pre = fmaf(sw1[j * SMAX + i], xv[i][e], pre);
                    dh = fmaf(sw2[i * SMAX + j], dv[i][e], dh);           // dh[j] = sum_p W2[p][j] dy[p]
                }
                const float hj = ActMath<TM>::gelu_erf_f(pre);
                dp[j][e] = dh * ActMath<TM>::gelu_erf_g(pre);
#pragma unroll
                for (int p = 0; p < SMAX; ++p) aw2[p][j] = fmaf(dv[p][e], hj, aw2[p][j]);   // gW2[p][j] += dy[p] h[j]
            }
        }
#pragma unroll
        for (int j = 0; j < SMAX; ++j) {
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                ab1[j] += dp[j][e];
                ab2[j] += dv[j][e];
#pragma unroll
                for (int i = 0; i < SMAX; ++i) aw1[j][i] = fmaf(dp[j][e], xv[i][e], aw1[j][i]);  // gW1[j][i] += dpre[j] x[i]
            }
        }
#pragma unroll
        for (int p = 0; p < SMAX; ++p) {
            if (p >= S) continue;
            float o[4];
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                float a = dv[p][e];
#pragma unroll
                for (int j = 0; j < SMAX; ++j) a = fmaf(sw1[j * SMAX + p], dp[j][e], a);   // dx[p] = dy[p] + sum_j W1[j][p] dpre[j]
                o[e] = a;
            }
            *reinterpret_cast<float4*>(dy + base + (size_t)p * IC) = make_float4(o[0], o[1], o[2], o[3]);
        }
    }
    const int lane = threadIdx.x % 32;
#pragma unroll
    for (int j = 0; j < SMAX; ++j) {
#pragma unroll
        for (int i = 0; i < SMAX; ++i) {
            const float s1 = warp_sum(aw1[j][i]), s2 = warp_sum(aw2[j][i]);
            if (lane == 0) { atomicAdd(&rg1[j * SMAX + i], s1); atomicAdd(&rg2[j * SMAX + i], s2); }
        }
        const float t1 = warp_sum(ab1[j]), t2 = warp_sum(ab2[j]);
        if (lane == 0) { atomicAdd(&rb1[j], t1); atomicAdd(&rb2[j], t2); }
    }
    __syncthreads();
    for (int e = threadIdx.x; e < S * S; e += blockDim.x) {
        const int j = e / S, i = e % S;
        atomicAdd(gW1 + e, rg1[j * SMAX + i]);
        atomicAdd(gW2 + e, rg2[j * SMAX + i]);
    }
    for (int j = threadIdx.x; j < S; j += blockDim.x) { atomicAdd(gb1 + j, rb1[j]); atomicAdd(gb2 + j, rb2[j]); }
}

}  // namespace tante
