// Attention axis 'C' of Attn_Backbone (reference models/attn_backbone.py:124-130,184-189): every latent token (b, t, h, w)
// becomes a SEQUENCE of its C channel values; each scalar is lifted to E = expanded_channel features by
// channel_blocks[j] = Linear(1, E/4) -> GELU(erf) -> Linear(E/4, E), a TransformerBlock(E) attends over the C channel
// tokens, and the LAST feature of its output is the new latent value:
//     x = rearrange(x, 'b t h w c -> (b t h w) c 1'); x = channel_blocks[j](x); x = blocks[i](x)[..., -1]
// The block itself runs on the existing LayerNorm / GEMM / attention kernels at width E over rows = tokens * C
// (tante_abi.cu: run_channel_layer, in chunks of latent tokens); this file holds the two ends of the pass.  Training recomputes
// the block chunk by chunk in the backward (only the layer's fp32 input is kept) and runs the usual backward kernels at width E.
#pragma once
#include "common.cuh"

namespace tante {

// rows = (token, channel) pairs in latent order: row r reads the scalar x[r] and writes xc[r][0..E).  One warp per row at a
// time: lane j evaluates hidden unit(s) j, j + 32 (E/4 <= 64), the hidden vector is exchanged through shared memory, then
// every lane produces the features e = lane, lane + 32, ... (coalesced 128-byte stores).  W2 is staged transposed,
// [E/4][E], so that the dot products read consecutive shared-memory words across the warp.
__global__ void __launch_bounds__(256) channel_lift_kernel(const float* __restrict__ x, const float* __restrict__ w0,
                                                           const float* __restrict__ b0, const float* __restrict__ w2,
                                                           const float* __restrict__ b2, float* __restrict__ xc, long long rows,
                                                           int E) {
    extern __shared__ float cl_smem[];
    const int E4 = E / 4;
    float* sW2 = cl_smem;                    // [E4][E]
    float* sB2 = sW2 + (size_t)E4 * E;       // [E]
    float* sW0 = sB2 + E;                    // [E4]
    float* sB0 = sW0 + E4;                   // [E4]
    float* sH = sB0 + E4;                    // [8 warps][E4]
    for (int i = threadIdx.x; i < E4 * E; i += blockDim.x) {
        const int e = i / E4, j = i % E4;    // w2 is [E][E4] row-major (nn.Linear weight)
        sW2[j * E + e] = w2[i];
    }
    for (int i = threadIdx.x; i < E; i += blockDim.x) sB2[i] = b2[i];
    for (int i = threadIdx.x; i < E4; i += blockDim.x) { sW0[i] = w0[i]; sB0[i] = b0[i]; }
    __syncthreads();
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    float* hv = sH + warp * E4;
    const long long nw = (long long)gridDim.x * 8;
    for (long long r = (long long)blockIdx.x * 8 + warp; r < rows; r += nw) {
        const float v = x[r];
        for (int j = lane; j < E4; j += 32) {
            const float a = fmaf(v, sW0[j], sB0[j]);
            hv[j] = 0.5f * a * (1.0f + erff(a * 0.70710678118654752f));
        }
        __syncwarp();
        float* dst = xc + (size_t)r * E;
        for (int e = lane; e < E; e += 32) {
            float acc = sB2[e];
            for (int j = 0; j < E4; ++j) acc = fmaf(hv[j], sW2[j * E + e], acc);
            dst[e] = acc;
        }
        __syncwarp();
    }
}

// x[r] = xc[r][E - 1]: the last feature of the block output is the new latent value (attn_backbone.py:188).
__global__ void __launch_bounds__(256) channel_extract_kernel(const float* __restrict__ xc, float* __restrict__ x, long long rows,
                                                              int E) {
    const long long r = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (r < rows) x[r] = xc[(size_t)r * E + (E - 1)];
}

// ---- training ----------------------------------------------------------------------------------------------------------------
// gradient of the block output: only the last feature carries the gradient of the new latent value (fp32 stream + TA mirror)
template <typename TA>
__global__ void __launch_bounds__(256) channel_extract_bwd_kernel(const float* __restrict__ gx, float* __restrict__ gxc, TA* __restrict__ gxb,
                                                                  long long rows, int E) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;      // element of [rows][E]
    if (i >= rows * E) return;
    const long long r = i / E;
    const int e = (int)(i % E);
    const float v = e == E - 1 ? gx[r] : 0.f;
    gxc[i] = v;
    if (gxb) gxb[i] = from_f32<TA>(v);
}

// backward of channel_lift_kernel for one chunk: g = gradient of the lifted features [rows][E] (fp32).  One warp per row at a time:
//   dh[j] = sum_e W2[e][j] g[e];  da[j] = dh[j] gelu'(a[j]), a[j] = x w0[j] + b0[j];  dx = sum_j da[j] w0[j]  (-> gx[row], overwritten)
//   gw0[j] += da[j] x, gb0[j] += da[j]  (per-lane registers over the rows of the warp, one atomic per entry and warp at the end)
// and stores the hidden activations h[row][0 .. E/4) (TA, row pitch ldh) for the weight-gradient GEMM  gW2 = g^T h,  gb2 = colsum(g).
template <typename TA>
__global__ void __launch_bounds__(256) channel_lift_bwd_kernel(const float* __restrict__ x, const float* __restrict__ g,
                                                               const float* __restrict__ w0, const float* __restrict__ b0,
                                                               const float* __restrict__ w2, float* __restrict__ gx, TA* __restrict__ hh,
                                                               int ldh, float* __restrict__ gw0, float* __restrict__ gb0, long long rows,
                                                               int E) {
    extern __shared__ float cl_smem[];
    const int E4 = E / 4;
    float* sW2 = cl_smem;                    // [E][E4] (as stored: lanes read consecutive j)
    float* sW0 = sW2 + (size_t)E * E4;       // [E4]
    float* sB0 = sW0 + E4;                   // [E4]
    float* sG = sB0 + E4;                    // [8 warps][E]
    for (int i = threadIdx.x; i < E * E4; i += blockDim.x) sW2[i] = w2[i];
    for (int i = threadIdx.x; i < E4; i += blockDim.x) { sW0[i] = w0[i]; sB0[i] = b0[i]; }
    __syncthreads();
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    float* gv = sG + warp * E;
    float aw0[2] = {0.f, 0.f}, ab0[2] = {0.f, 0.f};      // E4 <= 64: entries lane, lane + 32
    const long long nw = (long long)gridDim.x * 8;
    for (long long r = (long long)blockIdx.x * 8 + warp; r < rows; r += nw) {
        const float xv = x[r];
        for (int e = lane; e < E; e += 32) gv[e] = g[(size_t)r * E + e];
        __syncwarp();
        float dx = 0.f;
#pragma unroll
        for (int q = 0; q < 2; ++q) {
            const int j = lane + 32 * q;
            if (j < E4) {
                float dh = 0.f;
                for (int e = 0; e < E; ++e) dh = fmaf(sW2[e * E4 + j], gv[e], dh);
                const float a = fmaf(xv, sW0[j], sB0[j]);
                const float cdf = 0.5f * (1.0f + erff(a * 0.70710678118654752f));
                const float pdf = 0.3989422804014327f * expf(-0.5f * a * a);
                const float da = dh * (cdf + a * pdf);
                hh[(size_t)r * ldh + j] = from_f32<TA>(a * cdf);
                aw0[q] = fmaf(da, xv, aw0[q]);
                ab0[q] += da;
                dx = fmaf(da, sW0[j], dx);
            }
        }
        dx = warp_sum(dx);
        if (lane == 0) gx[r] = dx;
        __syncwarp();
    }
#pragma unroll
    for (int q = 0; q < 2; ++q) {
        const int j = lane + 32 * q;
        if (j < E4) { atomicAdd(gw0 + j, aw0[q]); atomicAdd(gb0 + j, ab0[q]); }
    }
}

}  // namespace tante
