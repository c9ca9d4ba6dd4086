// Attention axis 'C' of Attn_Backbone (reference models/attn_backbone.py:124-130,184-189): every latent token (b, t, h, w)
// becomes a SEQUENCE of its C channel values; each scalar is lifted to E = expanded_channel features by
// channel_blocks[j] = Linear(1, E/4) -> GELU(erf) -> Linear(E/4, E), a TransformerBlock(E) attends over the C channel
// tokens, and the LAST feature of its output is the new latent value:
//     x = rearrange(x, 'b t h w c -> (b t h w) c 1'); x = channel_blocks[j](x); x = blocks[i](x)[..., -1]
// The block itself runs on the existing LayerNorm / GEMM / attention kernels at width E over rows = tokens * C
// (tante_abi.cu: run_channel_layer, in chunks of latent tokens); this file holds the two ends of the pass.
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

}  // namespace tante
