// Backward of the residual axis MLP ("propagator", attn_backbone.py:111-119,140-146) on tensor cores (bf16 mode):
//   y = x + W2 gelu(W1 x + b1) + b2   along one axis of length S <= 64
//   dh   = W2^T dy,  dpre = dh o gelu'(pre)  (pre recomputed from x, as the forward computed it: bf16 operands)
//   dx   = dy + W1^T dpre                    (in place on the fp32 gradient stream)
//   gW2 += dy h^T, gb2 += rowsum(dy), gW1 += dpre x^T, gb1 += rowsum(dpre)
// Same slab geometry as propagator_mma_kernel: one slab = one `outer` x 128 columns, 4 warps x 32 columns for the
// three slab GEMMs (weights = A operands, slab = B operand through ldmatrix.trans).  The weight gradients contract
// over the COLUMNS of the slab: both operands are slab rows (k contiguous), read with plain ldmatrix; every warp owns
// one 16-row block of gW1 / gW2 (x a share of the 128 columns when S <= 32), keeps it in registers across the slabs
// of its persistent CTA and flushes it with one atomic per entry at the end.  The bias gradients ride along as one
// extra MMA against a fragment of ones.  HBM traffic: read x, read dy, write dx (3 x 4 B per element).
#pragma once
#include "propagator_mma.cuh"

namespace tante {

template <int MB /* padded S / 16: 1, 2 or 4 */>
__global__ void __launch_bounds__(128) propagator_bwd_mma_kernel(const float* __restrict__ xin, float* __restrict__ dy, int S,
                                                                 long long IC, long long n_outer,
                                                                 const float* __restrict__ W1, const float* __restrict__ b1,
                                                                 const float* __restrict__ W2, float* __restrict__ gW1,
                                                                 float* __restrict__ gb1, float* __restrict__ gW2,
                                                                 float* __restrict__ gb2) {
    constexpr int SP = MB * 16;
    constexpr int WPITCH = SP * 2 + 16;          // bytes; odd number of 16-B chunks -> conflict-free ldmatrix
    constexpr int KQ = 4 / MB;                   // warps sharing one 16-row block of the weight gradients
    constexpr int KSTEPS = 8 / KQ;               // 16-column k-steps per warp in the weight-gradient phase
    extern __shared__ __align__(128) uint8_t pb_smem[];
    uint8_t* sX = pb_smem;                       // x      [SP][128] bf16 swizzled
    uint8_t* sD = sX + SP * 256;                 // dy
    uint8_t* sH = sD + SP * 256;                 // h = gelu(pre)
    uint8_t* sP = sH + SP * 256;                 // dpre
    uint8_t* sW1 = sP + SP * 256;                // W1[j][i]
    uint8_t* sW2T = sW1 + SP * WPITCH;           // W2^T[i][j]
    uint8_t* sW1T = sW2T + SP * WPITCH;          // W1^T[i][j]
    float* sb1 = reinterpret_cast<float*>(sW1T + SP * WPITCH);

    const int tid = threadIdx.x, warp = tid / 32, lane = tid % 32;
    for (int i = tid; i < SP * SP; i += 128) {
        const int a = i / SP, b = i % SP;
        const bool ok = a < S && b < S;
        *reinterpret_cast<__nv_bfloat16*>(sW1 + a * WPITCH + b * 2) = __float2bfloat16_rn(ok ? W1[a * S + b] : 0.f);
        *reinterpret_cast<__nv_bfloat16*>(sW2T + a * WPITCH + b * 2) = __float2bfloat16_rn(ok ? W2[b * S + a] : 0.f);
        *reinterpret_cast<__nv_bfloat16*>(sW1T + a * WPITCH + b * 2) = __float2bfloat16_rn(ok ? W1[b * S + a] : 0.f);
    }
    for (int i = tid; i < SP; i += 128) sb1[i] = i < S ? b1[i] : 0.f;

    const uint32_t aX = (uint32_t)__cvta_generic_to_shared(sX), aD = (uint32_t)__cvta_generic_to_shared(sD);
    const uint32_t aH = (uint32_t)__cvta_generic_to_shared(sH), aP = (uint32_t)__cvta_generic_to_shared(sP);
    const uint32_t aW1 = (uint32_t)__cvta_generic_to_shared(sW1), aW2T = (uint32_t)__cvta_generic_to_shared(sW2T);
    const uint32_t aW1T = (uint32_t)__cvta_generic_to_shared(sW1T);
    const int g = lane >> 2, t = lane & 3;
    const int lrow = (lane & 7) + 8 * ((lane >> 3) & 1);
    const int lchk = lane >> 4;
    const int wmb = warp % MB, wkq = warp / MB;   // weight-gradient role of this warp

    float aw1[2 * MB][4], aw2[2 * MB][4], ab1[4], ab2[4];
#pragma unroll
    for (int i = 0; i < 2 * MB; ++i)
#pragma unroll
        for (int e = 0; e < 4; ++e) aw1[i][e] = aw2[i][e] = 0.f;
#pragma unroll
    for (int e = 0; e < 4; ++e) ab1[e] = ab2[e] = 0.f;

    const long long ncb = (IC + 127) / 128;
    const long long nslab = n_outer * ncb;
    for (long long slab = blockIdx.x; slab < nslab; slab += gridDim.x) {
        const long long outer = slab / ncb;
        const long long col0 = (slab % ncb) * 128;
        const int ncol = (int)min((long long)128, IC - col0);
        const float* xb = xin + (size_t)outer * S * IC + col0;
        float* db = dy + (size_t)outer * S * IC + col0;
        __syncthreads();                          // weight-gradient phase of the previous slab is done with the slabs
#pragma unroll 4
        for (int i = tid; i < SP * 32; i += 128) {
            const int p = i / 32, c4 = (i % 32) * 4;
            float4 xv = make_float4(0.f, 0.f, 0.f, 0.f), dv = xv;
            if (p < S && c4 < ncol) {
                xv = *reinterpret_cast<const float4*>(xb + (size_t)p * IC + c4);
                dv = *reinterpret_cast<const float4*>(db + (size_t)p * IC + c4);
            }
            const uint32_t off = slab_off(p, c4 / 8) + (c4 % 8) * 2;
            *reinterpret_cast<uint2*>(sX + off) = make_uint2(pack_bf16x2(xv.x, xv.y), pack_bf16x2(xv.z, xv.w));
            *reinterpret_cast<uint2*>(sD + off) = make_uint2(pack_bf16x2(dv.x, dv.y), pack_bf16x2(dv.z, dv.w));
        }
        __syncthreads();

        // ---- pre = W1 x + b1 and dh = W2^T dy, one 16-row block of hidden units at a time ----
#pragma unroll
        for (int mb = 0; mb < MB; ++mb) {
            float accp[4][4], acch[4][4];
#pragma unroll
            for (int nb = 0; nb < 4; ++nb)
#pragma unroll
                for (int e = 0; e < 4; ++e) accp[nb][e] = acch[nb][e] = 0.f;
#pragma unroll
            for (int kk = 0; kk < MB; ++kk) {
                uint32_t bx[4][2], bd[4][2];
#pragma unroll
                for (int j = 0; j < 2; ++j) {
                    const uint32_t o = slab_off(kk * 16 + lrow, warp * 4 + 2 * j + lchk);
                    ldsm_x4_t(aX + o, bx[2 * j][0], bx[2 * j][1], bx[2 * j + 1][0], bx[2 * j + 1][1]);
                    ldsm_x4_t(aD + o, bd[2 * j][0], bd[2 * j][1], bd[2 * j + 1][0], bd[2 * j + 1][1]);
                }
                uint32_t a1[4], a2[4];
                const uint32_t wo = (uint32_t)((mb * 16 + lrow) * WPITCH + (kk * 2 + lchk) * 16);
                ldsm_x4(aW1 + wo, a1[0], a1[1], a1[2], a1[3]);
                ldsm_x4(aW2T + wo, a2[0], a2[1], a2[2], a2[3]);
#pragma unroll
                for (int nb = 0; nb < 4; ++nb) {
                    mma_bf16_16816(accp[nb], a1, bx[nb][0], bx[nb][1]);
                    mma_bf16_16816(acch[nb], a2, bd[nb][0], bd[nb][1]);
                }
            }
            const int j0 = mb * 16 + g, j1 = j0 + 8;
            const float bb0 = sb1[j0], bb1 = sb1[j1];
#pragma unroll
            for (int nb = 0; nb < 4; ++nb) {
                const int chunk = warp * 4 + nb;
                const float p0 = accp[nb][0] + bb0, p1 = accp[nb][1] + bb0, p2 = accp[nb][2] + bb1, p3 = accp[nb][3] + bb1;
                *reinterpret_cast<uint32_t*>(sH + slab_off(j0, chunk) + t * 4) = pack_bf16x2(gelu_erf_fast(p0), gelu_erf_fast(p1));
                *reinterpret_cast<uint32_t*>(sH + slab_off(j1, chunk) + t * 4) = pack_bf16x2(gelu_erf_fast(p2), gelu_erf_fast(p3));
                *reinterpret_cast<uint32_t*>(sP + slab_off(j0, chunk) + t * 4) =
                    pack_bf16x2(acch[nb][0] * gelu_erf_grad_fast(p0), acch[nb][1] * gelu_erf_grad_fast(p1));
                *reinterpret_cast<uint32_t*>(sP + slab_off(j1, chunk) + t * 4) =
                    pack_bf16x2(acch[nb][2] * gelu_erf_grad_fast(p2), acch[nb][3] * gelu_erf_grad_fast(p3));
            }
        }
        __syncwarp();       // dpre columns of a warp are consumed by the same warp below

        // ---- dx = dy + W1^T dpre ----
        {
            float acc[MB][4][4];
#pragma unroll
            for (int mb = 0; mb < MB; ++mb)
#pragma unroll
                for (int nb = 0; nb < 4; ++nb) acc[mb][nb][0] = acc[mb][nb][1] = acc[mb][nb][2] = acc[mb][nb][3] = 0.f;
#pragma unroll
            for (int kk = 0; kk < MB; ++kk) {
                uint32_t bp[4][2];
#pragma unroll
                for (int j = 0; j < 2; ++j)
                    ldsm_x4_t(aP + slab_off(kk * 16 + lrow, warp * 4 + 2 * j + lchk), bp[2 * j][0], bp[2 * j][1], bp[2 * j + 1][0],
                              bp[2 * j + 1][1]);
#pragma unroll
                for (int mb = 0; mb < MB; ++mb) {
                    uint32_t af[4];
                    ldsm_x4(aW1T + (uint32_t)((mb * 16 + lrow) * WPITCH + (kk * 2 + lchk) * 16), af[0], af[1], af[2], af[3]);
#pragma unroll
                    for (int nb = 0; nb < 4; ++nb) mma_bf16_16816(acc[mb][nb], af, bp[nb][0], bp[nb][1]);
                }
            }
            // read-modify-write of the gradient stream in two sweeps (all loads first: a load may not be hoisted over
            // a store to the same array, so interleaving them would serialise 8*MB L2 round trips per thread)
#pragma unroll
            for (int mb = 0; mb < MB; ++mb) {
#pragma unroll
                for (int hh = 0; hh < 2; ++hh) {
                    const int p = mb * 16 + g + 8 * hh;
#pragma unroll
                    for (int nb = 0; nb < 4; ++nb) {
                        const int c = warp * 32 + nb * 8 + 2 * t;
                        if (p < S && c < ncol) {
                            const float2 dv = *reinterpret_cast<const float2*>(db + (size_t)p * IC + c);
                            acc[mb][nb][2 * hh + 0] += dv.x;
                            acc[mb][nb][2 * hh + 1] += dv.y;
                        }
                    }
                }
            }
#pragma unroll
            for (int mb = 0; mb < MB; ++mb) {
#pragma unroll
                for (int hh = 0; hh < 2; ++hh) {
                    const int p = mb * 16 + g + 8 * hh;
#pragma unroll
                    for (int nb = 0; nb < 4; ++nb) {
                        const int c = warp * 32 + nb * 8 + 2 * t;
                        if (p < S && c < ncol)
                            *reinterpret_cast<float2*>(db + (size_t)p * IC + c) = make_float2(acc[mb][nb][2 * hh], acc[mb][nb][2 * hh + 1]);
                    }
                }
            }
        }
        __syncthreads();    // every column of h / dpre is in shared memory

        // ---- weight gradients: rows wmb*16.. of gW2 = dy h^T and gW1 = dpre x^T over this warp's column share ----
#pragma unroll
        for (int ks = 0; ks < KSTEPS; ++ks) {
            const int kc = (wkq * KSTEPS + ks) * 2 + lchk;
            uint32_t ad[4], ap[4];
            ldsm_x4(aD + slab_off(wmb * 16 + lrow, kc), ad[0], ad[1], ad[2], ad[3]);
            ldsm_x4(aP + slab_off(wmb * 16 + lrow, kc), ap[0], ap[1], ap[2], ap[3]);
#pragma unroll
            for (int nbp = 0; nbp < MB; ++nbp) {
                uint32_t r0, r1, r2, r3;
                ldsm_x4(aH + slab_off(nbp * 16 + lrow, kc), r0, r1, r2, r3);
                mma_bf16_16816(aw2[2 * nbp], ad, r0, r2);
                mma_bf16_16816(aw2[2 * nbp + 1], ad, r1, r3);
                ldsm_x4(aX + slab_off(nbp * 16 + lrow, kc), r0, r1, r2, r3);
                mma_bf16_16816(aw1[2 * nbp], ap, r0, r2);
                mma_bf16_16816(aw1[2 * nbp + 1], ap, r1, r3);
            }
            const uint32_t ones = 0x3F803F80u;    // bf16 (1, 1)
            mma_bf16_16816(ab2, ad, ones, ones);
            mma_bf16_16816(ab1, ap, ones, ones);
        }
    }

    // ---- flush: this warp holds rows wmb*16 + g (+8), columns nb*8 + 2t (+1) of both gradients ----
#pragma unroll
    for (int hh = 0; hh < 2; ++hh) {
        const int j = wmb * 16 + g + 8 * hh;
        if (j >= S) continue;
#pragma unroll
        for (int nb = 0; nb < 2 * MB; ++nb) {
#pragma unroll
            for (int e = 0; e < 2; ++e) {
                const int i = nb * 8 + 2 * t + e;
                if (i < S) {
                    atomicAdd(gW1 + j * S + i, aw1[nb][2 * hh + e]);
                    atomicAdd(gW2 + j * S + i, aw2[nb][2 * hh + e]);
                }
            }
        }
        if (t == 0) {
            atomicAdd(gb1 + j, ab1[2 * hh]);
            atomicAdd(gb2 + j, ab2[2 * hh]);
        }
    }
}

template <int MB>
static cudaError_t launch_propagator_bwd_mma_inst(const float* xin, float* dy, int S, long long IC, long long outer,
                                                  const float* W1, const float* b1, const float* W2, float* gW1, float* gb1,
                                                  float* gW2, float* gb2, int num_sms, cudaStream_t st) {
    constexpr int SP = MB * 16;
    const size_t smem = (size_t)4 * SP * 256 + (size_t)3 * SP * (SP * 2 + 16) + SP * sizeof(float);
    static unsigned long long attr_done = 0;
    if (attrs_needed(attr_done)) {
        cudaError_t e = cudaFuncSetAttribute(propagator_bwd_mma_kernel<MB>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
    }
    const long long nslab = outer * ((IC + 127) / 128);
    const unsigned grid = (unsigned)std::min<long long>(nslab, (long long)num_sms * (MB == 4 ? 2 : 4));
    propagator_bwd_mma_kernel<MB><<<grid, 128, smem, st>>>(xin, dy, S, IC, outer, W1, b1, W2, gW1, gb1, gW2, gb2);
    return cudaGetLastError();
}

// Returns false for axis lengths the tensor-core kernel does not pay off for (short axes stay on the FFMA kernel).
static bool launch_propagator_bwd_mma(const float* xin, float* dy, int S, long long IC, long long outer, const float* W1,
                                      const float* b1, const float* W2, float* gW1, float* gb1, float* gW2, float* gb2,
                                      int num_sms, cudaStream_t st, cudaError_t* err) {
    if (S <= 8 || S > 64 || (IC & 3)) return false;
    if (S <= 16) *err = launch_propagator_bwd_mma_inst<1>(xin, dy, S, IC, outer, W1, b1, W2, gW1, gb1, gW2, gb2, num_sms, st);
    else if (S <= 32) *err = launch_propagator_bwd_mma_inst<2>(xin, dy, S, IC, outer, W1, b1, W2, gW1, gb1, gW2, gb2, num_sms, st);
    else *err = launch_propagator_bwd_mma_inst<4>(xin, dy, S, IC, outer, W1, b1, W2, gW1, gb1, gW2, gb2, num_sms, st);
    return true;
}

}  // namespace tante
