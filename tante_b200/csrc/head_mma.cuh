// Fused Taylor head for the bf16 tensor mode (boundary B of SURVEY.md 8(d)):
//   last transposed conv (C1=64 -> D, k0 x k0) of all K orders on mma.sync tensor cores
//   + Horner evaluation of  u(t + i*fi) = u0 + sum_k d_k (i*fi)^k / k!   for i = 1..n_b (per-sample n)
//   + residual u0 + multi-frame emit into (a) the caller's frames (B,n_cap,D,H,W), or (b) the rollout's
//     channels-last history AND the channels-first ring window -- one pass, replacing dec_conv_3
//     (enc_dec_cnn.py:273), the Python Taylor loop + cat (tante.py:165-171), the formatter transpose
//     (datamodule.py:191-192) and the sliding-window cat (r_evaler.py:98).
// Per stage-1 row (one k0 x k0 pixel block): reads K*64 bf16 + 4D fp32 (u0), writes n*4D fp32 (+ ring copy).
// With the 2*64*4D*K FLOP/row on FFMA the kernel would be compute-bound (~8-12 FLOP/B); on mma.sync it is
// bandwidth-bound, which is the point of this variant.
#pragma once
#include "attention_mma.cuh"
#include "kernels_simt.cuh"

namespace tante {

constexpr int kHeadRows = 64;          // rows per CTA: 4 warps x one 16-row m-block
constexpr int kHeadWPitch = 144;       // bytes per weight row (64 bf16 + 16 B pad): conflict-free 32-bit B-fragment loads

template <int KORD, int NB>
__global__ void __launch_bounds__(128) taylor_head_mma_kernel(HeadParams hp, PatchGeom g, long long rows_total, int B) {
    extern __shared__ __align__(1024) uint8_t hsm[];
    const int NO = g.k0 * g.k0 * g.D;
    uint8_t* sZ = hsm;                                              // [KORD][64 rows][128 B] swizzled
    uint8_t* sW = sZ + KORD * kHeadRows * 128;                      // [KORD][NB*8][144 B]
    float* sb = reinterpret_cast<float*>(sW + KORD * NB * 8 * kHeadWPitch);   // [KORD][D]
    const int tid = threadIdx.x, warp = tid / 32, lane = tid % 32;
    const long long row0 = (long long)blockIdx.x * kHeadRows;

    // ---- stage: z tiles by cp.async (coalesced 16-B chunks), weights converted to bf16 [o][c] ----
    {
        const uint32_t zb = (uint32_t)__cvta_generic_to_shared(sZ);
#pragma unroll
        for (int k = 0; k < KORD; ++k) {
            const __nv_bfloat16* zk = reinterpret_cast<const __nv_bfloat16*>(hp.z[k]);
            for (int i = tid; i < kHeadRows * 8; i += 128) {
                const int r = i / 8, c = i % 8;
                const bool ok = row0 + r < rows_total;
                const __nv_bfloat16* src = zk + (ok ? (size_t)(row0 + r) * 64 + c * 8 : 0);
                const uint32_t dst = zb + (uint32_t)(k * kHeadRows * 128 + r * 128 + ((c ^ (r & 7)) << 4));
                const int nbytes = ok ? 16 : 0;
                asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(nbytes) : "memory");
            }
        }
        asm volatile("cp.async.commit_group;" ::: "memory");
#pragma unroll
        for (int k = 0; k < KORD; ++k) {
            for (int i = tid; i < NB * 8 * 64; i += 128) {
                const int o = i / 64, c = i % 64;
                const float w = o < NO ? hp.w3[k][(size_t)c * NO + o] : 0.f;
                *reinterpret_cast<__nv_bfloat16*>(sW + (k * NB * 8 + o) * kHeadWPitch + c * 2) = __float2bfloat16_rn(w);
            }
            for (int i = tid; i < g.D; i += 128) sb[k * g.D + i] = hp.b3[k][i];
        }
        asm volatile("cp.async.wait_group 0;" ::: "memory");
    }
    __syncthreads();

    const int gq = lane >> 2, t = lane & 3;
    const int lrow = (lane & 7) + 8 * ((lane >> 3) & 1), lchk = lane >> 4;
    const uint32_t zb = (uint32_t)__cvta_generic_to_shared(sZ);

    // ---- last deconv as [16 x 64] x [64 x NB*8] per order ----
    float acc[KORD][NB][4];
#pragma unroll
    for (int k = 0; k < KORD; ++k) {
#pragma unroll
        for (int nb = 0; nb < NB; ++nb) acc[k][nb][0] = acc[k][nb][1] = acc[k][nb][2] = acc[k][nb][3] = 0.f;
#pragma unroll
        for (int ks = 0; ks < 4; ++ks) {
            uint32_t a[4];
            const int r = warp * 16 + lrow, c = ks * 2 + lchk;
            ldsm_x4(zb + (uint32_t)(k * kHeadRows * 128 + r * 128 + ((c ^ (r & 7)) << 4)), a[0], a[1], a[2], a[3]);
#pragma unroll
            for (int nb = 0; nb < NB; ++nb) {
                const uint8_t* wr = sW + (k * NB * 8 + nb * 8 + gq) * kHeadWPitch + (ks * 16 + 2 * t) * 2;
                const uint32_t b0 = *reinterpret_cast<const uint32_t*>(wr);
                const uint32_t b1 = *reinterpret_cast<const uint32_t*>(wr + 16);
                mma_bf16_16816(acc[k][nb], a, b0, b1);
            }
        }
    }

    // ---- Horner + emit; this thread owns rows (gq, gq+8) of its m-block, columns nb*8 + 2t + {0,1} ----
    const size_t HW = (size_t)g.H * g.W;
    float* y_out = hp.ptrs ? hp.ptrs->y_out : nullptr;
#pragma unroll
    for (int hh = 0; hh < 2; ++hh) {
        const long long row = row0 + warp * 16 + gq + 8 * hh;
        if (row >= rows_total) continue;
        long long tkn = row / g.R1;
        const int r1 = (int)(row % g.R1);
        const int wp = (int)(tkn % g.Wp); tkn /= g.Wp;
        const int hpp = (int)(tkn % g.Hp);
        const int b = (int)(tkn / g.Hp);
        int h1, w1;
        stage1_row_to_hw(g, hpp, wp, r1, h1, w1);
        const int n = hp.n_arr[b];
        if (n <= 0 && !hp.deriv_dbg) continue;
        const int fc = hp.fcount ? hp.fcount[b] : g.T;
        const int u_slot = (fc + g.T - 1) % g.T;
        const float* u0p = hp.u_ring + ((size_t)(b * g.T + u_slot) * g.D) * HW;
        const int cum = hp.cum ? hp.cum[b] : 0;
#pragma unroll
        for (int nb = 0; nb < NB; ++nb) {
#pragma unroll
            for (int j = 0; j < 2; ++j) {
                const int o = nb * 8 + 2 * t + j;
                if (o >= NO) continue;
                const int d = o % g.D;
                const int cp = (o / g.D) % g.k0;
                const int c = o / (g.D * g.k0);
                const size_t pix = (size_t)(h1 * g.k0 + c) * g.W + (w1 * g.k0 + cp);
                float dk[KORD];
#pragma unroll
                for (int k = 0; k < KORD; ++k) dk[k] = acc[k][nb][2 * hh + j] + sb[k * g.D + d];
                if (hp.deriv_dbg) {
#pragma unroll
                    for (int k = 0; k < KORD; ++k) hp.deriv_dbg[(((size_t)k * B + b) * g.D + d) * HW + pix] = dk[k];
                }
                if (n <= 0) continue;
                const float u0 = u0p[(size_t)d * HW + pix];
                for (int i = 1; i <= n; ++i) {
                    const float dt = (float)i * hp.fi;
                    float v = 0.f;
#pragma unroll
                    for (int k = KORD; k >= 1; --k) v = (dk[k - 1] + v) * (dt / (float)k);
                    const float val = v + u0;
                    if (hp.frames) hp.frames[(((size_t)b * hp.n_cap + (i - 1)) * g.D + d) * HW + pix] = val;
                    if (y_out) {
                        const int fidx = cum + i - 1;
                        if (fidx < hp.n_roll) y_out[(((size_t)b * hp.n_roll + fidx) * HW + pix) * g.D + d] = val;
                        if (i > n - g.T) {
                            const int slot = (fc + i - 1) % g.T;
                            hp.ring_out[((size_t)(b * g.T + slot) * g.D + d) * HW + pix] = val;
                        }
                    }
                }
            }
        }
    }
}

template <int KORD, int NB>
static cudaError_t launch_head_mma_inst(const HeadParams& hp, const PatchGeom& g, long long rows, int B, cudaStream_t st) {
    const size_t smem = (size_t)KORD * kHeadRows * 128 + (size_t)KORD * NB * 8 * kHeadWPitch + (size_t)KORD * g.D * 4;
    // (largest instantiation needs 42 KB of dynamic shared memory: below the 48 KB default limit)
    const unsigned blocks = (unsigned)((rows + kHeadRows - 1) / kHeadRows);
    taylor_head_mma_kernel<KORD, NB><<<blocks, 128, smem, st>>>(hp, g, rows, B);
    return cudaGetLastError();
}

// Returns false when (K, D) is outside the instantiated set (caller falls back to the FFMA kernel).
static bool launch_head_mma(const HeadParams& hp, const PatchGeom& g, int C1, long long rows, int B, cudaStream_t st,
                            cudaError_t* err) {
    if (C1 != 64) return false;
    const int NO = g.k0 * g.k0 * g.D;
    const int nb = (NO + 7) / 8;
    const int K = hp.K;
#define TANTE_HEAD(KO, NBv) \
    if (K == KO && nb <= NBv) { *err = launch_head_mma_inst<KO, NBv>(hp, g, rows, B, st); return true; }
    TANTE_HEAD(1, 2) TANTE_HEAD(2, 2) TANTE_HEAD(3, 2) TANTE_HEAD(4, 2)
    TANTE_HEAD(1, 4) TANTE_HEAD(2, 4) TANTE_HEAD(3, 4)
    TANTE_HEAD(1, 6) TANTE_HEAD(2, 6)
    TANTE_HEAD(1, 8)
#undef TANTE_HEAD
    return false;
}

}  // namespace tante
