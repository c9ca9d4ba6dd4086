// FP32 FFMA GEMM for the exact (<=1e-5) parity mode:  C[M,N] = epi(A[M,K] * W[N,K]^T).
// Both operands K-major ("TN"), the natural nn.Linear layout.  128 x BN x 16 CTA tile,
// 256 threads, 8 x (BN/16) register tile, double-buffered shared memory with register prefetch.
// Requirements: K % 16 == 0, N % BN == 0, lda/ldw % 4 == 0; M arbitrary (row-guarded).
#pragma once
#include "common.cuh"

namespace tante {

template <int BN, int EPI>
__global__ void __launch_bounds__(256) gemm_simt_kernel(const float* __restrict__ A, int lda,
                                                        const float* __restrict__ W, int ldw,
                                                        float* __restrict__ C, int ldc, int M, int N, int K,
                                                        EpiParams ep) {
    constexpr int BM = 128, BK = 16, TN = BN / 16;
    __shared__ __align__(16) float As[2][BK][BM + 4];
    __shared__ __align__(16) float Ws[2][BK][BN + 4];

    if (ep.m_dev) M = min(M, *ep.m_dev * ep.m_rows);      // device-side row count (rollout encoder cache)
    const int tid = threadIdx.x;
    const int m0 = blockIdx.y * BM;
    const int n0 = blockIdx.x * BN;
    if (m0 >= M) return;
    const int ty = tid / 16, tx = tid % 16;

    // global->smem mapping: each thread moves float4s along K.
    // A: 128 rows x 4 float4 = 512 -> 2 per thread.  W: BN rows x 4 float4 -> BN/64 per thread.
    const int lrow = tid / 4, lk = (tid % 4) * 4;
    float4 ra[2], rw[BN / 64];

    auto gload = [&](int k0) {
#pragma unroll
        for (int i = 0; i < 2; ++i) {
            const int r = m0 + lrow + i * 64;
            ra[i] = (r < M) ? *reinterpret_cast<const float4*>(A + (size_t)r * lda + k0 + lk) : make_float4(0, 0, 0, 0);
        }
#pragma unroll
        for (int i = 0; i < BN / 64; ++i) {
            const int r = n0 + lrow + i * 64;
            rw[i] = *reinterpret_cast<const float4*>(W + (size_t)r * ldw + k0 + lk);
        }
    };
    auto sstore = [&](int buf) {
#pragma unroll
        for (int i = 0; i < 2; ++i) {
            const int r = lrow + i * 64;
            As[buf][lk + 0][r] = ra[i].x; As[buf][lk + 1][r] = ra[i].y;
            As[buf][lk + 2][r] = ra[i].z; As[buf][lk + 3][r] = ra[i].w;
        }
#pragma unroll
        for (int i = 0; i < BN / 64; ++i) {
            const int r = lrow + i * 64;
            Ws[buf][lk + 0][r] = rw[i].x; Ws[buf][lk + 1][r] = rw[i].y;
            Ws[buf][lk + 2][r] = rw[i].z; Ws[buf][lk + 3][r] = rw[i].w;
        }
    };

    float acc[8][TN];
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < TN; ++j) acc[i][j] = 0.0f;

    gload(0);
    sstore(0);
    __syncthreads();
    const int nk = K / BK;
    for (int kt = 0; kt < nk; ++kt) {
        const int buf = kt & 1;
        if (kt + 1 < nk) gload((kt + 1) * BK);
#pragma unroll
        for (int k = 0; k < BK; ++k) {
            float a[8], w[TN];
            const float4 a0 = *reinterpret_cast<const float4*>(&As[buf][k][ty * 8]);
            const float4 a1 = *reinterpret_cast<const float4*>(&As[buf][k][ty * 8 + 4]);
            a[0] = a0.x; a[1] = a0.y; a[2] = a0.z; a[3] = a0.w; a[4] = a1.x; a[5] = a1.y; a[6] = a1.z; a[7] = a1.w;
#pragma unroll
            for (int j4 = 0; j4 < TN / 4; ++j4) {
                const float4 w4 = *reinterpret_cast<const float4*>(&Ws[buf][k][tx * TN + j4 * 4]);
                w[j4 * 4 + 0] = w4.x; w[j4 * 4 + 1] = w4.y; w[j4 * 4 + 2] = w4.z; w[j4 * 4 + 3] = w4.w;
            }
#pragma unroll
            for (int i = 0; i < 8; ++i)
#pragma unroll
                for (int j = 0; j < TN; ++j) acc[i][j] = fmaf(a[i], w[j], acc[i][j]);
        }
        if (kt + 1 < nk) {
            sstore(buf ^ 1);
            __syncthreads();
        }
    }

#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const int m = m0 + ty * 8 + i;
        if (m >= M) continue;
#pragma unroll
        for (int j4 = 0; j4 < TN / 4; ++j4) {
            const int n = n0 + tx * TN + j4 * 4;
            float4 o;
            o.x = apply_epilogue<EPI>(acc[i][j4 * 4 + 0], m, n + 0, ep);
            o.y = apply_epilogue<EPI>(acc[i][j4 * 4 + 1], m, n + 1, ep);
            o.z = apply_epilogue<EPI>(acc[i][j4 * 4 + 2], m, n + 2, ep);
            o.w = apply_epilogue<EPI>(acc[i][j4 * 4 + 3], m, n + 3, ep);
            *reinterpret_cast<float4*>(C + (size_t)m * ldc + n) = o;
        }
    }
}

template <int BN>
static cudaError_t launch_gemm_simt_bn(int epi, const float* A, int lda, const float* W, int ldw, float* C, int ldc,
                                       int M, int N, int K, const EpiParams& ep, cudaStream_t st) {
    dim3 grid(N / BN, (M + 127) / 128), block(256);
    switch (epi) {
#define TANTE_CASE(E) \
    case E: gemm_simt_kernel<BN, E><<<grid, block, 0, st>>>(A, lda, W, ldw, C, ldc, M, N, K, ep); break;
        TANTE_CASE(EPI_BIAS)
        TANTE_CASE(EPI_BIAS_RELU)
        TANTE_CASE(EPI_BIAS_GELU_ERF)
        TANTE_CASE(EPI_BIAS_GELU_TANH)
        TANTE_CASE(EPI_BIAS_RESID)
        TANTE_CASE(EPI_EMBED)
#undef TANTE_CASE
        default: return cudaErrorInvalidValue;
    }
    return cudaGetLastError();
}

static cudaError_t launch_gemm_simt(int epi, const float* A, int lda, const float* W, int ldw, float* C, int ldc, int M,
                                    int N, int K, const EpiParams& ep, cudaStream_t st) {
    if (M <= 0) return cudaSuccess;
    if (K % 16 != 0 || N % 64 != 0 || lda % 4 != 0 || ldw % 4 != 0 || ldc % 4 != 0) return cudaErrorInvalidValue;
    if (N % 128 == 0) return launch_gemm_simt_bn<128>(epi, A, lda, W, ldw, C, ldc, M, N, K, ep, st);
    return launch_gemm_simt_bn<64>(epi, A, lda, W, ldw, C, ldc, M, N, K, ep, st);
}

}  // namespace tante
