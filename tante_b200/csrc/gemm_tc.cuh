// bf16 tensor-core GEMM for sm_100a:  C[M,N] = epi(A[M,K] * W[N,K]^T), fp32 accumulation in TMEM.
//
// Design (weight-stationary, persistent, warp-specialised):
//   * every GEMM of the TANTE block has a short reduction (K = 64..512) and a huge M (tokens), so the
//     K-major weight slice W[n0:n0+BN, :] (<= 128 KB) is TMA-loaded into shared memory ONCE per CTA and
//     stays resident while the CTA walks over its M tiles; only activations stream (16 KB k-blocks
//     through a 4-6 stage mbarrier ring) -> ~256 FLOP per byte of L2/HBM traffic at BN = 256;
//   * one elected thread issues tcgen05.mma (M=128, N=BN, K=16) with both operands in SWIZZLE_128B
//     shared memory; the fp32 accumulator is double-buffered in TMEM (2 x BN columns) so the epilogue of
//     tile i overlaps the MMAs of tile i+1;
//   * 4 epilogue warps read TMEM with tcgen05.ld (32 lanes x 32 columns), transpose through a private,
//     conflict-free shared-memory patch and apply bias / activation / residual / FiLM+embeddings with
//     fully coalesced 128-byte global accesses.
// Warp roles: 0 = TMA producer, 1 = MMA issuer + TMEM owner, 2..5 = epilogue.
#pragma once
#include "common.cuh"
#include "sm100_ptx.cuh"

namespace tante {

constexpr int kTcBlockM = 128;
constexpr int kTcBlockK = 64;                       // 64 bf16 = one 128-byte swizzle row
constexpr int kTcStageBytes = kTcBlockM * 128;      // 16 KB
constexpr int kTcMaxStages = 8;
constexpr int kTcEpiBytes = 4 * 32 * 33 * 4;        // per-warp 32x33 fp32 transpose patches
constexpr int kTcThreads = 192;

__device__ __forceinline__ float epi_apply_rt(int epi, float acc, int m, int n, const EpiParams& p) {
    float v = acc + p.bias[n];
    switch (epi) {
        case EPI_BIAS_RELU: v = fmaxf(v, 0.0f); break;
        case EPI_BIAS_GELU_ERF: v = gelu_erf(v); break;
        case EPI_BIAS_GELU_TANH: v = gelu_tanh(v); break;
        case EPI_BIAS_RESID: v = p.resid[(size_t)m * p.ldr + n] + v; break;
        case EPI_EMBED: {
            const int hw = m % p.L;
            const int t = (m / p.L) % p.T;
            v = v + (v * p.film[(size_t)(t * 2 + 0) * p.ldr + n] + p.film[(size_t)(t * 2 + 1) * p.ldr + n]);
            v = v + p.s_emb[(size_t)hw * p.ldr + n];
            v = v + p.t_emb[(size_t)t * p.ldr + n];
            break;
        }
        default: break;
    }
    return v;
}

template <int BN>
__global__ void __launch_bounds__(kTcThreads, 1)
gemm_tc_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmW, void* __restrict__ C,
               int ldc, int M, int N, int nkb, int nstage, int epi, int out_bf16, EpiParams ep) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    uint8_t* sW = smem;                                   // [nkb][BN rows][128 B]
    uint8_t* sA = sW + (size_t)nkb * BN * 128;            // [nstage][128 rows][128 B]
    float* sEpi = reinterpret_cast<float*>(sA + (size_t)nstage * kTcStageBytes);
    uint64_t* bars = reinterpret_cast<uint64_t*>(reinterpret_cast<uint8_t*>(sEpi) + kTcEpiBytes);
    uint64_t* full = bars;                    // [kTcMaxStages]
    uint64_t* empty = bars + kTcMaxStages;    // [kTcMaxStages]
    uint64_t* w_full = bars + 2 * kTcMaxStages;
    uint64_t* tmem_full = w_full + 1;         // [2]
    uint64_t* tmem_empty = tmem_full + 2;     // [2]
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_empty + 2);

    const int warp = threadIdx.x / 32, lane = threadIdx.x % 32;
    const int n_slices = N / BN;
    const int slice = blockIdx.x % n_slices;
    const int rank = blockIdx.x / n_slices;
    const int per_slice = gridDim.x / n_slices;
    const int m_tiles = (M + kTcBlockM - 1) / kTcBlockM;
    const int n0 = slice * BN;

    if (warp == 0 && lane == 0) {
        ptx::prefetch_tmap(&tmA);
        ptx::prefetch_tmap(&tmW);
        for (int s = 0; s < nstage; ++s) { ptx::mbar_init(&full[s], 1); ptx::mbar_init(&empty[s], 1); }
        ptx::mbar_init(w_full, 1);
        for (int i = 0; i < 2; ++i) { ptx::mbar_init(&tmem_full[i], 1); ptx::mbar_init(&tmem_empty[i], 4); }
        ptx::fence_barrier_init();
    }
    if (warp == 1) {
        ptx::tmem_alloc(tmem_slot, 2 * BN);
        ptx::tmem_relinquish();
    }
    ptx::tc_fence_before();
    __syncthreads();
    ptx::tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        // ===== TMA producer =====
        if (lane == 0) {
            ptx::mbar_arrive_expect_tx(w_full, (uint32_t)nkb * BN * 128);
            for (int kb = 0; kb < nkb; ++kb) ptx::tma_load_2d(sW + (size_t)kb * BN * 128, &tmW, w_full, kb * kTcBlockK, n0);
            int stage = 0;
            uint32_t phase = 0;
            for (int mt = rank; mt < m_tiles; mt += per_slice) {
                for (int kb = 0; kb < nkb; ++kb) {
                    ptx::mbar_wait(&empty[stage], phase ^ 1);
                    ptx::mbar_arrive_expect_tx(&full[stage], kTcStageBytes);
                    ptx::tma_load_2d(sA + (size_t)stage * kTcStageBytes, &tmA, &full[stage], kb * kTcBlockK, mt * kTcBlockM);
                    if (++stage == nstage) { stage = 0; phase ^= 1; }
                }
            }
        }
    } else if (warp == 1) {
        // ===== MMA issuer =====
        constexpr uint32_t idesc = ptx::umma_idesc_bf16(kTcBlockM, BN);
        ptx::mbar_wait(w_full, 0);
        int stage = 0;
        uint32_t phase = 0;
        int t = 0;
        for (int mt = rank; mt < m_tiles; mt += per_slice, ++t) {
            const int acc = t & 1;
            ptx::mbar_wait(&tmem_empty[acc], ((t >> 1) & 1) ^ 1);
            ptx::tc_fence_after();
            const uint32_t tmem_d = tmem_base + (uint32_t)(acc * BN);
            for (int kb = 0; kb < nkb; ++kb) {
                ptx::mbar_wait(&full[stage], phase);
                ptx::tc_fence_after();
                if (lane == 0) {
                    const uint64_t da = ptx::umma_desc_k_sw128(ptx::smem_u32(sA + (size_t)stage * kTcStageBytes));
                    const uint64_t db = ptx::umma_desc_k_sw128(ptx::smem_u32(sW + (size_t)kb * BN * 128));
#pragma unroll
                    for (int k = 0; k < kTcBlockK / 16; ++k)
                        ptx::umma_bf16(tmem_d, da + (uint64_t)(2 * k), db + (uint64_t)(2 * k), idesc, (kb | k) != 0);
                    ptx::umma_commit(&empty[stage]);          // frees the smem stage when the MMAs retire
                    if (kb == nkb - 1) ptx::umma_commit(&tmem_full[acc]);
                }
                __syncwarp();
                if (++stage == nstage) { stage = 0; phase ^= 1; }
            }
        }
    } else {
        // ===== epilogue warps (2..5): TMEM lane quarter = warp % 4 =====
        const int q = warp & 3;
        float* patch = sEpi + q * 32 * 33;
        int t = 0;
        for (int mt = rank; mt < m_tiles; mt += per_slice, ++t) {
            const int acc = t & 1;
            ptx::mbar_wait(&tmem_full[acc], (t >> 1) & 1);
            ptx::tc_fence_after();
            const int row0 = mt * kTcBlockM + q * 32;
#pragma unroll 1
            for (int ch = 0; ch < BN / 32; ++ch) {
                uint32_t r[32];
                ptx::tmem_ld_32x32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(acc * BN + ch * 32), r);
                ptx::tc_wait_ld();
#pragma unroll
                for (int c = 0; c < 32; ++c) patch[lane * 33 + c] = __uint_as_float(r[c]);
                __syncwarp();
                const int n = n0 + ch * 32 + lane;
#pragma unroll 4
                for (int rr = 0; rr < 32; ++rr) {
                    const int m = row0 + rr;
                    if (m < M) {
                        const float v = epi_apply_rt(epi, patch[rr * 33 + lane], m, n, ep);
                        if (out_bf16) reinterpret_cast<__nv_bfloat16*>(C)[(size_t)m * ldc + n] = __float2bfloat16_rn(v);
                        else reinterpret_cast<float*>(C)[(size_t)m * ldc + n] = v;
                    }
                }
                __syncwarp();
            }
            ptx::tc_fence_before();
            if (lane == 0) ptx::mbar_arrive(&tmem_empty[acc]);
        }
    }
    ptx::tc_fence_before();
    __syncthreads();
    if (warp == 1) ptx::tmem_dealloc(tmem_base, 2 * BN);
}

// ---- host side ----------------------------------------------------------------------------------
typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                    const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                    CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static PFN_encodeTiled get_encode_tiled() {
    static PFN_encodeTiled fn = nullptr;
    if (!fn) {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess ||
            q != cudaDriverEntryPointSuccess)
            return nullptr;
        fn = reinterpret_cast<PFN_encodeTiled>(p);
    }
    return fn;
}

// 2-D bf16 K-major tensor map: dims {K, rows}, box {64, box_rows}, 128-byte swizzle, zero OOB fill.
static bool make_tmap_bf16(CUtensorMap* m, const void* base, int rows, int K, int ld_elems, int box_rows) {
    PFN_encodeTiled enc = get_encode_tiled();
    if (!enc) return false;
    cuuint64_t dims[2] = {(cuuint64_t)K, (cuuint64_t)rows};
    cuuint64_t strides[1] = {(cuuint64_t)ld_elems * 2};
    cuuint32_t box[2] = {(cuuint32_t)kTcBlockK, (cuuint32_t)box_rows};
    cuuint32_t estr[2] = {1, 1};
    return enc(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), dims, strides, box, estr,
               CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
               CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

struct TcPlan { int BN, nkb, nstage, grid; size_t smem; };

static bool tc_plan(int M, int N, int K, int num_sms, TcPlan* p) {
    if (K % kTcBlockK != 0 || K > 512 || N % 64 != 0) return false;
    const int nkb = K / kTcBlockK;
    int BN = (N % 256 == 0) ? 256 : (N % 128 == 0 ? 128 : 64);
    while ((size_t)nkb * BN * 128 > 128 * 1024 && BN > 64) BN /= 2;   // keep the resident slice <= 128 KB
    const size_t fixed = (size_t)nkb * BN * 128 + kTcEpiBytes + 256 + 1024;
    const size_t budget = 227 * 1024;
    int nstage = (int)((budget - fixed) / kTcStageBytes);
    nstage = nstage > 6 ? 6 : nstage;
    if (nstage < 2) return false;
    const int n_slices = N / BN;
    const int m_tiles = (M + kTcBlockM - 1) / kTcBlockM;
    int per_slice = num_sms / n_slices;
    if (per_slice < 1) per_slice = 1;
    if (per_slice > m_tiles) per_slice = m_tiles;
    p->BN = BN; p->nkb = nkb; p->nstage = nstage; p->grid = per_slice * n_slices;
    p->smem = fixed + (size_t)nstage * kTcStageBytes;
    return true;
}

static cudaError_t launch_gemm_tc(int epi, const __nv_bfloat16* A, int lda, const __nv_bfloat16* W, int ldw, void* C,
                                  int ldc, int out_bf16, int M, int N, int K, const EpiParams& ep, int num_sms,
                                  cudaStream_t st) {
    if (M <= 0) return cudaSuccess;
    TcPlan p;
    if (!tc_plan(M, N, K, num_sms, &p)) return cudaErrorInvalidValue;
    CUtensorMap tmA, tmW;
    if (!make_tmap_bf16(&tmA, A, M, K, lda, kTcBlockM)) return cudaErrorInvalidValue;
    if (!make_tmap_bf16(&tmW, W, N, K, ldw, p.BN)) return cudaErrorInvalidValue;
    static bool attr_done = false;
    if (!attr_done) {
        cudaError_t e;
        if ((e = cudaFuncSetAttribute(gemm_tc_kernel<256>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024)) != cudaSuccess) return e;
        if ((e = cudaFuncSetAttribute(gemm_tc_kernel<128>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024)) != cudaSuccess) return e;
        if ((e = cudaFuncSetAttribute(gemm_tc_kernel<64>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024)) != cudaSuccess) return e;
        attr_done = true;
    }
    switch (p.BN) {
        case 256: gemm_tc_kernel<256><<<p.grid, kTcThreads, p.smem, st>>>(tmA, tmW, C, ldc, M, N, p.nkb, p.nstage, epi, out_bf16, ep); break;
        case 128: gemm_tc_kernel<128><<<p.grid, kTcThreads, p.smem, st>>>(tmA, tmW, C, ldc, M, N, p.nkb, p.nstage, epi, out_bf16, ep); break;
        default: gemm_tc_kernel<64><<<p.grid, kTcThreads, p.smem, st>>>(tmA, tmW, C, ldc, M, N, p.nkb, p.nstage, epi, out_bf16, ep); break;
    }
    return cudaGetLastError();
}

}  // namespace tante
