// Weight-gradient GEMM on tcgen05 for the bf16 training mode:
//     dW[N][K] += dY[M][N]^T * X[M][K]          (reduction over the M tokens, fp32 accumulate)
// i.e. the `.grad` of every nn.Linear / patch-conv weight the reference gets from autograd.
//
// Both operands arrive with the REDUCTION dimension as rows (token-major activations, exactly as the forward
// wrote them), so they are fed to the tensor core as MN-major SWIZZLE_128B operands: a TMA box of 64 rows x 64
// columns (128 B per row) is one column block of the canonical layout ((8,n),(8,k)):((1,LBO),(8,SBO)) with
// SBO = 1024 B (next 8 reduction rows) and LBO = 8 KB (next 64-column block) -- no transposes anywhere.
// One CTA owns a 128 x BN tile of dW for a slice of M (split-M), accumulates it in TMEM over a 4-stage
// TMA/mbarrier ring and adds it to the fp32 gradient arena with TMA reduce-add stores (cp.reduce.async.bulk).
// Warp roles: 0 = TMA producer, 1 = MMA issuer + TMEM owner, 2..5 = epilogue.
#pragma once
#include "gemm_tc.cuh"

namespace tante {

constexpr int kWgRows = 64;                    // reduction rows per stage
constexpr int kWgBox = kWgRows * 128;          // one 64 x 64 bf16 box = 8 KB
constexpr int kWgThreads = 192;

namespace ptx {
// MN-major SWIZZLE_128B descriptor: LBO = byte distance between 64-element column blocks, SBO = between 8-row groups
__device__ __forceinline__ uint64_t umma_desc_mn_sw128(uint32_t smem_addr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
    return (uint64_t)((smem_addr >> 4) & 0x3FFF) | ((uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16) |
           ((uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32) | (1ull << 46) | (2ull << 61);
}
__host__ __device__ constexpr uint32_t umma_idesc_bf16_mn(int M, int N) {
    return umma_idesc_bf16(M, N) | (1u << 15) | (1u << 16);     // a_major = b_major = MN
}
// 1-D bulk copy global -> shared (async proxy), completion on an mbarrier; 16-byte aligned, size % 16 == 0
__device__ __forceinline__ void bulk_load_1d(void* smem_dst, const void* gsrc, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(gsrc)), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}
__device__ __forceinline__ void tma_reduce_add_2d(const CUtensorMap* m, const void* smem_src, int c0, int c1) {
    asm volatile("cp.reduce.async.bulk.tensor.2d.global.shared::cta.add.tile.bulk_group [%0, {%2, %3}], [%1];"
                 ::"l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(smem_src)), "r"(c0), "r"(c1)
                 : "memory");
}
}  // namespace ptx

template <int BN>
__global__ void __launch_bounds__(kWgThreads, 1)
wgrad_tc_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                const __grid_constant__ CUtensorMap tmC, int m_tiles, int tiles_per_split, int nstage,
                float* __restrict__ bias_grad, int n_total, const __nv_bfloat16* __restrict__ ones_g) {
    pdl_trigger();
    extern __shared__ uint8_t smem_raw[];
    // 1024-byte alignment (SWIZZLE_128B) as an OFFSET into the shared array: a pointer -> integer -> pointer round trip would
    // lose the address space and turn every shared-memory access of the epilogue into a generic LD / ST
    uint8_t* smem = smem_raw + ((1024u - (ptx::smem_u32(smem_raw) & 1023u)) & 1023u);
    constexpr int kStage = (2 + BN / 64) * kWgBox;
    uint8_t* sStage = smem + kWgBox;                                  // [nstage][A: 2 boxes][B: BN/64 boxes]
    // bias gradient db[n] = sum_m dY[m][n] rides along as one extra N = 64 MMA per k-step against a tile of ones
    // (accumulated in TMEM columns [BN, BN+64)): no separate column-sum pass over dY
    uint8_t* sOnes = smem;                                            // 64 rows x 128 B of bf16 1.0
    uint64_t* bars = reinterpret_cast<uint64_t*>(sStage + (size_t)nstage * kStage);
    constexpr int kTmemCols = BN >= 256 ? 512 : (BN >= 128 ? 256 : 128);
    const bool do_bias = bias_grad != nullptr && blockIdx.y == 0;
    uint64_t* full = bars;
    uint64_t* empty = bars + 8;
    uint64_t* tmem_full = bars + 16;
    uint64_t* ones_full = bars + 17;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 18);

    const int warp = threadIdx.x / 32, lane = threadIdx.x % 32;
    const int n0 = blockIdx.x * 128, k0 = blockIdx.y * BN;
    const int t_begin = blockIdx.z * tiles_per_split;
    const int t_end = min(m_tiles, t_begin + tiles_per_split);
    const int ntile = t_end - t_begin;

    if (warp == 0 && lane == 0) {
        ptx::prefetch_tmap(&tmA);
        ptx::prefetch_tmap(&tmB);
        ptx::prefetch_tmap(&tmC);
        for (int s = 0; s < nstage; ++s) { ptx::mbar_init(&full[s], 1); ptx::mbar_init(&empty[s], 1); }
        ptx::mbar_init(tmem_full, 1);
        ptx::mbar_init(ones_full, 1);
        ptx::fence_barrier_init();
    }
    if (warp == 1) {
        ptx::tmem_alloc(tmem_slot, kTmemCols);
        ptx::tmem_relinquish();
    }
    ptx::tc_fence_before();
    __syncthreads();
    ptx::tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    pdl_wait();     // PDL: barriers / TMEM above needed nothing from the previous kernel; the operands below do

    if (ntile > 0) {
        if (warp == 0) {
            if (lane == 0) {
                if (do_bias) {
                    // the tile of ones arrives like every other operand: through the async proxy (a bulk copy from a
                    // small global buffer), so the tensor core sees it without any generic->async proxy hand-over
                    ptx::mbar_arrive_expect_tx(ones_full, kWgBox);
                    ptx::bulk_load_1d(sOnes, ones_g, kWgBox, ones_full);
                }
                int stage = 0;
                uint32_t phase = 0;
                for (int t = t_begin; t < t_end; ++t) {
                    ptx::mbar_wait(&empty[stage], phase ^ 1);
                    ptx::mbar_arrive_expect_tx(&full[stage], kStage);
                    uint8_t* sa = sStage + (size_t)stage * kStage;
                    uint8_t* sb = sa + 2 * kWgBox;
                    ptx::tma_load_2d(sa, &tmA, &full[stage], n0, t * kWgRows);
                    ptx::tma_load_2d(sa + kWgBox, &tmA, &full[stage], n0 + 64, t * kWgRows);
#pragma unroll
                    for (int j = 0; j < BN / 64; ++j)
                        ptx::tma_load_2d(sb + j * kWgBox, &tmB, &full[stage], k0 + j * 64, t * kWgRows);
                    if (++stage == nstage) { stage = 0; phase ^= 1; }
                }
            }
        } else if (warp == 1) {
            constexpr uint32_t idesc = ptx::umma_idesc_bf16_mn(128, BN);
            constexpr uint32_t idesc_b = ptx::umma_idesc_bf16_mn(128, 64);     // one full 64-column SW128 atom (N < 64 is not a valid MN-major SW128 tile)
            const uint32_t sones = ptx::smem_u32(sOnes);
            int stage = 0;
            uint32_t phase = 0;
            if (do_bias) ptx::mbar_wait(ones_full, 0);
            for (int t = 0; t < ntile; ++t) {
                ptx::mbar_wait(&full[stage], phase);
                ptx::tc_fence_after();
                if (lane == 0) {
                    const uint32_t sa = ptx::smem_u32(sStage + (size_t)stage * kStage);
                    const uint32_t sb = sa + 2 * kWgBox;
#pragma unroll
                    for (int k = 0; k < kWgRows / 16; ++k) {
                        const uint64_t da = ptx::umma_desc_mn_sw128(sa + k * 2048, kWgBox, 1024);
                        const uint64_t db = ptx::umma_desc_mn_sw128(sb + k * 2048, kWgBox, 1024);
                        ptx::umma_bf16(tmem_base, da, db, idesc, (t | k) != 0);
                        if (do_bias) {
                            // the descriptor is rebuilt per MMA from an opaque register copy: a loop-invariant 64-bit
                            // descriptor hoisted out of the elected-lane block gets a predicated R2UR from ptxas 12.9 that
                            // is skipped at run time (stale uniform registers -> the MMA reads a garbage address)
                            uint32_t so;
                            asm volatile("mov.u32 %0, %1;" : "=r"(so) : "r"(sones));
                            ptx::umma_bf16(tmem_base + (uint32_t)BN, da, ptx::umma_desc_mn_sw128(so, kWgBox, 1024), idesc_b, (t | k) != 0);
                        }
                    }
                    ptx::umma_commit(&empty[stage]);
                    if (t == ntile - 1) ptx::umma_commit(tmem_full);
                }
                __syncwarp();
                if (++stage == nstage) { stage = 0; phase ^= 1; }
            }
        } else {
            // epilogue: TMEM lane quarter q (rows q*32..+31 of the 128-row tile), 32 fp32 columns per chunk
            const int q = warp & 3;
            ptx::mbar_wait(tmem_full, 0);
            ptx::tc_fence_after();
            // all MMAs have retired -> the stage ring is free: use it as per-warp staging (4 KB each, 1024-B aligned)
            uint8_t* buf = sStage + (size_t)(warp - 2) * kTcEpiBuf;
            const uint32_t tm = tmem_base + ((uint32_t)(q * 32) << 16);
#pragma unroll 1
            for (int ch = 0; ch < BN / 32; ++ch) {
                uint32_t r0[32];
                ptx::tmem_ld_32x32(tm + (uint32_t)(ch * 32), r0);
                if (lane == 0) ptx::bulk_wait_read<0>();
                __syncwarp();
                ptx::tc_wait_ld();
#pragma unroll
                for (int c = 0; c < 8; ++c)
                    *reinterpret_cast<uint4*>(buf + sw128_off(lane, c)) =
                        make_uint4(r0[c * 4 + 0], r0[c * 4 + 1], r0[c * 4 + 2], r0[c * 4 + 3]);
                ptx::fence_proxy_async();
                __syncwarp();
                if (lane == 0) {
                    ptx::tma_reduce_add_2d(&tmC, buf, k0 + ch * 32, n0 + q * 32);
                    ptx::bulk_commit();
                }
            }
            if (do_bias) {
                uint32_t r0[32];
                ptx::tmem_ld_32x32(tm + (uint32_t)BN, r0);
                ptx::tc_wait_ld();
                const int n = n0 + q * 32 + lane;
                if (n < n_total) atomicAdd(bias_grad + n, __uint_as_float(r0[0]));
            }
            if (lane == 0) ptx::bulk_wait_all<0>();
        }
    }
    ptx::tc_fence_before();
    __syncthreads();
    if (warp == 1) ptx::tmem_dealloc(tmem_base, kTmemCols);
}

static cudaError_t wg_set_attrs() {
    static unsigned long long done = 0;
    if (!attrs_needed(done)) return cudaSuccess;
    cudaError_t e;
    if ((e = cudaFuncSetAttribute(wgrad_tc_kernel<256>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024)) != cudaSuccess) return e;
    if ((e = cudaFuncSetAttribute(wgrad_tc_kernel<128>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024)) != cudaSuccess) return e;
    if ((e = cudaFuncSetAttribute(wgrad_tc_kernel<64>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024)) != cudaSuccess) return e;
    return cudaSuccess;
}

// 8 KB of bf16 1.0 per device (the B operand of the fused bias-gradient MMA)
static const __nv_bfloat16* wg_ones(cudaError_t* err) {
    static void* ptr[64] = {nullptr};
    int dev = 0;
    if ((*err = cudaGetDevice(&dev)) != cudaSuccess) return nullptr;
    if (dev < 0 || dev >= 64) { *err = cudaErrorInvalidDevice; return nullptr; }
    if (!ptr[dev]) {
        uint16_t host[kWgBox / 2];
        for (int i = 0; i < kWgBox / 2; ++i) host[i] = 0x3F80;
        void* p = nullptr;
        if ((*err = cudaMalloc(&p, kWgBox)) != cudaSuccess) return nullptr;
        if ((*err = cudaMemcpy(p, host, kWgBox, cudaMemcpyHostToDevice)) != cudaSuccess) return nullptr;
        ptr[dev] = p;
    }
    return reinterpret_cast<const __nv_bfloat16*>(ptr[dev]);
}

// true when the shape is covered by the tensor-core kernel (otherwise the caller uses wgrad_simt).
// N (columns of A = rows of dW) must be a multiple of 64: an odd 64-column half-tile is zero-filled by TMA on load
// and clipped on the reduce-store.  Kb = columns of B as stored (multiple of 64, zero-padded by the producer),
// Kc <= Kb = columns of dW actually kept (the rest is clipped by the output tensor map).
static bool wgrad_tc_supported(long long M, int N, int Kb, int Kc, int lda, int ldb, int ldc) {
    return M >= 64 && M < (1LL << 31) && N % 64 == 0 && Kb % 64 == 0 && Kc <= Kb && Kc >= 1 && lda % 8 == 0 &&
           ldb % 8 == 0 && ldc % 4 == 0;
}

static cudaError_t launch_wgrad_tc(const __nv_bfloat16* A, int lda, const __nv_bfloat16* Bm, int ldb, float* Cout,
                                   int ldc, long long M, int N, int Kb, int Kc, int num_sms, cudaStream_t st,
                                   float* bias_grad = nullptr) {
    if (!wgrad_tc_supported(M, N, Kb, Kc, lda, ldb, ldc)) return cudaErrorInvalidValue;
    const int BN = (Kb % 256 == 0) ? 256 : (Kb % 128 == 0 ? 128 : 64);
    const int stage_bytes = (2 + BN / 64) * kWgBox;
    int nstage = (int)((227 * 1024 - 2048 - kWgBox) / stage_bytes);
    nstage = nstage > 6 ? 6 : nstage;
    const size_t smem = (size_t)nstage * stage_bytes + kWgBox + 256 + 1024;
    const int m_tiles = (int)((M + kWgRows - 1) / kWgRows);
    const int n_tiles = (N + 127) / 128;
    const int tiles_xy = n_tiles * (Kb / BN);
    int splits = std::max(1, num_sms / tiles_xy);
    splits = std::min(splits, m_tiles);
    const int per = (m_tiles + splits - 1) / splits;
    splits = (m_tiles + per - 1) / per;
    CUtensorMap tmA, tmB, tmC;
    if (!make_tmap_2d(&tmA, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, A, (int)M, N, lda, 64, kWgRows)) return cudaErrorInvalidValue;
    if (!make_tmap_2d(&tmB, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, Bm, (int)M, Kb, ldb, 64, kWgRows)) return cudaErrorInvalidValue;
    if (!make_tmap_2d(&tmC, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, Cout, N, Kc, ldc, 32, 32)) return cudaErrorInvalidValue;
    { cudaError_t e = wg_set_attrs(); if (e != cudaSuccess) return e; }
    const __nv_bfloat16* ones = nullptr;
    if (bias_grad) { cudaError_t e = cudaSuccess; ones = wg_ones(&e); if (e != cudaSuccess) return e; }
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(n_tiles, Kb / BN, splits); cfg.blockDim = dim3(kWgThreads); cfg.dynamicSmemBytes = smem; cfg.stream = st;
    cudaLaunchAttribute at[1];
    int na = 0;
    if (pdl_enabled(st)) {
        at[na].id = cudaLaunchAttributeProgrammaticStreamSerialization;
        at[na].val.programmaticStreamSerializationAllowed = 1;
        ++na;
    }
    cfg.attrs = at; cfg.numAttrs = (unsigned)na;
    switch (BN) {
        case 256: return cudaLaunchKernelEx(&cfg, wgrad_tc_kernel<256>, tmA, tmB, tmC, m_tiles, per, nstage, bias_grad, N, ones);
        case 128: return cudaLaunchKernelEx(&cfg, wgrad_tc_kernel<128>, tmA, tmB, tmC, m_tiles, per, nstage, bias_grad, N, ones);
        default: return cudaLaunchKernelEx(&cfg, wgrad_tc_kernel<64>, tmA, tmB, tmC, m_tiles, per, nstage, bias_grad, N, ones);
    }
}

}  // namespace tante
