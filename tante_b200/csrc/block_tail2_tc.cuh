// Fused block tail, TWO TILES IN FLIGHT per SM (second generation of block_tail_tc.cuh; same math, same arguments):
//
//     x_mid = x + att * Wo^T + bo ;  h = gelu_tanh(LN2(x_mid) * W1^T + b1) ;  x_out = x_mid + h * W2^T + b2 ;  ln = LN1'(x_out)
//
// The first kernel runs the three GEMM phases and the three epilogues of a tile strictly one after the other (one TMEM
// accumulator + x_mid in the other 256 TMEM columns: nothing left for a second tile), so the tensor pipe idles during the
// epilogues and the epilogue warps idle during the MMAs: 20 % / 32 % busy, 50 % of the HBM roofline.  Here a CTA owns two
// independent tile STREAMS (tiles k = 0, 2, 4 .. and k = 1, 3, 5 .. of its persistent sequence), each with its own 64 KB A
// tile, its own 256-column TMEM accumulator and its own 8 epilogue warps; ONE MMA warp and ONE weight producer serve both in
// the fixed order s0.phase0, s1.phase0, s0.phase1, s1.phase1, s0.phase2, s1.phase2, so the MMAs of one stream run under the
// epilogue of the other.  What pays for the second accumulator is x_mid: it no longer lives in TMEM but takes a round trip
// through L2 (stored by the first epilogue -- in training that store exists anyway, it is the tape -- and TMA-loaded back by
// the third; inference parks it in the x_out rows the tile is about to overwrite).
//
// Shared memory (226 KB): two A tiles (128 KB), a ring of three 16 KB weight stages shared by both streams (48 KB), sixteen
// 2 KB staging buffers (one per epilogue warp), parameters, row statistics, barriers.  A warp owns 32 token rows x 128
// columns.  fp32 rows move in half-chunks of 32 rows x 16 columns (2 KB, SWIZZLE_64B boxes): one lands in the warp's staging
// buffer ahead of the accumulator, the others in the warp's OWN 8 KB of the A tile (the two 32-row x 128-B regions it will
// later fill with LayerNorm / hidden output), which is idle between a GEMM's completion and the rewrite -- five buffers
// cycle through the eight half-chunks of a phase; the updated rows go back to the accumulator columns (for the LayerNorm
// pass) and out by TMA from the same buffers.  No buffer is ever shared between warps, so the only cross-warp traffic of
// an epilogue is the row-statistics exchange of the two warps that share a token row.
// Warp roles: 0 = weight producer, 1 = MMA issuer + TMEM owner, 2..9 = epilogue of stream 0, 10..17 = epilogue of stream 1.
#pragma once
#include "block_tail_tc.cuh"

namespace tante {

constexpr int kB2EpiWarps = 16;
constexpr int kB2Threads = 64 + 32 * kB2EpiWarps;
constexpr int kB2WStages = 3;
constexpr int kB2Half = 32 * 64;            // half-chunk buffer: 32 rows x 16 fp32 = 2 KB (SWIZZLE_64B)
constexpr int kB2NBars = 6 + 10 + 5 * kB2EpiWarps;
constexpr size_t kB2Smem = 1024 + 2 * 4 * kBtKBlk + kB2WStages * kBtWStage + kB2EpiWarps * kB2Half + 7 * kBtC * 4 +
                           2 * 4 * 2 * 32 * 2 * 4 + kB2NBars * 8 + 64;

// byte offset of 16-byte chunk c (0..3) of row r inside a 64-byte-row SWIZZLE_64B box (address bits [4,6) ^= bits [7,9))
__device__ __forceinline__ uint32_t sw64_off(int r, int c) { return (uint32_t)(r * 64 + ((c ^ ((r >> 1) & 3)) << 4)); }

template <int MODE /* 0: inference, 1: training (stores the tape), 2: training with dropout */>
__global__ void __launch_bounds__(kB2Threads, 1)
block_tail2_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmWo,
                   const __grid_constant__ CUtensorMap tmW1, const __grid_constant__ CUtensorMap tmW2,
                   const __grid_constant__ CUtensorMap tmXin, const __grid_constant__ CUtensorMap tmXmid,
                   const __grid_constant__ CUtensorMap tmXout, const __grid_constant__ CUtensorMap tmLn2,
                   const __grid_constant__ CUtensorMap tmHpre, const __grid_constant__ CUtensorMap tmHact,
                   const __grid_constant__ CUtensorMap tmLnOut, BtParams p, long long* trace) {
    constexpr bool TRAIN = MODE != 0;
    constexpr bool dropping = MODE == 2;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = smem_raw + ((1024u - (ptx::smem_u32(smem_raw) & 1023u)) & 1023u);
    // TANTE_TAIL_TRACE: CTA 0 records (event, clock64) pairs -- role 0 = MMA warp, 1 / 2 = first epilogue warp of stream 0 / 1
    int tr_n = 0;
    auto TR = [&](int role, int ev) {
        if (trace && blockIdx.x == 0 && tr_n < 1024) { trace[(role * 1024 + tr_n) * 2] = ev; trace[(role * 1024 + tr_n) * 2 + 1] = clock64(); ++tr_n; }
    };
    uint8_t* sA = smem;                                        // [2 streams][4 k-blocks][128 rows][128 B]
    uint8_t* sW = sA + 2 * 4 * kBtKBlk;                        // [kB2WStages][128 rows][128 B]
    uint8_t* sE = sW + kB2WStages * kBtWStage;                 // [16 warps][2 KB]
    float* sP = reinterpret_cast<float*>(sE + kB2EpiWarps * kB2Half);    // bo, g2, be2, b1, b2, gn, ben
    float* sStat = sP + 7 * kBtC;                              // [2 streams][4 quarters][2 slices][32 rows][2]
    uint64_t* bars = reinterpret_cast<uint64_t*>(sStat + 2 * 4 * 2 * 32 * 2);
    uint64_t* w_full = bars;                       // [3]
    uint64_t* w_empty = bars + 3;                  // [3]
    uint64_t* a_full = bars + 6;                   // [2] att tile landed
    uint64_t* a_empty = bars + 8;                  // [2] the stream's epilogue warps are done with the A tile (8 arrivals)
    uint64_t* a_ready = bars + 10;                 // [2] the epilogue warps rewrote the A tile (8 arrivals)
    uint64_t* acc_full = bars + 12;                // [2] one MMA phase finished
    uint64_t* acc_free = bars + 14;                // [2] the last epilogue is done with the accumulator (8 arrivals)
    uint64_t* cbar = bars + 16;                    // [16 warps][5 buffers] half-chunk landed
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(cbar + 5 * kB2EpiWarps);

    pdl_trigger();
    const int warp = threadIdx.x / 32, lane = threadIdx.x % 32;
    const int tiles = (p.M + 127) / 128;
    const int nk = (int)blockIdx.x < tiles ? (tiles - 1 - (int)blockIdx.x) / (int)gridDim.x + 1 : 0;    // tiles of this CTA
    const int n0 = (nk + 1) / 2, n1 = nk / 2;                                                           // per stream

    for (int i = threadIdx.x; i < kBtC; i += kB2Threads) {
        sP[i] = p.bo[i]; sP[kBtC + i] = p.g2[i]; sP[2 * kBtC + i] = p.be2[i]; sP[3 * kBtC + i] = p.b1[i];
        sP[4 * kBtC + i] = p.b2[i];
        sP[5 * kBtC + i] = p.has_ln_out ? p.gn[i] : 1.f; sP[6 * kBtC + i] = p.has_ln_out ? p.ben[i] : 0.f;
    }
    if (warp == 0 && lane == 0) {
        ptx::prefetch_tmap(&tmA); ptx::prefetch_tmap(&tmWo); ptx::prefetch_tmap(&tmW1); ptx::prefetch_tmap(&tmW2);
        ptx::prefetch_tmap(&tmXin); ptx::prefetch_tmap(&tmXout);
        if (p.has_ln_out) ptx::prefetch_tmap(&tmLnOut);
        for (int s = 0; s < kB2WStages; ++s) { ptx::mbar_init(&w_full[s], 1); ptx::mbar_init(&w_empty[s], 1); }
        for (int s = 0; s < 2; ++s) {
            ptx::mbar_init(&a_full[s], 1);
            ptx::mbar_init(&a_empty[s], 8);
            ptx::mbar_init(&a_ready[s], 8);
            ptx::mbar_init(&acc_full[s], 1);
            ptx::mbar_init(&acc_free[s], 8);
        }
        for (int i = 0; i < 5 * kB2EpiWarps; ++i) ptx::mbar_init(&cbar[i], 1);
        ptx::fence_barrier_init();
    }
    if (warp == 1) { ptx::tmem_alloc(tmem_slot, 512); ptx::tmem_relinquish(); }
    ptx::tc_fence_before();
    __syncthreads();
    ptx::tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    pdl_wait();

    if (warp == 0) {
        // ===== weight producer: 8 stages per (stream, phase), in the order the MMA warp consumes them =====
        if (lane == 0) {
            int ws = 0;
            uint32_t wph = 0;
            for (int it = 0; it < n0; ++it) {
                for (int ph = 0; ph < 3; ++ph) {
                    const CUtensorMap* wm = ph == 0 ? &tmWo : (ph == 1 ? &tmW1 : &tmW2);
                    for (int s = 0; s < 2; ++s) {
                        if (s == 1 && it >= n1) break;
                        for (int kb = 0; kb < 4; ++kb) {
                            for (int nh = 0; nh < 2; ++nh) {
                                ptx::mbar_wait(&w_empty[ws], wph ^ 1u);
                                ptx::mbar_arrive_expect_tx(&w_full[ws], kBtWStage);
                                ptx::tma_load_2d(sW + ws * kBtWStage, wm, &w_full[ws], kb * 64, nh * 128);
                                if (++ws == kB2WStages) { ws = 0; wph ^= 1u; }
                            }
                        }
                    }
                }
            }
        }
    } else if (warp == 1) {
        // ===== MMA issuer: the phases of the two streams interleaved =====
        constexpr uint32_t idesc = ptx::umma_idesc_bf16(128, 128);
        int ws = 0;
        uint32_t wph = 0;
        uint32_t ar[2] = {0u, 0u};
        for (int it = 0; it < n0; ++it) {
            for (int ph = 0; ph < 3; ++ph) {
#pragma unroll
                for (int s = 0; s < 2; ++s) {
                    if (s == 1 && it >= n1) continue;
                    if (ph == 0) {
                        ptx::mbar_wait(&acc_free[s], (uint32_t)(it & 1) ^ 1u);    // the stream's previous tile is done with the accumulator
                        ptx::mbar_wait(&a_full[s], (uint32_t)(it & 1));
                    } else {
                        ptx::mbar_wait(&a_ready[s], ar[s] & 1u);                   // the A tile was rewritten by the stream's epilogue
                        ++ar[s];
                    }
                    ptx::tc_fence_after();
                    if (lane == 0) TR(0, 100 + s * 10 + ph);
                    const uint8_t* a_tile = sA + s * 4 * kBtKBlk;
                    const uint32_t tm = tmem_base + (uint32_t)(s * 256);
                    for (int kb = 0; kb < 4; ++kb) {
                        for (int nh = 0; nh < 2; ++nh) {
                            ptx::mbar_wait(&w_full[ws], wph);
                            ptx::tc_fence_after();
                            if (lane == 0) {
                                const uint64_t da = ptx::umma_desc_k_sw128(ptx::smem_u32(a_tile + kb * kBtKBlk));
                                const uint64_t db = ptx::umma_desc_k_sw128(ptx::smem_u32(sW + ws * kBtWStage));
#pragma unroll
                                for (int k = 0; k < 4; ++k)
                                    ptx::umma_bf16(tm + (uint32_t)(nh * 128), da + (uint64_t)(2 * k), db + (uint64_t)(2 * k), idesc,
                                                   (kb | k) != 0);
                                ptx::umma_commit(&w_empty[ws]);
                            }
                            __syncwarp();
                            if (++ws == kB2WStages) { ws = 0; wph ^= 1u; }
                        }
                    }
                    if (lane == 0) { ptx::umma_commit(&acc_full[s]); TR(0, 200 + s * 10 + ph); }
                    __syncwarp();
                }
            }
        }
    } else {
        // ===== epilogue: stream s, TMEM lane quarter q (thread = one token row), column slice cs (128 columns = 2 k-blocks) =====
        const int ewg = warp - 2;
        const int s = ewg >> 3;
        const int ew = ewg & 7;
        const int q = warp & 3;
        const int cs = ew >> 2;
        const int ns = s ? n1 : n0;
        const int col0 = cs * 128;
        uint8_t* a_tile = sA + s * 4 * kBtKBlk;
        uint8_t* ebuf = sE + (size_t)ewg * kB2Half;
        uint8_t* reg0 = a_tile + (size_t)(2 * cs) * kBtKBlk + (size_t)q * 32 * 128;     // own 32 rows of k-block 2cs
        uint8_t* reg1 = reg0 + kBtKBlk;                                                  // ... and of k-block 2cs + 1
        uint64_t* cb = cbar + ewg * 5;
        float* st = sStat + (size_t)s * (4 * 2 * 32 * 2);
        const uint32_t tm_acc = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(s * 256 + col0);
        const float* bo = sP; const float* g2 = sP + kBtC; const float* be2 = sP + 2 * kBtC; const float* b1 = sP + 3 * kBtC;
        const float* b2 = sP + 4 * kBtC; const float* gn = sP + 5 * kBtC; const float* ben = sP + 6 * kBtC;
        uint32_t af = 0;              // completed acc_full waits
        int row0 = 0;

        // buffer b of the five a phase cycles through: 0 = staging, 1..4 = the halves of the two own A-tile regions
        auto bufp = [&](int b) -> uint8_t* {
            return b == 0 ? ebuf : ((b <= 2 ? reg0 : reg1) + (size_t)((b - 1) & 1) * kB2Half);
        };
        auto load_att = [&](int tile) {
            ptx::mbar_arrive_expect_tx(&a_full[s], 4 * kBtKBlk);
            for (int kb = 0; kb < 4; ++kb) ptx::tma_load_2d(a_tile + kb * kBtKBlk, &tmA, &a_full[s], kb * 64, tile * 128);
        };
        // the two warps of a token row exchange their partial sums -> mean, 1 / std of the 256-column row
        auto row_stats = [&](float rsum, float rsq, float& mean, float& rstd) {
            float* mine = st + ((q * 2 + cs) * 32 + lane) * 2;
            mine[0] = rsum; mine[1] = rsq;
            named_bar_sync(1 + s * 4 + q, 64);
            const float* oth = st + ((q * 2 + (cs ^ 1)) * 32 + lane) * 2;
            const float sm = rsum + oth[0], sq = rsq + oth[1];
            mean = sm * (1.0f / kBtC);
            const float var = fmaxf(sq * (1.0f / kBtC) - mean * mean, 0.f);
            rstd = rsqrtf(var + 1e-5f);
        };
        // LayerNorm of the warp's two k-blocks (rows re-read from the accumulator columns) -> bf16 -> own A-tile regions
        auto ln_regions = [&](float mean, float rstd, const float* gam, const float* bet) {
            const float mr = mean * rstd;
#pragma unroll 1
            for (int kh = 0; kh < 4; ++kh) {
                uint8_t* dst = (kh & 2) ? reg1 : reg0;
                const int hh = kh & 1;
                uint32_t r0[32];
                ptx::tmem_ld_32x32(tm_acc + (uint32_t)(kh * 32), r0);
                ptx::tc_wait_ld();
                const float* gs = gam + col0 + kh * 32;
                const float* bs = bet + col0 + kh * 32;
#pragma unroll
                for (int c = 0; c < 4; ++c) {
                    uint32_t pk[4];
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        const int col = c * 8 + j * 2;
                        const float a = fmaf(fmaf(__uint_as_float(r0[col]), rstd, -mr), gs[col], bs[col]);
                        const float bb = fmaf(fmaf(__uint_as_float(r0[col + 1]), rstd, -mr), gs[col + 1], bs[col + 1]);
                        __nv_bfloat162 h2 = __floats2bfloat162_rn(a, bb);
                        pk[j] = *reinterpret_cast<uint32_t*>(&h2);
                    }
                    *reinterpret_cast<uint4*>(dst + sw128_off(lane, hh * 4 + c)) = make_uint4(pk[0], pk[1], pk[2], pk[3]);
                }
            }
        };
        // One residual pass over the warp's 8 half-chunks: rows = src + drop(acc + bias), written back to the accumulator
        // columns (keep) and TMA-stored to dst.  Half-chunk 0 was requested into the staging buffer by the caller.
        auto resid_pass = [&](const CUtensorMap* src, const CUtensorMap* dst, const float* bias, uint32_t ph3, bool keep,
                              uint32_t site, float& rsum, float& rsq) {
            if (lane == 0) {
                for (int h = 1; h <= 4; ++h) {
                    ptx::mbar_arrive_expect_tx(&cb[h], kB2Half);
                    ptx::tma_load_2d(bufp(h), src, &cb[h], col0 + h * 16, row0);
                }
            }
            rsum = 0.f; rsq = 0.f;
#pragma unroll 1
            for (int h = 0; h < 8; ++h) {
                const int b = h < 5 ? h : h - 5;
                uint8_t* buf = bufp(b);
                const uint32_t par = h >= 5 ? 1u : (h >= 3 ? ph3 : 0u);      // uses of a buffer per tile are even: fixed parities
                uint32_t r[16];
                ptx::tmem_ld_32x16(tm_acc + (uint32_t)(h * 16), r);
                ptx::mbar_wait(&cb[b], par);
                ptx::tc_wait_ld();
                const float* bs = bias + col0 + h * 16;
                uint4 dw = make_uint4(0u, 0u, 0u, 0u);
#pragma unroll
                for (int c = 0; c < 4; ++c) {
                    float4* pp = reinterpret_cast<float4*>(buf + sw64_off(lane, c));
                    const float4 x = *pp;
                    float4 o;
                    o.x = __uint_as_float(r[c * 4 + 0]) + bs[c * 4 + 0];
                    o.y = __uint_as_float(r[c * 4 + 1]) + bs[c * 4 + 1];
                    o.z = __uint_as_float(r[c * 4 + 2]) + bs[c * 4 + 2];
                    o.w = __uint_as_float(r[c * 4 + 3]) + bs[c * 4 + 3];
                    if (dropping) {       // 8 consecutive columns share one Philox group
                        if ((c & 1) == 0)
                            dw = drop_words(p.drop, site, (unsigned long long)(row0 + lane) * (kBtC / 8) + (col0 + h * 16) / 8 + (c >> 1));
                        const int l0 = (c & 1) * 4;
                        o.x *= drop_mul(p.drop, dw, l0); o.y *= drop_mul(p.drop, dw, l0 + 1);
                        o.z *= drop_mul(p.drop, dw, l0 + 2); o.w *= drop_mul(p.drop, dw, l0 + 3);
                    }
                    o.x += x.x; o.y += x.y; o.z += x.z; o.w += x.w;
                    *pp = o;
                    r[c * 4 + 0] = __float_as_uint(o.x); r[c * 4 + 1] = __float_as_uint(o.y);
                    r[c * 4 + 2] = __float_as_uint(o.z); r[c * 4 + 3] = __float_as_uint(o.w);
                    rsum += (o.x + o.y) + (o.z + o.w);
                    rsq += fmaf(o.x, o.x, fmaf(o.y, o.y, fmaf(o.z, o.z, o.w * o.w)));
                }
                if (keep) ptx::tmem_st_32x16(tm_acc + (uint32_t)(h * 16), r);
                ptx::fence_proxy_async();
                __syncwarp();
                if (lane == 0) {
                    ptx::tma_store_2d(dst, buf, col0 + h * 16, row0);
                    ptx::bulk_commit();
                    if (h >= 1 && h <= 3) {       // the store of half-chunk h - 1 has read its buffer: refill it with h + 4
                        ptx::bulk_wait_read<1>();
                        ptx::mbar_arrive_expect_tx(&cb[h - 1], kB2Half);
                        ptx::tma_load_2d(bufp(h - 1), src, &cb[h - 1], col0 + (h + 4) * 16, row0);
                    }
                }
                __syncwarp();
            }
            if (lane == 0) ptx::bulk_wait_read<0>();      // every buffer (staging + own A-tile regions) is free again
            __syncwarp();
        };

        if (ew == 0 && lane == 0 && ns > 0) load_att((int)blockIdx.x + s * (int)gridDim.x);
        for (int it = 0; it < ns; ++it) {
            const int tile = (int)blockIdx.x + (2 * it + s) * (int)gridDim.x;
            row0 = tile * 128 + q * 32;

            // ---------------- phase 1: x_mid = x + drop(acc + bo) ; LN2 -> A tile ----------------
            if (lane == 0) {
                ptx::mbar_arrive_expect_tx(&cb[0], kB2Half);
                ptx::tma_load_2d(ebuf, &tmXin, &cb[0], col0, row0);
            }
            if (ew == 0 && lane == 0) TR(1 + s, 10);
            ptx::mbar_wait(&acc_full[s], af & 1u); ++af;
            ptx::tc_fence_after();
            if (ew == 0 && lane == 0) TR(1 + s, 11);
            float rsum, rsq, mean, rstd;
            resid_pass(&tmXin, TRAIN ? &tmXmid : &tmXout, bo, 0u, true, p.site1, rsum, rsq);
            ptx::tc_wait_st();
            if (ew == 0 && lane == 0) TR(1 + s, 12);
            row_stats(rsum, rsq, mean, rstd);
            if (ew == 0 && lane == 0) TR(1 + s, 13);
            ln_regions(mean, rstd, g2, be2);
            if (ew == 0 && lane == 0) TR(1 + s, 14);
            ptx::fence_proxy_async();
            if (TRAIN) {      // the saved LN2 output leaves straight from the A tile
                __syncwarp();
                if (lane == 0) {
                    ptx::tma_store_2d(&tmLn2, reg0, col0, row0);
                    ptx::tma_store_2d(&tmLn2, reg1, col0 + 64, row0);
                    ptx::bulk_commit();
                }
            }
            ptx::tc_fence_before();
            __syncwarp();
            if (lane == 0) ptx::mbar_arrive(&a_ready[s]);

            // ---------------- phase 2: hidden = gelu_tanh(acc + b1) -> A tile ----------------
            ptx::mbar_wait(&acc_full[s], af & 1u); ++af;
            ptx::tc_fence_after();
            if (ew == 0 && lane == 0) TR(1 + s, 21);
            if (TRAIN) {
                // the LN2 stores have read the regions; the bf16 pre-activation goes out through them first
                if (lane == 0) ptx::bulk_wait_read<0>();
                __syncwarp();
#pragma unroll 1
                for (int kh = 0; kh < 4; ++kh) {
                    uint8_t* dst = (kh & 2) ? reg1 : reg0;
                    const int hh = kh & 1;
                    uint32_t r0[32];
                    ptx::tmem_ld_32x32(tm_acc + (uint32_t)(kh * 32), r0);
                    ptx::tc_wait_ld();
                    const float* bs = b1 + col0 + kh * 32;
#pragma unroll
                    for (int c = 0; c < 4; ++c) {
                        uint32_t pp[4];
#pragma unroll
                        for (int j = 0; j < 4; ++j) {
                            const int col = c * 8 + j * 2;
                            __nv_bfloat162 pr = __floats2bfloat162_rn(__uint_as_float(r0[col]) + bs[col], __uint_as_float(r0[col + 1]) + bs[col + 1]);
                            pp[j] = *reinterpret_cast<uint32_t*>(&pr);
                        }
                        *reinterpret_cast<uint4*>(dst + sw128_off(lane, hh * 4 + c)) = make_uint4(pp[0], pp[1], pp[2], pp[3]);
                    }
                }
                ptx::fence_proxy_async();
                __syncwarp();
                if (lane == 0) {
                    ptx::tma_store_2d(&tmHpre, reg0, col0, row0);
                    ptx::tma_store_2d(&tmHpre, reg1, col0 + 64, row0);
                    ptx::bulk_commit();
                    ptx::bulk_wait_read<0>();
                }
                __syncwarp();
            }
#pragma unroll 1
            for (int kh = 0; kh < 4; ++kh) {
                uint8_t* dst = (kh & 2) ? reg1 : reg0;
                const int hh = kh & 1;
                uint32_t r0[32];
                ptx::tmem_ld_32x32(tm_acc + (uint32_t)(kh * 32), r0);
                ptx::tc_wait_ld();
                const float* bs = b1 + col0 + kh * 32;
#pragma unroll
                for (int c = 0; c < 4; ++c) {
                    uint32_t pk[4];
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        const int col = c * 8 + j * 2;
                        float a = __uint_as_float(r0[col]) + bs[col];
                        float bb = __uint_as_float(r0[col + 1]) + bs[col + 1];
                        if (TRAIN) {
                            // the backward differentiates the activation at the SAVED (bf16) pre-activation: use it here too
                            const float2 f = __bfloat1622float2(__floats2bfloat162_rn(a, bb));
                            a = f.x; bb = f.y;
                        }
                        __nv_bfloat162 h2 = __floats2bfloat162_rn(gelu_tanh_fast(a), gelu_tanh_fast(bb));
                        pk[j] = *reinterpret_cast<uint32_t*>(&h2);
                    }
                    *reinterpret_cast<uint4*>(dst + sw128_off(lane, hh * 4 + c)) = make_uint4(pk[0], pk[1], pk[2], pk[3]);
                }
            }
            ptx::fence_proxy_async();
            if (TRAIN) {
                __syncwarp();
                if (lane == 0) {
                    ptx::tma_store_2d(&tmHact, reg0, col0, row0);
                    ptx::tma_store_2d(&tmHact, reg1, col0 + 64, row0);
                    ptx::bulk_commit();
                }
            }
            ptx::tc_fence_before();
            __syncwarp();
            if (lane == 0) ptx::mbar_arrive(&a_ready[s]);

            // ---------------- phase 3: x_out = x_mid + drop(acc + b2) ; LN1' ----------------
            if (ew == 0 && lane == 0) TR(1 + s, 24);
            if (lane == 0) {
                ptx::bulk_wait_all<0>();      // this warp's x_mid rows are in memory (and every earlier store has left its buffer)
                ptx::mbar_arrive_expect_tx(&cb[0], kB2Half);
                ptx::tma_load_2d(ebuf, TRAIN ? &tmXmid : &tmXout, &cb[0], col0, row0);
            }
            ptx::mbar_wait(&acc_full[s], af & 1u); ++af;
            ptx::tc_fence_after();
            if (ew == 0 && lane == 0) TR(1 + s, 31);
            resid_pass(TRAIN ? &tmXmid : &tmXout, &tmXout, b2, 1u, p.has_ln_out != 0, p.site2, rsum, rsq);
            if (ew == 0 && lane == 0) TR(1 + s, 32);
            if (p.has_ln_out) {
                ptx::tc_wait_st();
                row_stats(rsum, rsq, mean, rstd);
                if (ew == 0 && lane == 0) TR(1 + s, 33);
                ln_regions(mean, rstd, gn, ben);
                ptx::fence_proxy_async();
                __syncwarp();
                if (lane == 0) {
                    ptx::tma_store_2d(&tmLnOut, reg0, col0, row0);
                    ptx::tma_store_2d(&tmLnOut, reg1, col0 + 64, row0);
                    ptx::bulk_commit();
                    ptx::bulk_wait_read<0>();
                }
            }
            ptx::tc_fence_before();
            __syncwarp();
            if (ew == 0 && lane == 0) TR(1 + s, 34);
            if (lane == 0) {
                ptx::mbar_arrive(&acc_free[s]);       // the stream's next first GEMM may overwrite the accumulator ...
                ptx::mbar_arrive(&a_empty[s]);        // ... and its att rows may land in the A tile
            }
            if (ew == 0 && it + 1 < ns) {
                if (lane == 0) {
                    ptx::mbar_wait(&a_empty[s], (uint32_t)(it & 1));
                    load_att(tile + 2 * (int)gridDim.x);
                    TR(1 + s, 35);
                }
                __syncwarp();
            }
        }
        if (lane == 0) ptx::bulk_wait_all<0>();
    }
    ptx::tc_fence_before();
    __syncthreads();
    if (warp == 1) ptx::tmem_dealloc(tmem_base, 512);
}

static cudaError_t bt2_set_attrs() {
    static unsigned long long done = 0;
    if (!get_encode_tiled()) return cudaErrorNotSupported;
    if (!attrs_needed(done)) return cudaSuccess;
    cudaError_t e;
    if ((e = cudaFuncSetAttribute(block_tail2_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kB2Smem)) != cudaSuccess) return e;
    if ((e = cudaFuncSetAttribute(block_tail2_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kB2Smem)) != cudaSuccess) return e;
    if ((e = cudaFuncSetAttribute(block_tail2_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kB2Smem)) != cudaSuccess) return e;
    return cudaSuccess;
}

// TANTE_TAIL_STREAMS = 2 selects this kernel; the default (1) is the first-generation kernel: measured on B200 at M = 262144
// the two-stream version is 262 us vs 246 us.  Its clock64 trace (tools/tail_trace.py) shows why: the kernel is bound by the
// L2 -> SM delivery rate (~43 B/cycle/SM, 12.4 TB/s chip-wide), not by the serialisation -- per 128-token tile the weight
// stream alone is 384 KB, the activations 384 KB, and the x_mid round trip adds another 256 KB, so the overlap it buys is
// paid back in L2 bytes (MMA phases stretch from 3.4k to 4-10k cycles waiting for weight stages, residual passes to 11-13k).
static int tail_streams() {
    static const int n = getenv("TANTE_TAIL_STREAMS") ? atoi(getenv("TANTE_TAIL_STREAMS")) : 1;
    return n;
}

static cudaError_t launch_block_tail2(const BlockTailArgs& a, int M, bool train, int num_sms, cudaStream_t st) {
    if (M <= 0) return cudaSuccess;
    { cudaError_t e = bt2_set_attrs(); if (e != cudaSuccess) return e; }
    if (train && (!a.x_mid || !a.ln2 || !a.hpre || !a.hact)) return cudaErrorInvalidValue;
    CUtensorMap tmA, tmWo, tmW1, tmW2, tmXin, tmXmid, tmXout, tmLn2, tmHpre, tmHact, tmLnOut;
    const CUtensorMapDataType BF = CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, F32 = CU_TENSOR_MAP_DATA_TYPE_FLOAT32;
    const CUtensorMapSwizzle S64 = CU_TENSOR_MAP_SWIZZLE_64B;
    bool ok = make_tmap_2d(&tmA, BF, 2, a.att, M, kBtC, kBtC, 64, 128) &&
              make_tmap_2d(&tmWo, BF, 2, a.Wo, kBtC, kBtC, kBtC, 64, 128) &&
              make_tmap_2d(&tmW1, BF, 2, a.W1, kBtC, kBtC, kBtC, 64, 128) &&
              make_tmap_2d(&tmW2, BF, 2, a.W2, kBtC, kBtC, kBtC, 64, 128) &&
              make_tmap_2d(&tmXin, F32, 4, a.x_in, M, kBtC, kBtC, 16, 32, S64) &&
              make_tmap_2d(&tmXout, F32, 4, a.x_out, M, kBtC, kBtC, 16, 32, S64);
    if (!ok) return cudaErrorInvalidValue;
    tmXmid = tmXout; tmLn2 = tmA; tmHpre = tmA; tmHact = tmA; tmLnOut = tmA;
    if (a.ln_out && !make_tmap_2d(&tmLnOut, BF, 2, a.ln_out, M, kBtC, kBtC, 64, 32)) return cudaErrorInvalidValue;
    if (train) {
        ok = make_tmap_2d(&tmXmid, F32, 4, a.x_mid, M, kBtC, kBtC, 16, 32, S64) &&
             make_tmap_2d(&tmLn2, BF, 2, a.ln2, M, kBtC, kBtC, 64, 32) &&
             make_tmap_2d(&tmHpre, BF, 2, a.hpre, M, kBtC, kBtC, 64, 32) &&
             make_tmap_2d(&tmHact, BF, 2, a.hact, M, kBtC, kBtC, 64, 32);
        if (!ok) return cudaErrorInvalidValue;
    }
    BtParams p;
    p.bo = a.bo; p.g2 = a.g2; p.be2 = a.be2; p.b1 = a.b1; p.b2 = a.b2; p.gn = a.gn; p.ben = a.ben;
    p.M = M; p.has_ln_out = a.ln_out != nullptr;
    p.drop = train ? a.drop : DropCfg(); p.site1 = a.site1; p.site2 = a.site2;
    const int tiles = (M + 127) / 128;
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)std::min(tiles, num_sms)); cfg.blockDim = dim3(kB2Threads); cfg.dynamicSmemBytes = kB2Smem; cfg.stream = st;
    cudaLaunchAttribute at[1];
    int na = 0;
    if (pdl_enabled(st)) {
        at[na].id = cudaLaunchAttributeProgrammaticStreamSerialization;
        at[na].val.programmaticStreamSerializationAllowed = 1;
        ++na;
    }
    cfg.attrs = at; cfg.numAttrs = (unsigned)na;
    long long* trace = nullptr;
    static const char* trace_path = getenv("TANTE_TAIL_TRACE");      // debugging aid: per-phase clock64 stamps of CTA 0
    if (trace_path) {
        if (cudaMalloc(&trace, 3 * 1024 * 2 * sizeof(long long)) != cudaSuccess) return cudaErrorMemoryAllocation;
        cudaMemsetAsync(trace, 0, 3 * 1024 * 2 * sizeof(long long), st);
    }
    cudaError_t e;
    if (train && p.drop.p > 0.f) e = cudaLaunchKernelEx(&cfg, block_tail2_kernel<2>, tmA, tmWo, tmW1, tmW2, tmXin, tmXmid, tmXout, tmLn2, tmHpre, tmHact, tmLnOut, p, trace);
    else if (train) e = cudaLaunchKernelEx(&cfg, block_tail2_kernel<1>, tmA, tmWo, tmW1, tmW2, tmXin, tmXmid, tmXout, tmLn2, tmHpre, tmHact, tmLnOut, p, trace);
    else e = cudaLaunchKernelEx(&cfg, block_tail2_kernel<0>, tmA, tmWo, tmW1, tmW2, tmXin, tmXmid, tmXout, tmLn2, tmHpre, tmHact, tmLnOut, p, trace);
    if (trace) {
        std::vector<long long> hbuf(3 * 1024 * 2);
        cudaStreamSynchronize(st);
        cudaMemcpy(hbuf.data(), trace, hbuf.size() * sizeof(long long), cudaMemcpyDeviceToHost);
        cudaFree(trace);
        if (FILE* f = fopen(trace_path, "wb")) { fwrite(hbuf.data(), sizeof(long long), hbuf.size(), f); fclose(f); }
    }
    return e;
}

static cudaError_t launch_block_tail_any(const BlockTailArgs& a, int M, bool train, int num_sms, cudaStream_t st) {
    return tail_streams() >= 2 ? launch_block_tail2(a, M, train, num_sms, st) : launch_block_tail(a, M, train, num_sms, st);
}

}  // namespace tante
