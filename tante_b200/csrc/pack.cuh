// Weight packing: one launch turns the caller-owned fp32 master parameters (reference state_dict
// layout) into the library's GEMM-ready arena (fp32, plus a bf16 mirror in tensor mode).
#pragma once
#include "common.cuh"

namespace tante {

enum PackMode : int {
    PACK_COPY = 0,       // same layout
    PACK_CONV = 1,       // Conv2d (Co,Ci,k,k)          -> [Co][(di*k+dj)*Ci + ci]
    PACK_DECONV_NK = 2,  // ConvTranspose2d (Ci,Co,k,k) -> [(di*k+dj)*Co + co][Ci]   (GEMM W[N][K])
    PACK_DECONV_KN = 3,  // ConvTranspose2d (Ci,Co,k,k) -> [Ci][(di*k+dj)*Co + co]   (fused head)
    PACK_BIAS_REP = 4,   // bias (Co) -> [(di*k+dj)*Co + co] replicated k*k times
};

struct PackDesc {
    const float* src;
    long long dst_off;   // in elements, into the arena
    long long numel;     // destination elements
    int mode;
    int d0, d1, k;       // (Co,Ci,k) for CONV, (Ci,Co,k) for DECONV_*, (Co,-,k) for BIAS_REP
};

__global__ void __launch_bounds__(256) pack_params_kernel(const PackDesc* __restrict__ descs, float* __restrict__ arena,
                                                          __nv_bfloat16* __restrict__ arena_bf16) {
    const PackDesc d = descs[blockIdx.y];
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < d.numel;
         i += (long long)gridDim.x * blockDim.x) {
        long long s = i;
        const int kk = d.k * d.k;
        if (d.mode == PACK_CONV) {
            const int Ci = d.d1;
            const int ci = (int)(i % Ci);
            const int dd = (int)((i / Ci) % kk);
            const int co = (int)(i / ((long long)Ci * kk));
            s = ((long long)co * Ci + ci) * kk + dd;
        } else if (d.mode == PACK_DECONV_NK) {
            const int Ci = d.d0, Co = d.d1;
            const int ci = (int)(i % Ci);
            const int n = (int)(i / Ci);
            const int co = n % Co, dd = n / Co;
            s = ((long long)ci * Co + co) * kk + dd;
        } else if (d.mode == PACK_DECONV_KN) {
            const int Co = d.d1;
            const int n = (int)(i % ((long long)kk * Co));
            const int ci = (int)(i / ((long long)kk * Co));
            const int co = n % Co, dd = n / Co;
            s = ((long long)ci * Co + co) * kk + dd;
        } else if (d.mode == PACK_BIAS_REP) {
            s = i % d.d0;
        }
        const float v = d.src[s];
        arena[d.dst_off + i] = v;
        if (arena_bf16) arena_bf16[d.dst_off + i] = __float2bfloat16_rn(v);
    }
}

}  // namespace tante
