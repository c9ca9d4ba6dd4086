// patch_scale 16 / 32 / 64 (reference enc_dec_cnn.py:39-46,75-81,93-95,176-184): Patch_map stages with a 4x4 kernel use
// padding (k-1)//2 = 1, so their windows are SHIFTED by one pixel against the patch grid (window i covers rows 4i-1 .. 4i+2,
// zero-padded at the top / left, the last row / column unused), and the matching transposed convs produce 4h-2 rows that
// a bilinear (align_corners=False) resize stretches back to 4h.  The nested pixel order that turns the P <= 8 stages into
// plain GEMMs on reshaped matrices does not survive a shifted window or a resize that blends neighbouring patches, so
// these scales keep every stage in NATURAL grid order (channels-last [b][h_s][w_s][C_s]) and run
//     encoder stage = gather of the (shifted) k x k windows into a patch matrix  ->  GEMM (+ GELU / embed epilogue)
//     decoder stage = GEMM to the k x k sub-pixel matrix  ->  crop + bilinear resample (+ bias, GELU) back to the grid
//     head          = Horner sum + residual + emit over the decoded derivative FIELDS (boundary C of SURVEY.md 8(d))
// Inference / rollout only; the kernels are plain gather / scatter passes (HBM-bound, one pass each).
#pragma once
#include "common.cuh"
#include "kernels_simt.cuh"

namespace tante {

// conv1 windows from the channels-first (ring) frames: out[row = (bt, i, j)][(di*k + dj)*D + d], zero beyond k*k*D
template <typename TA>
__global__ void __launch_bounds__(256)
wide_im2col_cf_kernel(const float* __restrict__ x, const int* __restrict__ fcount, int T, int D, int H, int W, int k, int shift,
                      int Kpad, TA* __restrict__ out, long long total) {
    const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= total) return;
    const int col = (int)(idx % Kpad);
    long long row = idx / Kpad;
    float v = 0.f;
    if (col < k * k * D) {
        const int Ho = H / k, Wo = W / k;
        const int d = col % D, tap = col / D;
        const int di = tap / k, dj = tap % k;
        const int j = (int)(row % Wo); row /= Wo;
        const int i = (int)(row % Ho);
        const long long bt = row / Ho;
        const long long b = bt / T;
        const int t = (int)(bt % T);
        const int slot = fcount ? (fcount[b] + t) % T : t;
        const int y = i * k + di - shift, xx = j * k + dj - shift;
        if (y >= 0 && y < H && xx >= 0 && xx < W)
            v = x[((size_t)(b * T + slot) * D + d) * H * W + (size_t)y * W + xx];
    }
    out[idx] = from_f32<TA>(v);
}

// conv2 / conv3 windows from a channels-last grid [n][Hs][Ws][Cin]: out[row = (n, i, j)][(di*k + dj)*Cin + ci]; thread = 4 channels
template <typename TA>
__global__ void __launch_bounds__(256)
wide_im2col_cl_kernel(const TA* __restrict__ in, int Hs, int Ws, int Cin, int k, int shift, TA* __restrict__ out, long long total4) {
    const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= total4) return;
    const int C4 = Cin / 4;
    const int c4 = (int)(idx % C4);
    long long r = idx / C4;
    const int tap = (int)(r % (k * k)); r /= k * k;
    const int Ho = Hs / k, Wo = Ws / k;
    const int j = (int)(r % Wo); r /= Wo;
    const int i = (int)(r % Ho);
    const long long n = r / Ho;
    const int di = tap / k, dj = tap % k;
    const int y = i * k + di - shift, xx = j * k + dj - shift;
    float v[4] = {0.f, 0.f, 0.f, 0.f};
    if (y >= 0 && y < Hs && xx >= 0 && xx < Ws) Vec4<TA>::load(in + (((size_t)n * Hs + y) * Ws + xx) * Cin + c4 * 4, v);
    Vec4<TA>::store(out + idx * 4, v);
}

// source index / weight of torch's bilinear upsample with align_corners=False (aten UpSample.h: area_pixel_compute_source_index)
__device__ __forceinline__ void bilinear_src(int dst, int in_size, int out_size, int& i0, int& i1, float& w1) {
    const float scale = (float)in_size / (float)out_size;
    float src = scale * ((float)dst + 0.5f) - 0.5f;
    src = src < 0.f ? 0.f : src;
    i0 = min((int)src, in_size - 1);
    i1 = min(i0 + 1, in_size - 1);
    w1 = src - (float)i0;
}

// Transposed-conv output value at (y, x) of the (k*hi - 2*pad) grid from the sub-pixel matrix S[(n, i, j)][(di*k + dj)*Cout + co]:
// y + pad = k*i + di  (pad = (k-1)//2: 0 for k = 2, 1 for k = 4)
template <typename TA>
__device__ __forceinline__ float deconv_at(const TA* __restrict__ S, int ldS, long long n, int hi, int wi, int Cout, int k, int pad,
                                           int y, int x, int co) {
    const int yy = y + pad, xx = x + pad;
    const int i = yy / k, di = yy - i * k, j = xx / k, dj = xx - j * k;
    return to_f32(S[(((size_t)n * hi + i) * wi + j) * ldS + (di * k + dj) * Cout + co]);
}

// decoder stage tail: sub-pixel matrix (bias already added by the GEMM epilogue when `bias` is null) -> [crop + bilinear resize
// for k = 4] -> (+ bias) -> (GELU) -> channels-last grid out[n][k*hi][k*wi][Cout]  or, for the last stage, the channels-first
// fp32 derivative field field[n][Cout][k*hi][k*wi].
template <typename TA, bool ACT, bool FIELD>
__global__ void __launch_bounds__(256)
wide_deconv_post_kernel(const TA* __restrict__ S, int ldS, int hi, int wi, int Cout, int k, const float* __restrict__ bias,
                        TA* __restrict__ out, float* __restrict__ field, long long total) {
    const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= total) return;
    const int Ho = hi * k, Wo = wi * k;
    int co, X, Y;
    long long n;
    if (FIELD) {      // idx = ((n*Cout + co)*Ho + Y)*Wo + X : coalesced stores into the pixel planes
        long long r = idx;
        X = (int)(r % Wo); r /= Wo;
        Y = (int)(r % Ho); r /= Ho;
        co = (int)(r % Cout); n = r / Cout;
    } else {          // idx = ((n*Ho + Y)*Wo + X)*Cout + co : coalesced stores into the channels-last grid
        long long r = idx;
        co = (int)(r % Cout); r /= Cout;
        X = (int)(r % Wo); r /= Wo;
        Y = (int)(r % Ho); n = r / Ho;
    }
    const int pad = (k - 1) / 2;
    float v;
    if (pad == 0) {
        v = deconv_at(S, ldS, n, hi, wi, Cout, k, 0, Y, X, co);
    } else {
        const int Hd = Ho - 2 * pad, Wd = Wo - 2 * pad;       // what ConvTranspose2d produced (enc_dec_cnn.py:162-166)
        int y0, y1, x0, x1;
        float wy, wx;
        bilinear_src(Y, Hd, Ho, y0, y1, wy);
        bilinear_src(X, Wd, Wo, x0, x1, wx);
        const float v00 = deconv_at(S, ldS, n, hi, wi, Cout, k, pad, y0, x0, co), v01 = deconv_at(S, ldS, n, hi, wi, Cout, k, pad, y0, x1, co);
        const float v10 = deconv_at(S, ldS, n, hi, wi, Cout, k, pad, y1, x0, co), v11 = deconv_at(S, ldS, n, hi, wi, Cout, k, pad, y1, x1, co);
        // aten upsample_bilinear2d: h0lambda * (w0lambda * v00 + w1lambda * v01) + h1lambda * (w0lambda * v10 + w1lambda * v11)
        v = (1.f - wy) * ((1.f - wx) * v00 + wx * v01) + wy * ((1.f - wx) * v10 + wx * v11);
    }
    if (bias) v += bias[co];
    if (ACT) v = ActMath<TA>::gelu_erf_f(v);
    if (FIELD) field[idx] = v;
    else out[idx] = from_f32<TA>(v);
}

// Taylor / Horner emit over decoded derivative fields (tante.py:165-171 + formatter transpose + window cat, as head_mma.cuh);
// dfield = [K][B][D][HW] fp32.
struct EmitParams {
    const float* dfield;
    int K; float fi;
    const float* u_ring; const int* fcount; const int* n_arr;
    float* frames; int n_cap;
    const RolloutPtrs* ptrs; float* ring_out; const int* cum; int n_roll;
    int B, D, T; long long HW;
};
// grid = (ceil(HW / 4 / 256), D, B): thread = 4 consecutive pixels of one (sample, field) plane -- no index divisions,
// 128-bit loads / stores on the channels-first tensors
__global__ void __launch_bounds__(256) taylor_emit_kernel(EmitParams p) {
    const long long pix = ((long long)blockIdx.x * blockDim.x + threadIdx.x) * 4;
    if (pix >= p.HW) return;
    const int d = blockIdx.y, b = blockIdx.z;
    const int n = p.n_arr[b];
    if (n <= 0) return;
    const int fc = p.fcount ? p.fcount[b] : p.T;
    const long long per = (long long)p.D * p.HW;
    const size_t plane = (size_t)d * p.HW + pix;
    const float4 u0 = *reinterpret_cast<const float4*>(p.u_ring + (size_t)(b * p.T + (fc + p.T - 1) % p.T) * per + plane);
    float4 dk[kMaxOrder];
#pragma unroll
    for (int k = 0; k < kMaxOrder; ++k)
        dk[k] = k < p.K ? *reinterpret_cast<const float4*>(p.dfield + ((size_t)k * p.B + b) * per + plane) : make_float4(0.f, 0.f, 0.f, 0.f);
    const int cum = p.cum ? p.cum[b] : 0;
    float* y_out = p.ptrs ? p.ptrs->y_out : nullptr;
    for (int i = 1; i <= n; ++i) {
        const float dt = (float)i * p.fi;
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
        for (int k = kMaxOrder; k >= 1; --k)
            if (k <= p.K) {
                const float sc = dt / (float)k;
                v.x = (dk[k - 1].x + v.x) * sc; v.y = (dk[k - 1].y + v.y) * sc;
                v.z = (dk[k - 1].z + v.z) * sc; v.w = (dk[k - 1].w + v.w) * sc;
            }
        v.x += u0.x; v.y += u0.y; v.z += u0.z; v.w += u0.w;
        if (p.frames) *reinterpret_cast<float4*>(p.frames + ((size_t)b * p.n_cap + (i - 1)) * per + plane) = v;
        if (y_out) {
            const int fidx = cum + i - 1;
            if (fidx < p.n_roll) {
                float* yp = y_out + (((size_t)b * p.n_roll + fidx) * p.HW + pix) * p.D + d;
                yp[0] = v.x; yp[p.D] = v.y; yp[2 * p.D] = v.z; yp[3 * p.D] = v.w;
            }
            if (i > n - p.T) *reinterpret_cast<float4*>(p.ring_out + (size_t)(b * p.T + (fc + i - 1) % p.T) * per + plane) = v;
        }
    }
}

}  // namespace tante
