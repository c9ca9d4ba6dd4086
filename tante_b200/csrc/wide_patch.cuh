// patch_scale 16 / 32 / 64 (reference enc_dec_cnn.py:39-46,75-81,93-95,176-184): Patch_map stages with a 4x4 kernel use
// padding (k-1)//2 = 1, so their windows are SHIFTED by one pixel against the patch grid (window i covers rows 4i-1 .. 4i+2,
// zero-padded at the top / left, the last row / column unused), and the matching transposed convs produce 4h-2 rows that
// a bilinear (align_corners=False) resize stretches back to 4h.  The nested pixel order that turns the P <= 8 stages into
// plain GEMMs on reshaped matrices does not survive a shifted window or a resize that blends neighbouring patches, so
// these scales keep every stage in NATURAL grid order (channels-last [b][h_s][w_s][C_s]) and run
//     encoder stage = gather of the (shifted) k x k windows into a patch matrix  ->  GEMM (+ GELU / embed epilogue)
//     decoder stage = GEMM to the k x k sub-pixel matrix  ->  crop + bilinear resample (+ bias, GELU) back to the grid
//     head          = Horner sum + residual + emit over the decoded derivative FIELDS (boundary C of SURVEY.md 8(d))
// The kernels are plain gather passes (HBM-bound, one pass each); the training passes further down are their transposes.
#pragma once
#include "common.cuh"
#include "kernels_simt.cuh"

namespace tante {

// conv1 windows from the channels-first (ring) frames: out[row = (bt, i, j)][(di*k + dj)*D + d], zero beyond k*k*D
template <typename TA>
__global__ void __launch_bounds__(256)
wide_im2col_cf_kernel(const float* __restrict__ x, const int* __restrict__ fcount, int T, int D, int H, int W, int k, int shift,
                      int Kpad, TA* __restrict__ out, long long total, int stride = 0, int Hc = 0, int Wc = 0) {
    const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= total) return;
    const int col = (int)(idx % Kpad);
    long long row = idx / Kpad;
    float v = 0.f;
    if (col < k * k * D) {
        // windows: stride == k on the (H / k, W / k) grid, or (overlap_ratio != 0) stride < k on the (Hc, Wc) conv grid
        const int sd = stride > 0 ? stride : k;
        const int Ho = stride > 0 ? Hc : H / k, Wo = stride > 0 ? Wc : W / k;
        const int d = col % D, tap = col / D;
        const int di = tap / k, dj = tap % k;
        const int j = (int)(row % Wo); row /= Wo;
        const int i = (int)(row % Ho);
        const long long bt = row / Ho;
        const long long b = bt / T;
        const int t = (int)(bt % T);
        const int slot = fcount ? (fcount[b] + t) % T : t;
        const int y = i * sd + di - shift, xx = j * sd + dj - shift;
        if (y >= 0 && y < H && xx >= 0 && xx < W)
            v = x[((size_t)(b * T + slot) * D + d) * H * W + (size_t)y * W + xx];
    }
    out[idx] = from_f32<TA>(v);
}

// conv2 / conv3 windows from a channels-last grid [n][Hs][Ws][Cin]: out[row = (n, i, j)][(di*k + dj)*Cin + ci]; thread = 4 channels
template <typename TA>
__global__ void __launch_bounds__(256)
wide_im2col_cl_kernel(const TA* __restrict__ in, int Hs, int Ws, int Cin, int k, int shift, TA* __restrict__ out, long long total4,
                      int stride = 0, int Hc = 0, int Wc = 0) {
    const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= total4) return;
    const int C4 = Cin / 4;
    const int c4 = (int)(idx % C4);
    long long r = idx / C4;
    const int tap = (int)(r % (k * k)); r /= k * k;
    const int sd = stride > 0 ? stride : k;
    const int Ho = stride > 0 ? Hc : Hs / k, Wo = stride > 0 ? Wc : Ws / k;
    const int j = (int)(r % Wo); r /= Wo;
    const int i = (int)(r % Ho);
    const long long n = r / Ho;
    const int di = tap / k, dj = tap % k;
    const int y = i * sd + di - shift, xx = j * sd + dj - shift;
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
                                           int y, int x, int co, int sd = 0) {
    const int yy = y + pad, xx = x + pad;
    if (sd <= 0 || sd == k) {
        const int i = yy / k, di = yy - i * k, j = xx / k, dj = xx - j * k;
        return to_f32(S[(((size_t)n * hi + i) * wi + j) * ldS + (di * k + dj) * Cout + co]);
    }
    // overlap_ratio != 0: stride sd < k, every output sample sums the taps di = yy mod sd, + sd, ... of the inputs (yy - di) / sd
    float v = 0.f;
    for (int di = yy % sd; di < k; di += sd) {
        const int i = (yy - di) / sd;
        if (i < 0 || i >= hi) continue;
        for (int dj = xx % sd; dj < k; dj += sd) {
            const int j = (xx - dj) / sd;
            if (j < 0 || j >= wi) continue;
            v += to_f32(S[(((size_t)n * hi + i) * wi + j) * ldS + (di * k + dj) * Cout + co]);
        }
    }
    return v;
}

// decoder stage tail: sub-pixel matrix (bias already added by the GEMM epilogue when `bias` is null) -> [crop + bilinear resize
// for k = 4] -> (+ bias) -> (GELU) -> channels-last grid out[n][k*hi][k*wi][Cout]  or, for the last stage, the channels-first
// fp32 derivative field field[n][Cout][k*hi][k*wi].
template <typename TA, bool ACT, bool FIELD>
__global__ void __launch_bounds__(256)
wide_deconv_post_kernel(const TA* __restrict__ S, int ldS, int hi, int wi, int Cout, int k, const float* __restrict__ bias,
                        TA* __restrict__ out, float* __restrict__ field, long long total, int sd = 0) {
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
    const int se = sd > 0 ? sd : k;
    const int Hd = (hi - 1) * se - 2 * pad + k, Wd = (wi - 1) * se - 2 * pad + k;       // what ConvTranspose2d produced (enc_dec_cnn.py:162-166)
    float v;
    if (Hd == Ho && Wd == Wo) {
        v = deconv_at(S, ldS, n, hi, wi, Cout, k, pad, Y, X, co, sd);
    } else {
        int y0, y1, x0, x1;
        float wy, wx;
        bilinear_src(Y, Hd, Ho, y0, y1, wy);
        bilinear_src(X, Wd, Wo, x0, x1, wx);
        const float v00 = deconv_at(S, ldS, n, hi, wi, Cout, k, pad, y0, x0, co, sd), v01 = deconv_at(S, ldS, n, hi, wi, Cout, k, pad, y0, x1, co, sd);
        const float v10 = deconv_at(S, ldS, n, hi, wi, Cout, k, pad, y1, x0, co, sd), v11 = deconv_at(S, ldS, n, hi, wi, Cout, k, pad, y1, x1, co, sd);
        // aten upsample_bilinear2d: h0lambda * (w0lambda * v00 + w1lambda * v01) + h1lambda * (w0lambda * v10 + w1lambda * v11)
        v = (1.f - wy) * ((1.f - wx) * v00 + wx * v01) + wy * ((1.f - wx) * v10 + wx * v11);
    }
    if (bias) v += bias[co];
    if (ACT) v = ActMath<TA>::gelu_erf_f(v);
    if (FIELD) field[idx] = v;
    else out[idx] = from_f32<TA>(v);
}

// adaptive_avg_pool2d of the conv grid [n][Hc][Wc][C] (TA) to the patch grid [n][Ho][Wo][C] (enc_dec_cnn.py:109; torch's windows:
// rows floor(i Hc / Ho) .. ceil((i + 1) Hc / Ho) - 1), then the stage's GELU; output TA, or fp32 (OUT32: the pre-embedding of the
// last stage).  Thread = 4 channels.
template <typename TA, bool ACT, bool OUT32>
__global__ void __launch_bounds__(256)
wide_pool_kernel(const TA* __restrict__ in, int Hc, int Wc, int C, int Ho, int Wo, TA* __restrict__ out, float* __restrict__ out32,
                 long long total4) {
    const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= total4) return;
    const int C4 = C / 4;
    const int c4 = (int)(idx % C4);
    long long r = idx / C4;
    const int j = (int)(r % Wo); r /= Wo;
    const int i = (int)(r % Ho);
    const long long n = r / Ho;
    const int y0 = (i * Hc) / Ho, y1 = ((i + 1) * Hc + Ho - 1) / Ho;
    const int x0 = (j * Wc) / Wo, x1 = ((j + 1) * Wc + Wo - 1) / Wo;
    float acc[4] = {0.f, 0.f, 0.f, 0.f};
    for (int y = y0; y < y1; ++y)
        for (int x = x0; x < x1; ++x) {
            float v[4];
            Vec4<TA>::load(in + (((size_t)n * Hc + y) * Wc + x) * C + c4 * 4, v);
#pragma unroll
            for (int e = 0; e < 4; ++e) acc[e] += v[e];
        }
    const float inv = 1.0f / (float)((y1 - y0) * (x1 - x0));
#pragma unroll
    for (int e = 0; e < 4; ++e) {
        acc[e] *= inv;
        if (ACT) acc[e] = ActMath<TA>::gelu_erf_f(acc[e]);
    }
    if (OUT32) Vec4<float>::store(out32 + idx * 4, acc);
    else Vec4<TA>::store(out + idx * 4, acc);
}

// fp32 accumulator of a split-K GEMM -> (GELU) -> TA; thread = 4 elements
template <typename TA, bool ACT>
__global__ void __launch_bounds__(256) f32_to_ta_kernel(const float* __restrict__ src, TA* __restrict__ dst, long long n4) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n4) return;
    float v[4];
    Vec4<float>::load(src + i * 4, v);
    if (ACT) {
#pragma unroll
        for (int j = 0; j < 4; ++j) v[j] = ActMath<TA>::gelu_erf_f(v[j]);
    }
    Vec4<TA>::store(dst + i * 4, v);
}

// ---- training (backward) passes of the natural-order stages ------------------------------------------------------
// Every forward pass above is a gather, so its transpose is written as a gather too (no atomics): with stride == kernel size a
// grid pixel belongs to at most ONE window, and a transposed-conv sample feeds at most 4 x 4 resized pixels.

// transpose of wide_im2col_cl_kernel: grid gradient [n][Hs][Ws][Cin] from the patch-matrix gradient; thread = 4 channels
template <typename TA>
__global__ void __launch_bounds__(256)
wide_col2im_cl_kernel(const TA* __restrict__ dcols, int Hs, int Ws, int Cin, int k, int shift, TA* __restrict__ dgrid, long long total4,
                      int stride = 0, int Hc = 0, int Wc = 0) {
    const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= total4) return;
    const int C4 = Cin / 4;
    const int c4 = (int)(idx % C4);
    long long r = idx / C4;
    const int x = (int)(r % Ws); r /= Ws;
    const int y = (int)(r % Hs);
    const long long n = r / Hs;
    const int yy = y + shift, xx = x + shift;
    float v[4] = {0.f, 0.f, 0.f, 0.f};
    if (stride <= 0 || stride == k) {
        const int Ho = Hs / k, Wo = Ws / k;
        const int i = yy / k, di = yy - i * k, j = xx / k, dj = xx - j * k;
        if (i < Ho && j < Wo) Vec4<TA>::load(dcols + ((((size_t)n * Ho + i) * Wo + j) * (k * k) + (di * k + dj)) * Cin + c4 * 4, v);
    } else {
        // overlapping windows: the pixel is tap (di, dj) of every window ((yy - di) / stride, (xx - dj) / stride) on the conv grid
        for (int di = yy % stride; di < k; di += stride) {
            const int i = (yy - di) / stride;
            if (i < 0 || i >= Hc) continue;
            for (int dj = xx % stride; dj < k; dj += stride) {
                const int j = (xx - dj) / stride;
                if (j < 0 || j >= Wc) continue;
                float t4[4];
                Vec4<TA>::load(dcols + ((((size_t)n * Hc + i) * Wc + j) * (k * k) + (di * k + dj)) * Cin + c4 * 4, t4);
#pragma unroll
                for (int e = 0; e < 4; ++e) v[e] += t4[e];
            }
        }
    }
    Vec4<TA>::store(dgrid + idx * 4, v);
}

// transpose of wide_im2col_cf_kernel (contiguous (B, T, D, H, W) input): grad_input += the window-matrix gradient
template <typename TA>
__global__ void __launch_bounds__(256)
wide_col2im_cf_kernel(const TA* __restrict__ dcols, int D, int H, int W, int k, int shift, int Kpad, float* __restrict__ gin,
                      long long total, int stride = 0, int Hc = 0, int Wc = 0) {
    const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;      // ((bt * D + d) * H + y) * W + x
    if (idx >= total) return;
    long long r = idx;
    const int x = (int)(r % W); r /= W;
    const int y = (int)(r % H); r /= H;
    const int d = (int)(r % D);
    const long long bt = r / D;
    const int yy = y + shift, xx = x + shift;
    if (stride <= 0 || stride == k) {
        const int Ho = H / k, Wo = W / k;
        const int i = yy / k, di = yy - i * k, j = xx / k, dj = xx - j * k;
        if (i < Ho && j < Wo) gin[idx] += to_f32(dcols[(((size_t)bt * Ho + i) * Wo + j) * Kpad + (di * k + dj) * D + d]);
        return;
    }
    float v = 0.f;
    for (int di = yy % stride; di < k; di += stride) {
        const int i = (yy - di) / stride;
        if (i < 0 || i >= Hc) continue;
        for (int dj = xx % stride; dj < k; dj += stride) {
            const int j = (xx - dj) / stride;
            if (j < 0 || j >= Wc) continue;
            v += to_f32(dcols[(((size_t)bt * Hc + i) * Wc + j) * Kpad + (di * k + dj) * D + d]);
        }
    }
    gin[idx] += v;
}

// weight with which resized pixel `dst` reads source sample `src` (bilinear_src above; both taps may hit the same sample at the border)
__device__ __forceinline__ float bilinear_w(int dst, int src, int in_size, int out_size) {
    int i0, i1;
    float w1;
    bilinear_src(dst, in_size, out_size, i0, i1, w1);
    return (i0 == src ? 1.f - w1 : 0.f) + (i1 == src ? w1 : 0.f);
}

// transpose of wide_deconv_post_kernel (without its bias / GELU): gradient of the sub-pixel matrix dS[(n, i, j)][ldS] from the
// gradient of the stage output -- a channels-last TA grid g[n][k*hi][k*wi][Cout], or (FIELD) the channels-first fp32 field
// gradient gf[n][Cout][k*hi][k*wi].  Columns beyond k*k*Cout (padding of the last stage) are zeroed.
template <typename TA, bool FIELD>
__global__ void __launch_bounds__(256)
wide_deconv_post_bwd_kernel(const TA* __restrict__ g, const float* __restrict__ gf, int ldS, int hi, int wi, int Cout, int k,
                            TA* __restrict__ dS, long long total, int sd = 0) {
    const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;      // row * ldS + col
    if (idx >= total) return;
    const int col = (int)(idx % ldS);
    long long row = idx / ldS;
    float v = 0.f;
    if (col < k * k * Cout) {
        const int co = col % Cout, tap = col / Cout;
        const int di = tap / k, dj = tap - di * k;
        const int j = (int)(row % wi); row /= wi;
        const int i = (int)(row % hi);
        const long long n = row / hi;
        const int Ho = hi * k, Wo = wi * k;
        const int pad = (k - 1) / 2;
        const int se = sd > 0 ? sd : k;      // (with overlap several (i, di) pairs land on the same sample: each receives its gradient)
        const int y = se * i + di - pad, x = se * j + dj - pad;
        const int Hd = (hi - 1) * se - 2 * pad + k, Wd = (wi - 1) * se - 2 * pad + k;
        auto at = [&](int Y, int X) -> float {
            return FIELD ? gf[(((size_t)n * Cout + co) * Ho + Y) * Wo + X] : to_f32(g[(((size_t)n * Ho + Y) * Wo + X) * Cout + co]);
        };
        if (Hd == Ho && Wd == Wo) {
            if (y >= 0 && y < Ho && x >= 0 && x < Wo) v = at(y, x);
        } else {
            if (y >= 0 && y < Hd && x >= 0 && x < Wd) {
                // resized pixel Y reads the samples floor(src(Y)), + 1 with src(Y) = (Hd / Ho)(Y + 0.5) - 0.5 (clamped at 0): sample y can only
                // be read by Y with src(Y) in (y - 1, y + 1); bilinear_w decides exactly
                const float ry = (float)Ho / (float)Hd, rx = (float)Wo / (float)Wd;
                const int Ylo = max((int)floorf(((float)y - 0.5f) * ry - 0.5f) - 1, 0), Yhi = min((int)ceilf(((float)y + 1.5f) * ry - 0.5f) + 1, Ho - 1);
                const int Xlo = max((int)floorf(((float)x - 0.5f) * rx - 0.5f) - 1, 0), Xhi = min((int)ceilf(((float)x + 1.5f) * rx - 0.5f) + 1, Wo - 1);
                for (int Y = Ylo; Y <= Yhi; ++Y) {
                    const float wy = bilinear_w(Y, y, Hd, Ho);
                    if (wy == 0.f) continue;
                    float acc = 0.f;
                    for (int X = Xlo; X <= Xhi; ++X) {
                        const float wx = bilinear_w(X, x, Wd, Wo);
                        if (wx != 0.f) acc = fmaf(wx, at(Y, X), acc);
                    }
                    v = fmaf(wy, acc, v);
                }
            }
        }
    }
    dS[idx] = from_f32<TA>(v);
}

// transpose of wide_pool_kernel: conv-grid gradient [n][Hc][Wc][C] from the patch-grid gradient g [n][Ho][Wo][C] (TA; or fp32 G32 for
// the last encoder stage, whose pooled output is the fp32 pre-embedding).  A conv-grid pixel may sit in two neighbouring windows
// per axis (torch's adaptive windows overlap when Hc is not a multiple of Ho).  Thread = 4 channels.
template <typename TA, bool G32>
__global__ void __launch_bounds__(256)
wide_pool_bwd_kernel(const TA* __restrict__ g, const float* __restrict__ g32, int Hc, int Wc, int C, int Ho, int Wo, TA* __restrict__ dconv,
                     long long total4) {
    const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= total4) return;
    const int C4 = C / 4;
    const int c4 = (int)(idx % C4);
    long long r = idx / C4;
    const int x = (int)(r % Wc); r /= Wc;
    const int y = (int)(r % Hc);
    const long long n = r / Hc;
    float acc[4] = {0.f, 0.f, 0.f, 0.f};
    const int ilo = max((int)(((long long)y * Ho) / Hc) - 1, 0), ihi = min((int)(((long long)(y + 1) * Ho + Hc - 1) / Hc), Ho - 1);
    const int jlo = max((int)(((long long)x * Wo) / Wc) - 1, 0), jhi = min((int)(((long long)(x + 1) * Wo + Wc - 1) / Wc), Wo - 1);
    for (int i = ilo; i <= ihi; ++i) {
        const int y0 = (i * Hc) / Ho, y1 = ((i + 1) * Hc + Ho - 1) / Ho;
        if (y < y0 || y >= y1) continue;
        for (int j = jlo; j <= jhi; ++j) {
            const int x0 = (j * Wc) / Wo, x1 = ((j + 1) * Wc + Wo - 1) / Wo;
            if (x < x0 || x >= x1) continue;
            const float inv = 1.0f / (float)((y1 - y0) * (x1 - x0));
            float v[4];
            const size_t off = (((size_t)n * Ho + i) * Wo + j) * C + c4 * 4;
            if (G32) Vec4<float>::load(g32 + off, v);
            else Vec4<TA>::load(g + off, v);
#pragma unroll
            for (int e = 0; e < 4; ++e) acc[e] = fmaf(v[e], inv, acc[e]);
        }
    }
    Vec4<TA>::store(dconv + idx * 4, acc);
}

// bias gradient of an overlapped last decoder stage: db[d] += sum over (n, pixels) of the field gradient gf [n][D][HW]
__global__ void __launch_bounds__(256) field_bias_grad_kernel(const float* __restrict__ gf, int D, long long HW, long long N,
                                                              float* __restrict__ gb) {
    const int d = blockIdx.x;
    const long long chunk = blockIdx.y, nchunk = gridDim.y;
    float acc = 0.f;
    for (long long n = 0; n < N; ++n)
        for (long long i = chunk * blockDim.x + threadIdx.x; i < HW; i += nchunk * blockDim.x) acc += gf[((size_t)n * D + d) * HW + i];
    __shared__ float red[8];
    acc = warp_sum(acc);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
    __syncthreads();
    if (threadIdx.x == 0) {
        float t = 0.f;
        for (int i = 0; i < 8; ++i) t += red[i];
        atomicAdd(gb + d, t);
    }
}

// transpose of the Horner emit (training: contiguous window, frames (B, n_cap, D, H, W)): gradient of the K derivative fields
// gk[k-1] = sum_{i <= n_b} gframes_i (i fi)^k / k!  and of u0 (the last window frame): sum_i gframes_i, accumulated into gu0.
struct EmitBwdParams {
    const float* gframes; long long gf_bs;      // frame i of sample b at gframes + b * gf_bs + i * D * HW
    const int* n_arr; int n_g;
    float* gfield;                               // [K][B][D][HW]
    float* gu0; long long gu0_bs;                // nullable; sample stride
    int K; float fi; int B, D; long long HW;
};
__global__ void __launch_bounds__(256) taylor_emit_bwd_kernel(EmitBwdParams p) {
    const long long pix = ((long long)blockIdx.x * blockDim.x + threadIdx.x) * 4;
    if (pix >= p.HW) return;
    const int d = blockIdx.y, b = blockIdx.z;
    const int n = min(p.n_arr[b], p.n_g);
    const long long per = (long long)p.D * p.HW;
    const size_t plane = (size_t)d * p.HW + pix;
    float4 gk[kMaxOrder], g0 = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
    for (int k = 0; k < kMaxOrder; ++k) gk[k] = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int i = 1; i <= n; ++i) {
        const float4 gv = *reinterpret_cast<const float4*>(p.gframes + (size_t)b * p.gf_bs + (size_t)(i - 1) * per + plane);
        g0.x += gv.x; g0.y += gv.y; g0.z += gv.z; g0.w += gv.w;
        const float dt = (float)i * p.fi;
        float c = 1.f;
#pragma unroll
        for (int k = 1; k <= kMaxOrder; ++k)
            if (k <= p.K) {
                c *= dt / (float)k;
                gk[k - 1].x = fmaf(c, gv.x, gk[k - 1].x); gk[k - 1].y = fmaf(c, gv.y, gk[k - 1].y);
                gk[k - 1].z = fmaf(c, gv.z, gk[k - 1].z); gk[k - 1].w = fmaf(c, gv.w, gk[k - 1].w);
            }
    }
#pragma unroll
    for (int k = 0; k < kMaxOrder; ++k)
        if (k < p.K) *reinterpret_cast<float4*>(p.gfield + ((size_t)k * p.B + b) * per + plane) = gk[k];
    if (p.gu0) {
        float4* q = reinterpret_cast<float4*>(p.gu0 + (size_t)b * p.gu0_bs + plane);
        float4 o = *q;
        o.x += g0.x; o.y += g0.y; o.z += g0.z; o.w += g0.w;
        *q = o;
    }
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
