// enc_dec_type = 'fno' (reference models/enc_dec_fno.py:184-323): SpectralLayer = rfft2 (ortho) -> the low modes (the first and
// last m1 rows, the first m2 columns of the half spectrum) mixed across channels by a complex weight -> irfft2 (ortho), plus a
// 1x1 convolution.  Only 2 m1 x m2 modes survive, so both transforms are TRUNCATED DFTs -- five small passes, no FFT:
//     A[n,c,h,k2]  = sum_w x[n,c,h,w] e^{-2 pi i k2 w / W}                         (dft_w)
//     X[n,c,j,k2]  = sum_h A[n,c,h,k2] e^{-2 pi i k1(j) h / H}                     (dft_h; k1(j) = j or H - 2 m1 + j)
//     Y[n,o,j,k2]  = sum_c X[n,c,j,k2] Wt[c,o,j mod m1,k2]                         (mix)
//     B[n,o,h,k2]  = sum_j Y[n,o,j,k2] e^{+2 pi i k1(j) h / H}                     (idft_h)
//     y[n,o,h,w]   = (1 / HW) (Re B[..,0] + 2 sum_{k2>=1} Re(B[..,k2] e^{+2 pi i k2 w / W})) + sum_c w0[o,c] x[n,c,h,w] + b[o]
// (the C2R half of irfft2 ignores the imaginary part of the k2 = 0 column).  All five in fp32 in both precision modes; the
// patch convolutions between the spectral layers are the gather + GEMM stages of wide_patch.cuh.  The training passes at the end
// of the file are the adjoints of the same five passes.
#pragma once
#include "common.cuh"

namespace tante {

// where a spectral layer reads x[n, c, h, w] from / writes y[n, o, h, w] to
struct SpecView {
    const void* p;          // fp32 channels-first frames (mode 0) or TA channels-last grid [n][h][w][c] (mode 1)
    int mode;
    const int* fcount;      // mode 0: ring position per sample (nullable); frame t of sample b sits in slot (fcount[b] + t) % T
    int T;
};

template <typename TA>
__device__ __forceinline__ float spec_read(const SpecView& v, long long n, int c, int h, int w, int C, int H, int W) {
    if (v.mode == 0) {
        long long img = n;
        if (v.fcount) { const long long b = n / v.T; const int t = (int)(n % v.T); img = b * v.T + (v.fcount[b] + t) % v.T; }
        return reinterpret_cast<const float*>(v.p)[((size_t)img * C + c) * H * W + (size_t)h * W + w];
    }
    return to_f32(reinterpret_cast<const TA*>(v.p)[(((size_t)n * H + h) * W + w) * C + c]);
}

// twiddle table tw[j] = (cos, sin)(2 pi j / N), j < N (written once per axis length in double precision on the host)
template <typename TA>
__global__ void __launch_bounds__(128) spec_dft_w_kernel(SpecView in, int C, int H, int W, int m2, const float2* __restrict__ twW,
                                                         float2* __restrict__ A, long long rows) {
    // one warp per (n, c, h) row; lane = k2 (m2 <= 32 per pass, looped otherwise)
    const long long row = (long long)blockIdx.x * 4 + (threadIdx.x >> 5);
    if (row >= rows) return;
    const int lane = threadIdx.x & 31;
    const int h = (int)(row % H);
    const int c = (int)((row / H) % C);
    const long long n = row / ((long long)H * C);
    for (int k2 = lane; k2 < m2; k2 += 32) {
        float re = 0.f, im = 0.f;
        int idx = 0;                                   // (k2 * w) mod W
        for (int w = 0; w < W; ++w) {
            const float v = spec_read<TA>(in, n, c, h, w, C, H, W);
            const float2 t = twW[idx];
            re = fmaf(v, t.x, re);
            im = fmaf(-v, t.y, im);
            idx += k2; if (idx >= W) idx -= W;
        }
        A[(size_t)row * m2 + k2] = make_float2(re, im);
    }
}

__device__ __forceinline__ int spec_k1(int j, int m1, int H) { return j < m1 ? j : H - 2 * m1 + j; }

__global__ void __launch_bounds__(256) spec_dft_h_kernel(const float2* __restrict__ A, int H, int m1, int m2,
                                                         const float2* __restrict__ twH, float2* __restrict__ X, long long total) {
    const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;      // ((nc * 2 m1 + j) * m2 + k2)
    if (idx >= total) return;
    const int k2 = (int)(idx % m2);
    const int j = (int)((idx / m2) % (2 * m1));
    const long long nc = idx / ((long long)m2 * 2 * m1);
    const int k1 = spec_k1(j, m1, H);
    const float2* a = A + (size_t)nc * H * m2 + k2;
    float re = 0.f, im = 0.f;
    int t = 0;                                          // (k1 * h) mod H
    for (int h = 0; h < H; ++h) {
        const float2 v = a[(size_t)h * m2];
        const float2 w = twH[t];                        // e^{-i theta} = (cos, -sin)
        re = fmaf(v.x, w.x, fmaf(v.y, w.y, re));
        im = fmaf(v.y, w.x, fmaf(-v.x, w.y, im));
        t += k1; if (t >= H) t -= H;
    }
    X[idx] = make_float2(re, im);
}

__global__ void __launch_bounds__(256) spec_mix_kernel(const float2* __restrict__ X, const float2* __restrict__ Wt, int Cin, int Cout,
                                                       int m1, int m2, int wm2 /* stored modes2 of the weight */, int wm1,
                                                       float2* __restrict__ Y, long long total) {
    const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;      // ((n * Cout + o) * 2 m1 + j) * m2 + k2
    if (idx >= total) return;
    const int k2 = (int)(idx % m2);
    const int j = (int)((idx / m2) % (2 * m1));
    const int o = (int)((idx / ((long long)m2 * 2 * m1)) % Cout);
    const long long n = idx / ((long long)m2 * 2 * m1 * Cout);
    const int jw = j < m1 ? j : j - m1;
    float re = 0.f, im = 0.f;
    for (int c = 0; c < Cin; ++c) {
        const float2 x = X[(((size_t)n * Cin + c) * 2 * m1 + j) * m2 + k2];
        const float2 w = Wt[(((size_t)c * Cout + o) * wm1 + jw) * wm2 + k2];
        re = fmaf(x.x, w.x, fmaf(-x.y, w.y, re));
        im = fmaf(x.x, w.y, fmaf(x.y, w.x, im));
    }
    Y[idx] = make_float2(re, im);
}

__global__ void __launch_bounds__(256) spec_idft_h_kernel(const float2* __restrict__ Y, int H, int m1, int m2,
                                                          const float2* __restrict__ twH, float2* __restrict__ Bh, long long total) {
    const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;      // ((no * H + h) * m2 + k2)
    if (idx >= total) return;
    const int k2 = (int)(idx % m2);
    const int h = (int)((idx / m2) % H);
    const long long no = idx / ((long long)m2 * H);
    const float2* y = Y + (size_t)no * 2 * m1 * m2 + k2;
    float re = 0.f, im = 0.f;
    for (int j = 0; j < 2 * m1; ++j) {
        const int k1 = spec_k1(j, m1, H);
        const float2 w = twH[(int)(((long long)k1 * h) % H)];      // e^{+i theta}
        const float2 v = y[(size_t)j * m2];
        re = fmaf(v.x, w.x, fmaf(-v.y, w.y, re));
        im = fmaf(v.x, w.y, fmaf(v.y, w.x, im));
    }
    Bh[idx] = make_float2(re, im);
}

// y = irfft-along-W of Bh (scaled 1 / HW) + w0 x + b, optional GELU; out: TA channels-last grid (o fastest) or fp32 channels-first
template <typename TA, bool ACT, bool FIELD>
__global__ void __launch_bounds__(256) spec_out_kernel(const float2* __restrict__ Bh, SpecView in, const float* __restrict__ w0,
                                                       const float* __restrict__ b0, int Cin, int Cout, int H, int W, int m2,
                                                       const float2* __restrict__ twW, TA* __restrict__ out, float* __restrict__ field,
                                                       long long total, int conv_done = 0) {
    const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= total) return;
    int o, w, h;
    long long n;
    // thread order: w fastest in both cases, so that a warp shares one Bh row (broadcast loads) -- with the channel fastest every
    // thread walked its own row (32 cache lines per load); the channels-last stores are strided instead, a far smaller cost
    if (FIELD) { long long r = idx; w = (int)(r % W); r /= W; h = (int)(r % H); r /= H; o = (int)(r % Cout); n = r / Cout; }
    else { long long r = idx; w = (int)(r % W); r /= W; o = (int)(r % Cout); r /= Cout; h = (int)(r % H); n = r / H; }
    const float2* bh = Bh + (((size_t)n * Cout + o) * H + h) * m2;
    float acc = bh[0].x;
    int t = 0;
    for (int k2 = 1; k2 < m2; ++k2) {
        t += w; if (t >= W) t -= W;                     // (k2 * w) mod W
        const float2 tw = twW[t];
        const float2 v = bh[k2];
        acc = fmaf(2.f, fmaf(v.x, tw.x, -v.y * tw.y), acc);
    }
    float v = acc / ((float)H * (float)W);
    // 1x1 conv + bias: computed here (thin layers), or already sitting in `out` (conv_done: a GEMM wrote it -- run_spectral)
    float s;
    if (conv_done) {
        s = to_f32(out[(((size_t)n * H + h) * W + w) * Cout + o]);
    } else {
        s = b0[o];
        for (int c = 0; c < Cin; ++c) s = fmaf(w0[(size_t)o * Cin + c], spec_read<TA>(in, n, c, h, w, Cin, H, W), s);
    }
    v += s;
    if (ACT) v = ActMath<TA>::gelu_erf_f(v);
    if (FIELD) field[idx] = v;
    else out[(((size_t)n * H + h) * W + w) * Cout + o] = from_f32<TA>(v);
}

// ---- training: adjoints of the five passes ----------------------------------------------------------------------------------
// With g = dL/dy (after the GELU backward), every pass above is real-linear, so the input gradient runs the same transforms in
// reverse order on g:
//     gB[n,o,h,k2] = sum_w g[n,o,h,w] e^{-2 pi i k2 w / W}                          (dft_w of g; the weights c_k2 / HW, c_0 = 1,
//     gY[n,o,j,k2] = sum_h gB[n,o,h,k2] e^{-2 pi i k1(j) h / H}                      c_k = 2, are applied where gY is consumed)
//     gX[n,c,j,k2] = (c_k2 / HW) sum_o gY[n,o,j,k2] conj(Wt[c,o,j mod m1,k2])       (spec_mix_adj_kernel)
//     gA[n,c,h,k2] = sum_j gX[n,c,j,k2] e^{+2 pi i k1(j) h / H}                      (idft_h)
//     dx[n,c,h,w]  = sum_k2 Re(gA[n,c,h,k2] e^{+2 pi i k2 w / W}) + sum_o w0[o,c] g[n,o,h,w]     (spec_in_bwd_kernel)
// and the parameter gradients are  dWt[c,o,jw,k2] = (c_k2 / HW) sum_{n, j = jw, jw + m1} conj(X[n,c,j,k2]) gY[n,o,j,k2]  (PyTorch's
// convention for complex leaves: real part = d/dRe, imaginary part = d/dIm),  dw0[o,c] = sum g[n,o,h,w] x[n,c,h,w],  db[o] = sum g.

__device__ __forceinline__ float spec_ck(int k2, int H, int W) { return (k2 == 0 ? 1.f : 2.f) / ((float)H * (float)W); }

__global__ void __launch_bounds__(256) spec_mix_adj_kernel(const float2* __restrict__ gY, const float2* __restrict__ Wt, int Cin, int Cout,
                                                           int m1, int m2, int wm2, int wm1, int H, int W, float2* __restrict__ gX,
                                                           long long total) {
    const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;      // ((n * Cin + c) * 2 m1 + j) * m2 + k2
    if (idx >= total) return;
    const int k2 = (int)(idx % m2);
    const int j = (int)((idx / m2) % (2 * m1));
    const int c = (int)((idx / ((long long)m2 * 2 * m1)) % Cin);
    const long long n = idx / ((long long)m2 * 2 * m1 * Cin);
    const int jw = j < m1 ? j : j - m1;
    float re = 0.f, im = 0.f;
    for (int o = 0; o < Cout; ++o) {
        const float2 g = gY[(((size_t)n * Cout + o) * 2 * m1 + j) * m2 + k2];
        const float2 w = Wt[(((size_t)c * Cout + o) * wm1 + jw) * wm2 + k2];
        re = fmaf(g.x, w.x, fmaf(g.y, w.y, re));          // g * conj(w)
        im = fmaf(g.y, w.x, fmaf(-g.x, w.y, im));
    }
    const float s = spec_ck(k2, H, W);
    gX[idx] = make_float2(re * s, im * s);
}

__global__ void __launch_bounds__(256) spec_wt_grad_kernel(const float2* __restrict__ X, const float2* __restrict__ gY, int Cin, int Cout,
                                                           int m1, int m2, int wm2, int wm1, int H, int W, long long N,
                                                           float2* __restrict__ gWt, long long total) {
    const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;      // ((c * Cout + o) * m1 + jw) * m2 + k2
    if (idx >= total) return;
    const int k2 = (int)(idx % m2);
    const int jw = (int)((idx / m2) % m1);
    const int o = (int)((idx / ((long long)m2 * m1)) % Cout);
    const int c = (int)(idx / ((long long)m2 * m1 * Cout));
    float re = 0.f, im = 0.f;
    for (long long n = 0; n < N; ++n) {
#pragma unroll
        for (int half = 0; half < 2; ++half) {
            const int j = jw + half * m1;
            const float2 x = X[(((size_t)n * Cin + c) * 2 * m1 + j) * m2 + k2];
            const float2 g = gY[(((size_t)n * Cout + o) * 2 * m1 + j) * m2 + k2];
            re = fmaf(x.x, g.x, fmaf(x.y, g.y, re));      // conj(x) * g
            im = fmaf(x.x, g.y, fmaf(-x.y, g.x, im));
        }
    }
    const float s = spec_ck(k2, H, W);
    float2* dst = gWt + (((size_t)c * Cout + o) * wm1 + jw) * wm2 + k2;
    dst->x += re * s;
    dst->y += im * s;
}

// dx of one spectral layer: the C2R-like pass over gA plus the 1x1-conv path.  Output: TA channels-last grid dx[n][h][w][c]
// (written) or, for the first encoder layer, the fp32 channels-first input gradient (ACCUMULATED: gin[((n * Cin + c) * H + h) * W + w]).
template <typename TA, bool CF>
__global__ void __launch_bounds__(256) spec_in_bwd_kernel(const float2* __restrict__ gA, SpecView g, const float* __restrict__ w0, int Cin,
                                                          int Cout, int H, int W, int m2, const float2* __restrict__ twW,
                                                          TA* __restrict__ dx, float* __restrict__ gin, long long total) {
    const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= total) return;
    int c, w, h;
    long long n;
    if (CF) { long long r = idx; w = (int)(r % W); r /= W; h = (int)(r % H); r /= H; c = (int)(r % Cin); n = r / Cin; }
    else { long long r = idx; w = (int)(r % W); r /= W; c = (int)(r % Cin); r /= Cin; h = (int)(r % H); n = r / H; }      // (w fastest: see spec_out_kernel)
    const float2* ga = gA + (((size_t)n * Cin + c) * H + h) * m2;
    float acc = ga[0].x;
    int t = 0;
    for (int k2 = 1; k2 < m2; ++k2) {
        t += w; if (t >= W) t -= W;
        const float2 tw = twW[t];
        const float2 v = ga[k2];
        acc += fmaf(v.x, tw.x, -v.y * tw.y);
    }
    for (int o = 0; o < Cout; ++o) acc = fmaf(w0[(size_t)o * Cin + c], spec_read<TA>(g, n, o, h, w, Cout, H, W), acc);
    if (CF) gin[idx] += acc;
    else dx[(((size_t)n * H + h) * W + w) * Cin + c] = from_f32<TA>(acc);
}

// dw0[o][c] += sum_{n,h,w} g[n,o,h,w] x[n,c,h,w] and db[o] += sum g  for the thin layers (a field or frame side: few channels);
// one CTA per (o, c), c == Cin computes the bias entry
template <typename TA>
__global__ void __launch_bounds__(256) spec_w0_grad_kernel(SpecView g, SpecView x, int Cin, int Cout, int H, int W, long long N,
                                                           float* __restrict__ gw0, float* __restrict__ gb0) {
    const int o = blockIdx.x / (Cin + 1), c = blockIdx.x % (Cin + 1);
    const long long total = N * H * W;
    float acc = 0.f;
    for (long long i = threadIdx.x; i < total; i += blockDim.x) {
        const int w = (int)(i % W);
        const int h = (int)((i / W) % H);
        const long long n = i / ((long long)W * H);
        const float gv = spec_read<TA>(g, n, o, h, w, Cout, H, W);
        acc = c < Cin ? fmaf(gv, spec_read<TA>(x, n, c, h, w, Cin, H, W), acc) : acc + gv;
    }
    __shared__ float red[8];
    acc = warp_sum(acc);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
    __syncthreads();
    if (threadIdx.x == 0) {
        float s = 0.f;
        for (int i = 0; i < 8; ++i) s += red[i];
        if (c < Cin) gw0[(size_t)o * Cin + c] += s;
        else gb0[o] += s;
    }
}

}  // namespace tante
