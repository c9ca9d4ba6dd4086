// On-device evaluation metrics (reference trainer/metrics.py:53-164).  Every metric of the reference's evaluators
// (MSE, NMSE, L2RE, NNMSE, RMSE, NRMSE, VMSE, VRMSE) is a function of three spatial moments per (batch, frame, field):
// sum (x - y)^2, sum y^2 and sum y over (H, W) of the channels-last tensors -- ONE pass over prediction and target
// (8 B per element, HBM-bound) instead of the reference's sub / pow / mean / std / norm passes per metric.
#pragma once
#include "common.cuh"

namespace tante {

constexpr int kMetricMaxC = 16;

// x, y: f32 [BT][HW][C] channels-last.  out: f64 [BT][C][3] (zeroed by the launcher), accumulated with one atomic per
// block, field and moment.  Thread = pixel (its C values are contiguous); the grid is (chunks of HW, BT).
__global__ void __launch_bounds__(256)
metric_moments_kernel(const float* __restrict__ x, const float* __restrict__ y, long long HW, int C, long long per_chunk,
                      double* __restrict__ out) {
    const long long bt = blockIdx.y;
    const long long p0 = (long long)blockIdx.x * per_chunk;
    const long long p1 = min(HW, p0 + per_chunk);
    const float* xb = x + (size_t)bt * HW * C;
    const float* yb = y + (size_t)bt * HW * C;
    float a[kMetricMaxC][3];
#pragma unroll
    for (int c = 0; c < kMetricMaxC; ++c) { a[c][0] = 0.f; a[c][1] = 0.f; a[c][2] = 0.f; }
    for (long long p = p0 + threadIdx.x; p < p1; p += blockDim.x) {
        const float* xp = xb + (size_t)p * C;
        const float* yp = yb + (size_t)p * C;
#pragma unroll
        for (int c = 0; c < kMetricMaxC; ++c) {
            if (c < C) {
                const float yv = __ldg(yp + c), d = __ldg(xp + c) - yv;
                a[c][0] = fmaf(d, d, a[c][0]);
                a[c][1] = fmaf(yv, yv, a[c][1]);
                a[c][2] += yv;
            }
        }
    }
    __shared__ float red[8][kMetricMaxC * 3];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
#pragma unroll
    for (int c = 0; c < kMetricMaxC; ++c) {
        if (c < C) {
#pragma unroll
            for (int j = 0; j < 3; ++j) {
                const float s = warp_sum(a[c][j]);
                if (lane == 0) red[warp][c * 3 + j] = s;
            }
        }
    }
    __syncthreads();
    if ((int)threadIdx.x < C * 3) {
        double s = 0.0;
        for (int w = 0; w < (int)(blockDim.x >> 5); ++w) s += (double)red[w][threadIdx.x];
        atomicAdd(out + (size_t)bt * C * 3 + threadIdx.x, s);
    }
}

static cudaError_t launch_metric_moments(const float* x, const float* y, long long BT, long long HW, int C, double* out,
                                         int num_sms, cudaStream_t st) {
    if (C < 1 || C > kMetricMaxC || BT < 1 || HW < 1 || BT > 65535) return cudaErrorInvalidValue;
    cudaError_t e = cudaMemsetAsync(out, 0, (size_t)BT * C * 3 * sizeof(double), st);
    if (e != cudaSuccess) return e;
    // ~8 blocks per SM in total; every thread sees at least 4 pixels
    long long chunks = std::max<long long>(1, (8LL * num_sms + BT - 1) / BT);
    chunks = std::min<long long>(chunks, (HW + 1023) / 1024);
    const long long per_chunk = (HW + chunks - 1) / chunks;
    chunks = (HW + per_chunk - 1) / per_chunk;
    metric_moments_kernel<<<dim3((unsigned)chunks, (unsigned)BT), 256, 0, st>>>(x, y, HW, C, per_chunk, out);
    return cudaGetLastError();
}

}  // namespace tante
