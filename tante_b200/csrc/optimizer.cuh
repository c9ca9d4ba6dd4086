// Optimizer tail of a training step on the library's flat gradient (`allreduce_grads`, and the clip + AdamW that
// follow it in the reference drivers -- trainer/trainer.py:192-198 `clip_grad_norm_(1.0)`, trainer/r_trainer.py:155-157
// `clip_grad_value_(1.0)`, then `optimizer.step()` of torch.optim.AdamW, configs/tante.yaml:38-41):
//
//   [all-reduce of the flat bucket over NCCL]  ->  sum of squares  ->  clip + AdamW on every parameter through the bound
//   master pointers (+ the clipped gradient written back, as clip_grad_* leaves it)  ->  repack (tante_pack_params)
//
// = the NCCL kernel + two launches + the repack, instead of ~30 multi-tensor launches.  The moments live in two caller-owned
// flat buffers laid out like the gradient.  NCCL is reached through dlopen of the library the process already has (torch
// bundles one); nothing links against it.
#pragma once
#include <dlfcn.h>

#include "common.cuh"

namespace tante {

struct OptSeg { long long off, numel; float* ptr; };
constexpr int kOptMaxSegs = 1024;

__global__ void __launch_bounds__(256) grad_sumsq_kernel(const float* __restrict__ g, long long n, float scale, double* __restrict__ out) {
    float acc = 0.f;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        const float v = g[i] * scale;
        acc = fmaf(v, v, acc);
    }
    __shared__ float red[8];
    acc = warp_sum(acc);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
    __syncthreads();
    if (threadIdx.x == 0) {
        double t = 0.0;
        for (int w = 0; w < 8; ++w) t += (double)red[w];
        atomicAdd(out, t);
    }
}

struct AdamWArgs {
    float gscale;            // applied to the gradient first (1 / world size after a sum all-reduce)
    int clip_mode;           // 0 none, 1 global norm (clip_grad_norm_), 2 value (clip_grad_value_)
    float clip;
    float lr, beta1, beta2, eps, weight_decay;
    float bc1, bc2_sqrt;     // 1 - beta1^step, sqrt(1 - beta2^step)
};

// One thread per flat element; the parameter a flat index belongs to is found by bisection over the segment offsets.
__global__ void __launch_bounds__(256) adamw_flat_kernel(const OptSeg* __restrict__ segs, int nseg, float* __restrict__ g,
                                                         float* __restrict__ m, float* __restrict__ v, long long n,
                                                         const double* __restrict__ sumsq, AdamWArgs a) {
    __shared__ long long s_off[kOptMaxSegs];
    for (int i = threadIdx.x; i < nseg; i += blockDim.x) s_off[i] = segs[i].off;
    __syncthreads();
    float coef = a.gscale;
    if (a.clip_mode == 1) {
        const float norm = (float)sqrt(*sumsq);
        coef *= fminf(1.0f, a.clip / (norm + 1e-6f));
    }
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        int lo = 0, hi = nseg - 1;
        while (lo < hi) {
            const int mid = (lo + hi + 1) >> 1;
            if (s_off[mid] <= i) lo = mid; else hi = mid - 1;
        }
        const OptSeg sg = segs[lo];
        const long long j = i - sg.off;
        if (j >= sg.numel) continue;          // padding between segments
        float gi = g[i] * coef;
        if (a.clip_mode == 2) gi = fminf(fmaxf(gi, -a.clip), a.clip);
        g[i] = gi;
        const float mi = fmaf(1.0f - a.beta1, gi - m[i], m[i]);                  // exp_avg.lerp_(grad, 1 - beta1)
        const float vi = fmaf(1.0f - a.beta2, gi * gi, a.beta2 * v[i]);          // exp_avg_sq.mul_(beta2).addcmul_(grad, grad, 1 - beta2)
        m[i] = mi; v[i] = vi;
        float p = sg.ptr[j];
        p *= 1.0f - a.lr * a.weight_decay;                                      // decoupled weight decay
        const float denom = sqrtf(vi) / a.bc2_sqrt + a.eps;
        sg.ptr[j] = p - (a.lr / a.bc1) * (mi / denom);
    }
}

// ---- NCCL through dlopen -----------------------------------------------------------------------------------------------
struct NcclId { char b[128]; };           // ncclUniqueId (NCCL_UNIQUE_ID_BYTES = 128), passed by value
struct NcclApi {
    void* lib = nullptr;
    int (*GetUniqueId)(NcclId*) = nullptr;
    int (*CommInitRank)(void**, int, NcclId, int) = nullptr;
    int (*AllReduce)(const void*, void*, size_t, int, int, void*, cudaStream_t) = nullptr;
    int (*CommDestroy)(void*) = nullptr;
    const char* (*GetErrorString)(int) = nullptr;
    bool ok() const { return lib && GetUniqueId && CommInitRank && AllReduce && CommDestroy; }
};

static NcclApi& nccl_api() {
    static NcclApi api;
    static bool tried = false;
    if (tried) return api;
    tried = true;
    const char* names[] = {"libnccl.so.2", "libnccl.so"};
    for (const char* n : names) {       // the copy the process already loaded (torch's) first, then the system one
        if ((api.lib = dlopen(n, RTLD_NOW | RTLD_NOLOAD))) break;
    }
    for (const char* n : names) {
        if (api.lib) break;
        api.lib = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
    }
    if (!api.lib) return api;
    api.GetUniqueId = reinterpret_cast<decltype(api.GetUniqueId)>(dlsym(api.lib, "ncclGetUniqueId"));
    api.CommInitRank = reinterpret_cast<decltype(api.CommInitRank)>(dlsym(api.lib, "ncclCommInitRank"));
    api.AllReduce = reinterpret_cast<decltype(api.AllReduce)>(dlsym(api.lib, "ncclAllReduce"));
    api.CommDestroy = reinterpret_cast<decltype(api.CommDestroy)>(dlsym(api.lib, "ncclCommDestroy"));
    api.GetErrorString = reinterpret_cast<decltype(api.GetErrorString)>(dlsym(api.lib, "ncclGetErrorString"));
    return api;
}

}  // namespace tante
