// Counter-based dropout masks (reference: nn.Dropout / nn.MultiheadAttention(dropout=p) in TransformerBlock,
// models/attn_backbone.py:47-57,81-83; p = 0.1 in configs/tante.yaml:29).
//
// A mask bit is a pure function of (per-call key, site, element index): Philox4x32 with 7 rounds keyed by the 64-bit seed
// the host draws for every taped model call, counter = (element group, site).  Nothing is stored: the backward kernels
// regenerate exactly the bits the forward used.  One Philox call yields 8 independent 16-bit lanes = the masks of 8
// consecutive elements; an element is KEPT when its lane >= round(p * 65536) and then scaled by 1 / (1 - p).
// PyTorch's own Philox stream is not reproduced (bit-parity with torch's dropout is out of scope); the
// tests check the keep statistics, forward/backward consistency and the train-mode output distribution vs the reference.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace tante {

struct DropCfg {
    float p = 0.f;          // 0: dropout disabled
    float scale = 1.f;      // 1 / (1 - p)
    uint32_t thr = 0;       // round(p * 65536)
    uint32_t k0 = 0, k1 = 0;    // per-call key
};

// sites of one transformer layer (order o, layer i): 4 * (o * 64 + i) + {0: attention probabilities, 1: residual after the
// attention out-projection, 2: residual after the MLP}
__host__ __device__ inline uint32_t drop_site(int order, int layer, int which) { return (uint32_t)(4 * (order * 64 + layer) + which); }

__device__ __forceinline__ uint4 philox4x32_7(uint4 c, uint32_t k0, uint32_t k1) {
    constexpr uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u, W0 = 0x9E3779B9u, W1 = 0xBB67AE85u;
#pragma unroll
    for (int r = 0; r < 7; ++r) {
        const uint32_t hi0 = __umulhi(M0, c.x), lo0 = M0 * c.x;
        const uint32_t hi1 = __umulhi(M1, c.z), lo1 = M1 * c.z;
        c = make_uint4(hi1 ^ c.y ^ k0, lo1, hi0 ^ c.w ^ k1, lo0);
        k0 += W0; k1 += W1;
    }
    return c;
}

// the four 32-bit words holding the 16-bit lanes of element group `grp` (elements 8 * grp .. 8 * grp + 7) of `site`
__device__ __forceinline__ uint4 drop_words(const DropCfg& d, uint32_t site, unsigned long long grp) {
    return philox4x32_7(make_uint4((uint32_t)grp, (uint32_t)(grp >> 32), site, 0x7A17E0D0u), d.k0, d.k1);
}
__device__ __forceinline__ uint32_t drop_lane(const uint4& w, int lane /* 0..7 */) {
    const uint32_t v = (lane >> 1) == 0 ? w.x : ((lane >> 1) == 1 ? w.y : ((lane >> 1) == 2 ? w.z : w.w));
    return (lane & 1) ? (v >> 16) : (v & 0xFFFFu);
}
// multiplier (0 or 1 / (1 - p)) of one element
__device__ __forceinline__ float drop_mul(const DropCfg& d, const uint4& w, int lane) {
    return drop_lane(w, lane) >= d.thr ? d.scale : 0.f;
}
__device__ __forceinline__ float drop_elem(const DropCfg& d, uint32_t site, unsigned long long elem) {
    const uint4 w = drop_words(d, site, elem >> 3);
    return drop_mul(d, w, (int)(elem & 7));
}
// attention probabilities: element = (query token, head, key position); 8 consecutive key positions share a group
__device__ __forceinline__ unsigned long long drop_attn_grp(long long tok_q, int n_head, int head, int kpos) {
    return ((unsigned long long)(tok_q * n_head + head) << 13) | (unsigned long long)(kpos >> 3);
}

static inline DropCfg make_drop_cfg(float p, unsigned long long seed) {
    DropCfg d;
    if (p > 0.f) {
        d.p = p;
        d.scale = 1.0f / (1.0f - p);
        d.thr = (uint32_t)(p * 65536.0f + 0.5f);
        d.k0 = (uint32_t)seed;
        d.k1 = (uint32_t)(seed >> 32);
    }
    return d;
}

}  // namespace tante
