"""Test-side restatement of the library's counter-based dropout masks (tante_b200/csrc/dropout.cuh): Philox4x32 with 7
rounds, key = the per-call 64-bit seed, counter = (element group, site); 8 x 16-bit lanes per call; keep iff lane >= p*65536."""
import numpy as np
import torch

M0, M1, W0, W1 = 0xD2511F53, 0xCD9E8D57, 0x9E3779B9, 0xBB67AE85
MASK32 = np.uint64(0xFFFFFFFF)


def philox4x32_7(c0, c1, c2, c3, k0, k1):
    c0, c1, c2, c3 = [np.asarray(c, dtype=np.uint64) for c in (c0, c1, c2, c3)]
    k0, k1 = np.uint64(k0), np.uint64(k1)
    for _ in range(7):
        p0 = np.uint64(M0) * c0
        p1 = np.uint64(M1) * c2
        hi0, lo0 = p0 >> np.uint64(32), p0 & MASK32
        hi1, lo1 = p1 >> np.uint64(32), p1 & MASK32
        c0, c1, c2, c3 = (hi1 ^ c1 ^ k0) & MASK32, lo1, (hi0 ^ c3 ^ k1) & MASK32, lo0
        k0 = (k0 + np.uint64(W0)) & MASK32
        k1 = (k1 + np.uint64(W1)) & MASK32
    return c0, c1, c2, c3


def multipliers(seed: int, site: int, grp: np.ndarray, lane: np.ndarray, p: float) -> np.ndarray:
    """0 or 1/(1-p) for elements given by (group, lane) arrays (uint64 / int)."""
    grp = np.asarray(grp, dtype=np.uint64)
    w = philox4x32_7(grp & MASK32, grp >> np.uint64(32), np.full(grp.shape, site, np.uint64),
                     np.full(grp.shape, 0x7A17E0D0, np.uint64), seed & 0xFFFFFFFF, (seed >> 32) & 0xFFFFFFFF)
    lane = np.asarray(lane)
    word = np.choose(lane >> 1, w)
    v = np.where(lane & 1, word >> np.uint64(16), word & np.uint64(0xFFFF))
    thr = int(p * 65536.0 + 0.5)
    return np.where(v >= thr, 1.0 / (1.0 - p), 0.0)


def site_id(order, layer, which):
    return 4 * (order * 64 + layer) + which


class DropMasks:
    """`drop_fn` for oracle.tante_oracle.backbone: explicit multipliers of every dropout site of one model call."""

    def __init__(self, seed: int, p: float, n_head: int, C: int, dtype=torch.float32):
        self.seed, self.p, self.n_head, self.C, self.dtype = seed, p, n_head, C, dtype

    def __call__(self, order, layer, tok):
        tok = tok.numpy().astype(np.uint64)                      # (N, S)
        N, S = tok.shape
        C, nh = self.C, self.n_head
        col = np.arange(C, dtype=np.uint64)
        e = tok[:, :, None] * np.uint64(C) + col[None, None, :]                  # residual sites: element = token * C + column
        res = [multipliers(self.seed, site_id(order, layer, w), e >> np.uint64(3), (e & np.uint64(7)).astype(np.int64), self.p)
               for w in (1, 2)]
        head = np.arange(nh, dtype=np.uint64)
        kpos = np.arange(S, dtype=np.uint64)
        base = (tok[:, None, :, None] * np.uint64(nh) + head[None, :, None, None]) << np.uint64(13)      # (N, heads, Sq, 1)
        grp = base | (kpos[None, None, None, :] >> np.uint64(3))
        lane = np.broadcast_to((kpos & np.uint64(7)).astype(np.int64)[None, None, None, :], grp.shape)
        attn = multipliers(self.seed, site_id(order, layer, 0), grp, lane, self.p)
        t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(self.dtype)
        return {"attn": t(attn), "res1": t(res[0]), "res2": t(res[1])}


def next_call_seed(torch_seed: int) -> int:
    """The key tante_b200.TANTE draws for the next taped call after torch.manual_seed(torch_seed)."""
    torch.manual_seed(torch_seed)
    return int(torch.randint(0, 2 ** 62, (1,)).item())
