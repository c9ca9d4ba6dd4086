"""Per-phase clock64 trace of CTA 0 of the two-stream block-tail kernel (TANTE_TAIL_TRACE; debugging aid)."""
import os, struct, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
out = sys.argv[1] if len(sys.argv) > 1 else "gpurun_out/tail_trace.bin"
os.environ["TANTE_TAIL_TRACE"] = out
import torch
from tante_b200 import _abi
lib = _abi.load()
C, M = 256, int(os.environ.get("M", 262144))
st = torch.cuda.current_stream().cuda_stream
att = torch.randn(M, C, device="cuda").bfloat16()
Wo, W1, W2 = [(torch.randn(C, C, device="cuda") / 16).bfloat16() for _ in range(3)]
v = torch.randn(7, C, device="cuda") * 0.1
x = torch.randn(M, C, device="cuda"); xo = torch.empty_like(x)
ln = torch.empty(M, C, device="cuda", dtype=torch.bfloat16)
flush = torch.empty(256 << 20, device="cuda", dtype=torch.uint8)
for _ in range(3):
    flush.zero_()
    _abi.check(lib.tante_test_block_tail(att.data_ptr(), Wo.data_ptr(), W1.data_ptr(), W2.data_ptr(), v.data_ptr(), x.data_ptr(),
                                         xo.data_ptr(), ln.data_ptr(), None, None, None, None, M, 1, st))
    torch.cuda.synchronize()
raw = open(out, "rb").read()
vals = struct.unpack(f"{len(raw)//8}q", raw)
t0 = min(vals[(r * 1024) * 2 + 1] for r in range(3) if vals[(r * 1024) * 2 + 1] > 0)
for r, name in enumerate(("mma", "epi s0", "epi s1")):
    ev = [(vals[(r * 1024 + i) * 2], vals[(r * 1024 + i) * 2 + 1] - t0) for i in range(1024) if vals[(r * 1024 + i) * 2 + 1] > 0]
    print(name, len(ev), "events")
    prev = 0
    for e, t in ev[:int(os.environ.get("NEV", 70))]:
        print(f"   ev {e:4d}  t {t:8d}  (+{t - prev})")
        prev = t
