"""Stand-alone timing of the fused block-tail kernel vs the three GEMM launches it replaces (tante_test_gemm)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from tante_b200 import _abi

lib = _abi.load()
C = 256
st = torch.cuda.current_stream().cuda_stream
for M in (65536, 262144):
    att = torch.randn(M, C, device="cuda").bfloat16()
    Wo, W1, W2 = [(torch.randn(C, C, device="cuda") / 16).bfloat16() for _ in range(3)]
    v = torch.randn(7, C, device="cuda") * 0.1
    x = torch.randn(M, C, device="cuda")
    xo = torch.empty_like(x)
    ln = torch.empty(M, C, device="cuda", dtype=torch.bfloat16)
    hid = torch.empty(M, C, device="cuda", dtype=torch.bfloat16)
    xm = torch.empty_like(x); l2 = torch.empty_like(ln); hp = torch.empty_like(ln); ha = torch.empty_like(ln)
    flush = torch.empty(256 << 20, device="cuda", dtype=torch.uint8)

    def timeit(fn, iters=10):
        fn(); torch.cuda.synchronize()
        tot = 0.0
        for _ in range(iters):
            flush.zero_()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); fn(); e1.record(); torch.cuda.synchronize()
            tot += e0.elapsed_time(e1)
        return 1e3 * tot / iters

    def fused(train):
        def f():
            _abi.check(lib.tante_test_block_tail(att.data_ptr(), Wo.data_ptr(), W1.data_ptr(), W2.data_ptr(), v.data_ptr(),
                                                 x.data_ptr(), xo.data_ptr(), ln.data_ptr(), xm.data_ptr() if train else None,
                                                 l2.data_ptr() if train else None, hp.data_ptr() if train else None,
                                                 ha.data_ptr() if train else None, M, 1, st))
        return f

    def three():
        _abi.check(lib.tante_test_gemm(1, 6, att.data_ptr(), Wo.data_ptr(), v[0].data_ptr(), x.data_ptr(), xo.data_ptr(), 0, M, C, C, 1,
                                       v[1].data_ptr(), v[2].data_ptr(), ln.data_ptr(), st))
        _abi.check(lib.tante_test_gemm(1, 3, ln.data_ptr(), W1.data_ptr(), v[3].data_ptr(), None, hid.data_ptr(), 1, M, C, C, 1, None, None,
                                       None, st))
        _abi.check(lib.tante_test_gemm(1, 6, hid.data_ptr(), W2.data_ptr(), v[4].data_ptr(), xo.data_ptr(), xo.data_ptr(), 0, M, C, C, 1,
                                       v[5].data_ptr(), v[6].data_ptr(), ln.data_ptr(), st))
    t3, tf, tt = timeit(three), timeit(fused(False)), timeit(fused(True))
    by = M * (512 + 1024 + 1024 + 512)
    print(f"M={M}: three GEMM launches {t3:.1f} us | fused tail {tf:.1f} us ({by / tf / 1e3:.0f} GB/s algorithmic, "
          f"{3 * 2 * M * C * C / tf / 1e6:.0f} TFLOP/s) | fused tail (training stores) {tt:.1f} us "
          f"({(by + M * 2560) / tt / 1e3:.0f} GB/s)", flush=True)
