"""Time the stand-alone GEMM kernels through the C-ABI test hook (CUDA events, L2-sized-out inputs)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from tante_b200 import _abi

lib = _abi.load()
shapes = [(262144, 768, 256), (262144, 256, 256), (65536, 768, 256), (65536, 256, 256), (16384, 256, 256),
          (262144, 128, 256), (262144, 256, 512)]
if len(sys.argv) > 1:
    shapes = [tuple(int(v) for v in sys.argv[1].split(","))]
iters = int(sys.argv[2]) if len(sys.argv) > 2 else 10
for (M, N, K) in shapes:
    for epi, out_bf16 in ((0, 1), (3, 1), (4, 0), (6, 0)):
        if epi in (4, 6) and (N != 256 or K > 256):
            continue
        A = torch.randn(M, K, device="cuda").to(torch.bfloat16)
        W = torch.randn(N, K, device="cuda").to(torch.bfloat16)
        bias = torch.randn(N, device="cuda")
        C = torch.empty(M, N, device="cuda", dtype=torch.bfloat16 if out_bf16 else torch.float32)
        st = torch.cuda.current_stream().cuda_stream
        args = (1, epi, A.data_ptr(), W.data_ptr(), bias.data_ptr(), C.data_ptr() if epi in (4, 6) else None, C.data_ptr(),
                out_bf16, M, N, K)
        gam = torch.ones(N, device='cuda'); bet = torch.zeros(N, device='cuda'); lno = torch.empty(M, N, device='cuda', dtype=torch.bfloat16)
        tail = (gam.data_ptr(), bet.data_ptr(), lno.data_ptr(), st) if epi == 6 else (None, None, None, st)
        _abi.check(lib.tante_test_gemm(*args, 3, *tail))
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        _abi.check(lib.tante_test_gemm(*args, iters, *tail))
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / iters
        print(f"M={M} N={N} K={K} epi={epi} out_bf16={out_bf16}: {ms*1e3:.1f} us  {2*M*N*K/ms/1e9:.1f} TFLOP/s", flush=True)
