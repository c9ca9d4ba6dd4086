import sys, os, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tante_b200 import _abi
lib = _abi.load()
M, N, K = 4096, 256, 256
g = torch.Generator(device="cuda").manual_seed(5)
A = torch.randn(M, N, device="cuda", generator=g).to(torch.bfloat16)
B = torch.randn(M, K, device="cuda", generator=g).to(torch.bfloat16)
C = torch.zeros(N, K, device="cuda"); bias = torch.zeros(N, device="cuda")
_abi.check(lib.tante_test_wgrad(1, A.data_ptr(), B.data_ptr(), C.data_ptr(), bias.data_ptr(), M, N, K, 1, torch.cuda.current_stream().cuda_stream))
torch.cuda.synchronize()
print("bias", bias[:6].tolist(), bias[128:131].tolist())
print("colsum", A.float().sum(0)[:6].tolist())
print("dW[:,0]", C[:6, 0].tolist(), "dW[:,1]", C[:3, 1].tolist())
print("nan count", int(torch.isnan(bias).sum()))
