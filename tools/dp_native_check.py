"""torchrun check of the native data-parallel exchange: tante_allreduce_grads (the library's own NCCL communicator) against
torch.distributed.all_reduce on the same bucket, then two FusedAdamW training steps on rank-dependent batches -- the
replicas must stay bit-identical.  Run: python -m torch.distributed.run --nproc-per-node 2 tools/dp_native_check.py"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist
from tante_b200 import TANTE, TanteMetadata, FusedAdamW
from tante_b200.trainer import GradBucket, train_step

rank, local, world = int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
torch.manual_seed(211)
model = TANTE(4, TanteMetadata(spatial_resolution=(64, 64), n_fields=3), taylor_order=1, attn_axes="THW", patch_scale=8,
              deg=True, precision="bf16").to(dev).train()
bucket = GradBucket(model)
assert bucket.init_native_comm(model), "native communicator was not created"
g = torch.Generator(device=dev).manual_seed(100 + rank)
bucket.flat.copy_(torch.randn(bucket.flat.numel(), device=dev, generator=g))
ref = bucket.flat.clone()
dist.all_reduce(ref)
assert bucket.all_reduce_sum() == world
torch.cuda.synchronize()
err = (bucket.flat - ref).abs().max().item()
print(f"rank {rank}: native all-reduce vs torch.distributed: max abs diff {err:.3e}", flush=True)
assert err <= 1e-5 * ref.abs().max().item()      # ring order may differ between two communicators
opt = FusedAdamW(model.parameters(), lr=1e-3, weight_decay=1e-5, model=model)
x = torch.randn(2, 4, 3, 64, 64, device=dev, generator=g)
y = torch.randn(2, 2, 64, 64, 3, device=dev, generator=g)
for _ in range(2):
    loss = train_step(model, opt, x, y, 2, bucket)
flat_p = torch.cat([p.detach().reshape(-1) for p in model.parameters()])
hi, lo = flat_p.clone(), flat_p.clone()
dist.all_reduce(hi, op=dist.ReduceOp.MAX)
dist.all_reduce(lo, op=dist.ReduceOp.MIN)
spread = (hi - lo).abs().max().item()
print(f"rank {rank}: loss {loss.item():.5f}, parameter spread across ranks after 2 steps {spread:.3e}", flush=True)
assert spread == 0.0
dist.barrier()
if rank == 0:
    print("dp_native_check OK")
dist.destroy_process_group()
