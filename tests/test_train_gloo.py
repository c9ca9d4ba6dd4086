"""Host-side training logic on CPU (no CUDA extension involved): the loss mirror against the oracle's restatement
of trainer/metrics.py, and the data-parallel gradient path (ONE flat bucket, all-reduce before clipping) with
world_size-2 gloo processes -- the averaged gradient of the two half-batches must equal the full-batch gradient."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import tante_oracle as O
from tante_b200.trainer import GradBucket, mse_loss


def test_mse_loss_matches_oracle_restatement():
    g = torch.Generator().manual_seed(0)
    y = torch.randn(2, 4, 8, 12, 3, generator=g)
    r = torch.randn(2, 4, 8, 12, 3, generator=g)
    assert torch.allclose(mse_loss(y, r), O.train_loss(y, r, None))
    for rts in (torch.tensor([1.01, 1.2, 1.4]), torch.tensor([4.5, 5.0]), torch.tensor([2.0, 3.0])):
        assert torch.allclose(mse_loss(y, r, rts, 0.5, 2), O.train_loss(y, r, rts, 0.5, 2))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _model():
    torch.manual_seed(3)
    return torch.nn.Sequential(torch.nn.Linear(6, 16), torch.nn.GELU(), torch.nn.Linear(16, 4))


def _worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    model = _model()                                   # identical seeds -> identical replicas
    bucket = GradBucket(model)
    g = torch.Generator().manual_seed(7)
    x = torch.randn(8, 6, generator=g)
    y = torch.randn(8, 4, generator=g)
    lo, hi = rank * 4, rank * 4 + 4                    # contiguous half of the global batch
    for _ in range(2):                                 # second pass: zero() must reset the shared storage
        bucket.zero()
        torch.nn.functional.mse_loss(model(x[lo:hi]), y[lo:hi]).backward()
        assert all(p.grad.data_ptr() >= bucket.flat.data_ptr() for p in bucket.params)   # still views of the bucket
        bucket.all_reduce_mean()
    q.put((rank, bucket.flat.clone()))
    dist.barrier()
    dist.destroy_process_group()


def test_grad_bucket_allreduce_equals_full_batch_gradient():
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    got = dict(q.get(timeout=120) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    # reference: the full batch on one process
    model = _model()
    g = torch.Generator().manual_seed(7)
    x = torch.randn(8, 6, generator=g)
    y = torch.randn(8, 4, generator=g)
    torch.nn.functional.mse_loss(model(x), y).backward()
    ref = torch.cat([p.grad.reshape(-1) for p in model.parameters()])
    assert torch.allclose(got[0], got[1])
    assert torch.allclose(got[0], ref, atol=1e-6)
