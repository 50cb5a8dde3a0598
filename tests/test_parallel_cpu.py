"""Host-side logic of the data-parallel path on CPU: world_size 2 over gloo.  (The CUDA model has no CPU
fallback, so a small stand-in module with the same parameter plumbing is used; the GPU DP test lives in
tests/test_gpu_dp.py.)  Checks: flat-bucket views, allreduce(mean) == single-process nAveGrad accumulation
(the reference's train_parent.py:163-172 semantics), frame sharding."""
import os

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp
import torch.nn as nn

from osvos_pytorch_b200.parallel import GradientBucket, shard_range, trainable_parameters


class Tiny(nn.Module):
    def __init__(self):
        super().__init__()
        self.upscale = nn.ModuleList([nn.ConvTranspose2d(1, 1, 4, stride=2, bias=False)])
        self.stages = nn.ModuleList([nn.Sequential(nn.Conv2d(3, 4, 3, padding=1), nn.ReLU())])
        self.fuse = nn.Conv2d(4, 1, 1)

    def forward(self, x):
        return self.fuse(self.stages[0](x))


def _worker(rank, world, port, tmp):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.manual_seed(0)
    net = Tiny()
    bucket = GradientBucket(trainable_parameters(net))
    assert bucket.numel == sum(p.numel() for n, p in net.named_parameters() if not n.startswith("upscale"))
    g = torch.Generator().manual_seed(100)
    xs = torch.randn(world, 2, 3, 8, 8, generator=g)
    lo, hi = shard_range(world, rank, world)
    for i in range(lo, hi):
        net(xs[i]).pow(2).sum().backward()
    for p in bucket.params:                                    # grads accumulated INTO the bucket
        assert p.grad.untyped_storage().data_ptr() == bucket.flat.untyped_storage().data_ptr()
    bucket.allreduce_mean()
    torch.save(bucket.flat.clone(), os.path.join(tmp, f"r{rank}.pt"))
    if rank == 0:                                              # single-process oracle: nAveGrad = world
        ref = Tiny()
        ref.load_state_dict(net.state_dict())
        for i in range(world):
            (ref(xs[i]).pow(2).sum() / world).backward()
        flat = torch.cat([p.grad.flatten() for p in trainable_parameters(ref)])
        torch.save(flat, os.path.join(tmp, "ref.pt"))
    bucket.zero_()
    assert float(bucket.flat.abs().max()) == 0.0 and all(float(p.grad.abs().max()) == 0.0 for p in bucket.params)
    dist.destroy_process_group()


def test_allreduce_mean_equals_gradient_accumulation(tmp_path):
    world, port = 2, 29500 + os.getpid() % 400
    mp.spawn(_worker, args=(world, port, str(tmp_path)), nprocs=world, join=True)
    r0, r1, ref = (torch.load(tmp_path / f) for f in ("r0.pt", "r1.pt", "ref.pt"))
    assert torch.equal(r0, r1)
    assert torch.allclose(r0, ref, rtol=1e-5, atol=1e-6)


def test_shard_range_partitions():
    for total in (1, 7, 12, 13, 2079):
        for world in (1, 2, 4, 8):
            spans = [shard_range(total, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == total
            assert all(spans[i][1] == spans[i + 1][0] for i in range(world - 1))
            sizes = [hi - lo for lo, hi in spans]
            assert max(sizes) - min(sizes) <= 1


def test_optimizer_groups_match_reference_recipe():
    from osvos_pytorch_b200.networks.vgg_osvos import OSVOS
    from osvos_pytorch_b200.training import make_optimizer
    net = OSVOS(pretrained=0, verbose=False)
    on = make_optimizer(net, "online")
    pa = make_optimizer(net, "parent")
    assert len(on.param_groups) == 8 and len(pa.param_groups) == 10          # train_online.py:79-88 / train_parent.py:87-103
    lr = 1e-8
    assert [g["lr"] for g in on.param_groups] == [lr, 2 * lr, lr, 2 * lr, 0, 0, lr / 100, 2 * lr / 100]
    assert [g["lr"] for g in pa.param_groups] == [lr, 2 * lr, lr, 2 * lr, lr / 10, 2 * lr / 10, 0, 0, lr / 100, 2 * lr / 100]
    assert [g["weight_decay"] for g in pa.param_groups] == [2e-4, 0, 2e-4, 0, 2e-4, 0, 0, 0, 2e-4, 0]
    online_ids = {id(p) for g in on.param_groups for p in g["params"]}
    assert all(id(p) not in online_ids for p in net.score_dsn.parameters())    # score_dsn not optimised online
