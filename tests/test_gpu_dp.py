"""Data-parallel parity on real GPUs (needs >= 2 devices: `gpurun --gpus 2 -- pytest tests/test_gpu_dp.py -m gpu`).
R ranks x 1 frame with allreduce(mean) must equal the single-process oracle run with nAveGrad = R
(reference train_parent.py:163-172), parent objective."""
import os

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import osvos_oracle as oc

pytestmark = pytest.mark.gpu
H, W = 64, 96


def _worker(rank, world, port, tmp):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world),
                      LOCAL_RANK=str(rank))
    from osvos_pytorch_b200 import parallel, training
    from osvos_pytorch_b200.networks.vgg_osvos import OSVOS
    parallel.init_distributed("nccl")
    dev = torch.device("cuda", rank)
    net = OSVOS(pretrained=0, verbose=False)
    net.load_state_dict(oc.he_params(seed=0), strict=False)
    net.to(dev).train()
    bucket = parallel.GradientBucket(parallel.trainable_parameters(net), dev)
    assert bucket.numel == 14917637
    x, gt = oc.synthetic_frame(1, H, W, 500 + rank)
    outs = net(x.to(dev))
    losses = [training.class_balanced_cross_entropy_loss(o, gt.to(dev), size_average=False) for o in outs]
    (0.5 * sum(losses[:-1]) + losses[-1]).backward()
    bucket.allreduce_mean()
    torch.cuda.synchronize()
    if rank == 0:
        torch.save({n: p.grad.detach().cpu() for n, p in net.named_parameters() if not n.startswith("upscale")},
                   os.path.join(tmp, "dp.pt"))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.skipif(not torch.cuda.is_available() or torch.cuda.device_count() < 2, reason="needs >= 2 GPUs")
def test_dp_allreduce_matches_reference_accumulation(tmp_path):
    world = 2
    mp.spawn(_worker, args=(world, 29600 + os.getpid() % 300, str(tmp_path)), nprocs=world, join=True)
    got = torch.load(tmp_path / "dp.pt")
    params = oc.he_params(seed=0)
    acc = None
    for r in range(world):
        x, gt = oc.synthetic_frame(1, H, W, 500 + r)
        _, _, g = oc.forward_backward(params, x, gt, objective="parent", side_weight=0.5, grad_scale=1.0 / world)
        acc = g if acc is None else {k: acc[k] + g[k] for k in g}
    worst = 0.0
    for k, v in acc.items():
        err = float((got[k].double() - v.double()).norm() / v.double().norm())
        worst = max(worst, err)
        assert err < (1e-3 if k.startswith(("fuse", "score_dsn", "side_prep")) else 4e-2), (k, err)   # tiny-map flip bound, see test_gpu_backward.py
    print(f"DP x{world}: worst per-parameter gradient error vs single-process nAveGrad={world} oracle: {worst:.2e}")


def _worker_default_init(rank, world, port, tmp):
    """The reference's own (unseeded) initialisation on every rank, made ONE model by parallel.broadcast_parameters; one
    optimizer step of the package's parent loop; every parameter must still be identical on all ranks afterwards."""
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world),
                      LOCAL_RANK=str(rank))
    from osvos_pytorch_b200 import parallel, training
    from osvos_pytorch_b200.networks.vgg_osvos import OSVOS
    parallel.init_distributed("nccl")
    dev = torch.device("cuda", rank)
    torch.manual_seed(1000 + rank)                       # different seeds: the replicas START different
    net = OSVOS(pretrained=0, verbose=False)
    with torch.no_grad():                                # logits of a usable size (the stock N(0, 1e-3) init gives ~1e-12)
        for p in net.parameters():
            if p.dim() == 4 and p.shape[2] in (1, 3):     # the 3x3 / 1x1 convs, not the fixed deconvolution taps
                p.mul_(30.0)
    net.to(dev)
    before = net.side_prep[0].weight.detach().clone()
    parallel.broadcast_parameters(net, src=0)
    if rank != 0:
        assert not torch.equal(before, net.side_prep[0].weight.detach())      # it really was a different model
    opt = training.make_optimizer(net, "parent", lr=1e-9, fused=True)
    bucket = parallel.GradientBucket(parallel.trainable_parameters(net), dev)
    batch = training.synthetic_batch(1, H, W, 700 + rank, dev)               # different frames per rank
    training.parent_epoch(net, opt, bucket, [batch], 0, 240, 1)
    torch.cuda.synchronize()
    flat = torch.cat([p.detach().flatten() for n, p in net.named_parameters() if not n.startswith("upscale")])
    gathered = [torch.empty_like(flat) for _ in range(world)]
    dist.all_gather(gathered, flat)
    if rank == 0:
        torch.save({"equal": all(torch.equal(gathered[0], g) for g in gathered[1:])}, os.path.join(tmp, "sync.pt"))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.skipif(not torch.cuda.is_available() or torch.cuda.device_count() < 2, reason="needs >= 2 GPUs")
def test_dp_replicas_are_one_model_after_broadcast_and_a_step(tmp_path):
    world = 2
    mp.spawn(_worker_default_init, args=(world, 29650 + os.getpid() % 300, str(tmp_path)), nprocs=world, join=True)
    assert torch.load(tmp_path / "sync.pt")["equal"]
