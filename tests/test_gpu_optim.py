"""SURVEY.md 8(f) item 3 - fused SGD step + weight repack, against torch.optim.SGD (the reference's own optimizer,
train_online.py:79-88) and the oracle's fp64 restatement."""
import pytest
import torch

from oracle import osvos_oracle as oc

pytestmark = pytest.mark.gpu


def _net(seed=0):
    from osvos_pytorch_b200.networks.vgg_osvos import OSVOS, he_init_
    return he_init_(OSVOS(pretrained=0, verbose=False), seed=seed).cuda().train()


def _fill_grads(net, seed, skip_score=False):
    g = torch.Generator().manual_seed(seed)
    for n, p in net.named_parameters():
        if n.startswith("upscale") or (skip_score and n.startswith("score_dsn")):
            p.grad = None
            continue
        p.grad = (torch.randn(p.shape, generator=g) * 10.0).cuda()


@pytest.mark.parametrize("mode", ["online", "parent"])
def test_fused_sgd_matches_torch_sgd(mode):
    from osvos_pytorch_b200 import training
    a, b = _net(), _net()
    # large lr / wd so that every term of the update matters in fp32
    oa = training.make_optimizer(a, mode, lr=1e-3, wd=0.05, fused=False)
    ob = training.make_optimizer(b, mode, lr=1e-3, wd=0.05, fused=True)
    for step in range(4):
        _fill_grads(a, 10 + step, skip_score=(mode == "online"))
        _fill_grads(b, 10 + step, skip_score=(mode == "online"))
        oa.step()
        ob.step()
    torch.cuda.synchronize()
    for (n, pa), (_, pb) in zip(a.named_parameters(), b.named_parameters()):
        err = float((pa - pb).abs().max() / pa.abs().max().clamp_min(1e-30))
        assert err < 2e-6, (n, err)
        if n.startswith("upscale") or (mode == "online" and n.startswith("score_dsn")):
            assert torch.equal(pa, pb)                                     # untouched
    # momentum buffers interchangeable with torch.optim.SGD's state
    for pa, pb in zip(a.parameters(), b.parameters()):
        sa, sb = oa.state.get(pa, {}), ob.state.get(pb, {})
        if sa.get("momentum_buffer") is not None:
            ma, mb = sa["momentum_buffer"], sb["momentum_buffer"]
            assert float((ma - mb).abs().max() / ma.abs().max().clamp_min(1e-30)) < 2e-6


def test_fused_sgd_against_fp64_oracle_and_zero_grad():
    from osvos_pytorch_b200.optim import FusedSGD
    g = torch.Generator().manual_seed(3)
    shapes = [(5,), (1,), (4609,), (64, 3, 3, 3), (2, 4608), (9221,)]     # odd sizes: scalar tails, unaligned views
    flat = torch.randn(sum(torch.Size(s).numel() for s in shapes) + 3, generator=g).cuda()
    params, off = [], 3                                                     # offset 3 floats: 12-byte aligned views
    for s in shapes:
        n = torch.Size(s).numel()
        params.append(torch.nn.Parameter(flat[off:off + n].view(s).clone()))
        off += n
    gflat = torch.randn(off, generator=g).cuda()
    off = 3
    for p in params:
        p.grad = gflat[off:off + p.numel()].view(p.shape)                   # deliberately unaligned gradient views
        off += p.numel()
    opt = FusedSGD([{"params": params[:3], "lr": 0.1, "weight_decay": 0.01},
                    {"params": params[3:], "lr": 0.02}], lr=0.5, momentum=0.9)
    ref_p = [p.detach().clone() for p in params]
    ref_b = [None] * len(params)
    for step in range(3):
        grads = [p.grad.clone() for p in params]
        opt.step()
        for i in range(len(params)):
            lr, wd = (0.1, 0.01) if i < 3 else (0.02, 0.0)
            ref_p[i], ref_b[i] = oc.sgd_momentum_step(ref_p[i], grads[i], ref_b[i], lr, wd, 0.9)
        for p, r in zip(params, ref_p):
            assert float((p - r).abs().max()) < 1e-5 * float(r.abs().max())
    v = params[0]._version
    opt.step(zero_grad=True)
    assert params[0]._version > v
    assert all(float(p.grad.abs().max()) == 0.0 for p in params)


def test_fused_sgd_reemits_packed_layouts_bit_exactly():
    from osvos_pytorch_b200 import ops, training
    net = _net()
    opt = training.make_optimizer(net, "parent", lr=1e-3, wd=0.05, fused=True)
    _fill_grads(net, 1)
    table_before = [(w, f.clone(), t.clone()) for w, f, t in net._engine.packed_weight_table()]
    opt.step()
    opt.step()
    for (w, f_old, t_old), (w2, f, t) in zip(table_before, net._engine.packed_weight_table()):
        assert w is w2
        assert f.data_ptr() != f_old.data_ptr() or True
        want_f = ops.pack_conv3x3_weights(w, False, 64)
        want_t = ops.pack_conv3x3_weights(w, True, 64)
        assert torch.equal(f.view(torch.int16), want_f.view(torch.int16)), tuple(w.shape)
        assert torch.equal(t.view(torch.int16), want_t.view(torch.int16)), tuple(w.shape)
        assert not torch.equal(f.view(torch.int16), f_old.view(torch.int16))            # it did change
    # the engine serves the re-emitted buffers without repacking (same storage, current version stamp)
    ptrs = [f.data_ptr() for _, f, _ in net._engine.packed_weight_table()]
    opt.step()
    assert ptrs == [f.data_ptr() for _, f, _ in net._engine.packed_weight_table()]
    # and a forward after the fused step equals a forward of a fresh module holding the same weights
    from osvos_pytorch_b200.networks.vgg_osvos import OSVOS
    fresh = OSVOS(pretrained=0, verbose=False).cuda().eval()
    fresh.load_state_dict(net.state_dict())
    x = oc.synthetic_frame(1, 40, 56, 3)[0].cuda()
    with torch.no_grad():
        a, b = net.eval()(x), fresh(x)
    for u, v in zip(a, b):
        assert torch.equal(u, v)


def test_online_finetune_fused_equals_unfused():
    from osvos_pytorch_b200 import training
    x, gt = oc.synthetic_frame(1, 40, 56, 11)
    sample = {"image": x.cuda(), "gt": gt.cuda()}
    hist = {}
    for fused in (False, True):
        net = _net(seed=1)
        with torch.no_grad():
            for m in list(net.side_prep) + [net.fuse]:
                m.weight.mul_(0.1)
        hist[fused] = training.online_finetune(net, lambda it: sample, 20, n_ave_grad=5, lr=1e-9, log_every=5,
                                               log=lambda s: None, fused_optimizer=fused)
    for a, b in zip(hist[False], hist[True]):
        assert abs(a - b) <= 3e-4 * abs(a), (hist[False], hist[True])     # wgrad atomics are order-nondeterministic
