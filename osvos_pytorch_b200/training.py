"""Loop bodies of the two entry points, factored so that train_online.py / train_parent.py stay thin:
optimizer construction with the reference's per-group learning rates, synthetic DAVIS-shaped data,
the online fine-tune loop (train_online.py:112-149 of the reference) and the parent loop with the new
data-parallel exchange step (train_parent.py:129-176 + parallel.py)."""
import torch

from .layers.osvos_layers import class_balanced_cross_entropy_loss

MEANVAL = (104.00699, 116.66877, 122.67892)      # dataloaders/davis_2016.py:19 of the reference
ONLINE_WEIGHTS = (0.0, 0.0, 0.0, 0.0, 1.0)       # train_online.py:127: only the fused map is supervised


def _named(module, key):
    return [p for n, p in module.named_parameters() if key in n]


def make_optimizer(net, mode, lr=1e-8, wd=0.0002, momentum=0.9, fused=False):
    """SGD with the reference's parameter groups (``fused``: optim.FusedSGD - one launch per step, packed conv
    layouts re-emitted in the same pass - instead of torch.optim.SGD).
    online (train_online.py:77-88): stages / side_prep weights (wd) and biases (2 lr), deconvs lr 0,
    fuse at lr/100; score_dsn is NOT optimised.  parent (train_parent.py:85-103): additionally score_dsn at lr/10."""
    groups = [
        {"params": _named(net.stages, "weight"), "weight_decay": wd, "initial_lr": lr},
        {"params": _named(net.stages, "bias"), "lr": 2 * lr, "initial_lr": 2 * lr},
        {"params": _named(net.side_prep, "weight"), "weight_decay": wd, "initial_lr": lr},
        {"params": _named(net.side_prep, "bias"), "lr": 2 * lr, "initial_lr": 2 * lr},
    ]
    if mode == "parent":
        groups += [
            {"params": _named(net.score_dsn, "weight"), "lr": lr / 10, "weight_decay": wd, "initial_lr": lr / 10},
            {"params": _named(net.score_dsn, "bias"), "lr": 2 * lr / 10, "initial_lr": 2 * lr / 10},
        ]
    groups += [
        {"params": _named(net.upscale, "weight"), "lr": 0, "initial_lr": 0},
        {"params": _named(net.upscale_, "weight"), "lr": 0, "initial_lr": 0},
        {"params": [net.fuse.weight], "lr": lr / 100, "initial_lr": lr / 100, "weight_decay": wd},
        {"params": [net.fuse.bias], "lr": 2 * lr / 100, "initial_lr": 2 * lr / 100},
    ]
    if fused:
        from .optim import FusedSGD
        return FusedSGD(groups, lr=lr, momentum=momentum, engine=net._engine)
    return torch.optim.SGD(groups, lr=lr, momentum=momentum)


def synthetic_batch(n, h, w, seed, device):
    """DAVIS-shaped synthetic sample: BGR 0..255 mean-subtracted image, ~30 % positive mask."""
    g = torch.Generator().manual_seed(seed)
    img = torch.rand(n, 3, h, w, generator=g) * 255.0 - torch.tensor(MEANVAL).view(1, 3, 1, 1)
    gt = (torch.rand(n, 1, h, w, generator=g) > 0.7).float()
    return {"image": img.to(device), "gt": gt.to(device)}


class GraphedTrainStep:
    """Forward + objective + backward of one fixed-shape micro-batch captured in a CUDA graph (SURVEY.md 8f item 1:
    at > 250 fwd+bwd/s the ~120 kernel launches and the Python around them cost as much as the kernels).

    Gradients ACCUMULATE into the static ``p.grad`` buffers exactly like repeated ``loss.backward()`` calls
    (train_online.py:140-149): call ``zero_grads()`` after ``optimizer.step()`` (``optimizer.zero_grad()`` with
    set_to_none would detach the graph from its buffers).  The weight-packing kernels are part of the graph, so
    parameter updates between replays are picked up - unless ``external_pack`` is set: then the packed conv
    layouts are static buffers outside the graph that ``optim.FusedSGD`` rewrites in its update kernel (no
    packing work per micro-batch; only valid with that optimizer).  ``objective(outputs, gts) -> scalar tensor``, or a
    tuple of five loss weights = the package's fused objective (``OSVOS.forward_objective``).
    """

    def __init__(self, net, objective, sample, grad_scale=1.0, external_pack=False):
        self.net, self.objective, self.grad_scale = net, objective, float(grad_scale)
        self.x = sample["image"].detach().clone()
        self.gt = sample["gt"].detach().clone()
        self.params = [p for n, p in net.named_parameters() if not n.startswith("upscale")]
        for p in self.params:
            if p.grad is None:
                p.grad = torch.zeros_like(p)
        keep = [p.grad.clone() for p in self.params]
        cur = torch.cuda.current_stream()
        side = torch.cuda.Stream()
        side.wait_stream(cur)
        with torch.cuda.stream(side):
            for _ in range(2):                              # warm-up: lazy kernel attributes, allocator, autograd
                self._body()
        cur.wait_stream(side)
        for p, g in zip(self.params, keep):                 # undo the warm-up accumulation
            p.grad.copy_(g)
        if external_pack:
            net._engine.packed_weight_table()               # packed now, outside the graph
        # everything else derived from parameters is recomputed inside the graph
        net._engine.drop_derived_caches(keep_packed=external_pack)
        torch.cuda.synchronize()
        self.graph = torch.cuda.CUDAGraph()
        try:
            with torch.cuda.graph(self.graph):
                self.loss = self._body()
        finally:
            # graph-private buffers must not serve eager calls (also after a failed capture: the cached packed layouts
            # would point into the aborted capture's memory pool)
            net._engine.drop_derived_caches(keep_packed=external_pack)

    def _body(self):
        if not callable(self.objective):
            # loss weights of the package's fused objective (tail + five losses = one kernel each way); grad_scale is
            # folded into the weights, so no scaling kernel runs either
            w = [float(v) * self.grad_scale for v in self.objective]
            _, total, per_map = self.net.forward_objective(self.x, self.gt, w)
            with self.net._engine.direct_grad_accumulation():
                total.backward()
            self.per_map = per_map
            nz = [k for k, v in enumerate(self.objective) if float(v) != 0.0]
            if len(nz) == 1 and float(self.objective[nz[0]]) == 1.0:
                return per_map[nz[0]]                       # the unscaled objective IS that map's loss: no extra kernel
            return total.detach() if self.grad_scale == 1.0 else total.detach() / self.grad_scale
        outputs = self.net(self.x)
        loss = self.objective(outputs, self.gt)
        with self.net._engine.direct_grad_accumulation():   # p.grad buffers are static: add into them in the kernels
            (loss * self.grad_scale).backward()
        return loss.detach()

    def __call__(self, sample=None):
        if sample is not None:
            self.x.copy_(sample["image"], non_blocking=True)
            self.gt.copy_(sample["gt"], non_blocking=True)
        self.graph.replay()
        return self.loss

    def zero_grads(self, skip=()):
        """Zero the accumulated gradients (``skip``: parameters already cleared, e.g. by FusedSGD.step(zero_grad=True))."""
        skip = {id(p) for p in skip}
        for p in self.params:
            if id(p) not in skip:
                p.grad.zero_()


def online_finetune(net, sample_fn, iters, n_ave_grad=5, lr=1e-8, wd=0.0002, log_every=0, log=print, use_graph=True,
                    fused_optimizer=True):
    """`iters` forward/backward passes on the annotated frame, SGD step every `n_ave_grad` (reference
    train_online.py:112-149).  Losses are kept on the device; one host read per `log_every` iterations
    instead of the reference's per-iteration .item() sync.  With `use_graph` the fwd+loss+bwd of a micro-batch
    is a replayed CUDA graph (shapes must not change between iterations); with `fused_optimizer` the SGD step,
    the gradient zeroing and the repack of the conv weights are one kernel (optim.FusedSGD).  Returns the list of
    logged losses."""
    net.train()
    opt = make_optimizer(net, "online", lr, wd, fused=fused_optimizer)
    opt.zero_grad()
    opt_params = [p for g in opt.param_groups for p in g["params"]]
    history, running = [], None
    step = None
    for it in range(iters):
        sample = sample_fn(it)
        inputs, gts = sample["image"], sample["gt"]
        if use_graph:
            if step is None:
                step = GraphedTrainStep(net, ONLINE_WEIGHTS, sample, grad_scale=1.0 / n_ave_grad,
                                        external_pack=fused_optimizer)
            loss_val = step(sample)
            running = loss_val.clone() if running is None else running + loss_val
            if (it + 1) % n_ave_grad == 0:
                if fused_optimizer:
                    opt.step(zero_grad=True)
                    step.zero_grads(skip=opt_params)
                else:
                    opt.step()
                    step.zero_grads()
        else:
            # fuse-map loss only (train_online.py:127), 1/nAveGrad folded into the weight (train_online.py:140)
            _, loss, per_map = net.forward_objective(inputs, gts, [v / n_ave_grad for v in ONLINE_WEIGHTS])
            running = per_map[4].clone() if running is None else running + per_map[4]
            with net._engine.direct_grad_accumulation():
                loss.backward()
            if (it + 1) % n_ave_grad == 0:
                if fused_optimizer:
                    opt.step(zero_grad=True)
                else:
                    opt.step()
                    opt.zero_grad()
        if log_every and (it + 1) % log_every == 0:
            val = float(running) / log_every
            history.append(val)
            running = None
            log(f"[iter {it + 1:6d}] loss {val:.6f}")
    return history


def parent_epoch(net, opt, bucket, batches, epoch, n_epochs, n_ave_grad=1, group=None, state=None):
    """One epoch of the parent objective on this rank's shard: deep-supervision loss
    (1 - epoch/nEpochs) * sum_{k<4} L_k + L_fuse (train_parent.py:143-147), gradient accumulation over
    `n_ave_grad` local micro-batches, then ONE allreduce(mean) and one SGD step.
    Returns the mean of the five losses over the micro-batches as a device tensor (the reference reads `.item()` of
    every loss in every iteration, train_parent.py:146: a host sync per step that exposes launch latency and, in data
    parallel, lets rank skew accumulate; here the host only waits when it logs).
    `state`: a dict the caller keeps across epochs; it carries the accumulation counter, which the reference does NOT
    reset at epoch boundaries (`aveGrad`, train_parent.py:125,165-172) - leftover micro-batches of an epoch whose length
    is not a multiple of nAveGrad complete their group in the next epoch instead of inflating its first step."""
    net.train()
    if state is None:
        state = {}
    state.setdefault("ave_grad", 0)
    side_w = 1.0 - epoch / n_epochs
    totals = torch.zeros(5, device=bucket.flat.device)
    count = 0
    for it, sample in enumerate(batches):
        # (1 - epoch/nEpochs) * sum(side losses) + fuse loss, / nAveGrad (train_parent.py:143-147,163), as the fused
        # objective: tail + five losses are one kernel forward and one backward
        w = [side_w / n_ave_grad] * 4 + [1.0 / n_ave_grad]
        _, loss, per_map = net.forward_objective(sample["image"], sample["gt"], w)
        totals += per_map
        count += 1
        with net._engine.direct_grad_accumulation():
            loss.backward()
        state["ave_grad"] += 1
        if state["ave_grad"] % n_ave_grad == 0:
            state["ave_grad"] = 0
            bucket.allreduce_mean(group)
            if hasattr(opt, "_engine"):                     # optim.FusedSGD: update + zeroing + repack in one kernel
                opt.step(zero_grad=True)
            else:
                opt.step()
                bucket.zero_()
    return totals / max(count, 1)          # DEVICE tensor [5]: no host sync per call; callers read it when they log


def timed_parent_steps(net, opt, bucket, make_batch, steps, warmup, epoch=0, n_epochs=240, group=None):
    """Benchmark helper: `steps` optimizer steps (1 micro-batch each); device time via CUDA events."""
    def one(i):
        parent_epoch(net, opt, bucket, [make_batch(i)], epoch, n_epochs, 1, group)
    for i in range(warmup):
        one(i)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(steps):
        one(warmup + i)
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / steps
