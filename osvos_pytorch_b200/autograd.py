"""Training path: one torch.autograd.Function spanning the whole network, so that autograd sees
a single node whose forward/backward are sequences of native kernels (no PyTorch op does
arithmetic).  Replaces the autograd graph of reference networks/vgg_osvos.py:59-74 built at
train_online.py:124 / train_parent.py:140 and walked at :141 / :164.

Gradient bookkeeping mirrors the reference: parameters that do not influence the objective get
``None`` (e.g. score_dsn.* under the fuse-only online loss); the fixed
bilinear deconvolution weights (lr = 0 in both scripts) receive no gradient.
"""
import torch
import torch.nn as nn

from . import ops


def _trunk_convs(m):
    return [[c for c in stage if isinstance(c, nn.Conv2d)] for stage in m.stages]


class _OSVOSFunction(torch.autograd.Function):
    @staticmethod
    def forward(ctx, engine, x, objective, *params):
        """objective: None (plain forward: the five maps, each differentiable) or (label, loss_weights[5], divisor):
        the package's own objective fused into the tail - outputs are then (5 maps, total, losses[5]) with only
        `total = sum_k w_k * class_balanced_cross_entropy_loss(map_k, label)` differentiable."""
        m = engine.m
        fast = m.precision == "fast"
        engine._check_deconvs()
        ctx.set_materialize_grads(False)
        xin = x.detach().contiguous().float()
        n, _, h, w = (int(v) for v in xin.shape)
        convs = _trunk_convs(m)
        acts = []      # acts[i][j] = output act of conv j of stage i
        pooled = [None]
        a = ops.conv_first(xin, convs[0][0].weight.detach(), convs[0][0].bias.detach(), relu=True, fast=fast)
        stage_acts = [a]
        full, a = ops.conv3x3(a, engine._packed(convs[0][1], "s0c1"), convs[0][1].bias.detach(), 64, relu=True,
                              fast=fast, pool=True)               # 2x2 max pool fused into the epilogue
        stage_acts.append(full)
        acts.append(stage_acts)
        for i in range(1, 5):
            pooled.append(a)
            stage_acts = []
            for j, conv in enumerate(convs[i]):
                if j == len(convs[i]) - 1 and i < 4:
                    full, a = ops.conv3x3(a, engine._packed(conv, f"s{i}c{j}"), conv.bias.detach(), conv.out_channels,
                                          relu=True, fast=fast, pool=True)
                else:
                    a, _, _ = ops.conv3x3(a, engine._packed(conv, f"s{i}c{j}"), conv.bias.detach(), conv.out_channels,
                                          relu=True, fast=fast)
                    full = a
                stage_acts.append(full)
            acts.append(stage_acts)
        # side_prep has no ReLU: side_prep o (score_dsn, fuse slice) is ONE 3x3 conv C -> 2 (csrc/side_conv.cu), all four
        # scales in one launch; its backward needs neither the 16 features nor their gradient (csrc/side_bwd_folded.cu)
        pqs = ops.side_folded_multi([acts[i][-1] for i in range(1, 5)], engine._folded_side_all(), fast=fast)
        ctx.engine = engine
        ctx.dims = (n, h, w)
        ctx.fast = fast
        if getattr(engine, "debug_capture", None) is not None:      # tests: the saved activations of this pass
            engine.debug_capture.update(acts=acts, pooled=pooled)
        if objective is None:
            out, _ = ops.tail_fwd(pqs, m.fuse.bias.detach(), n, h, w)
            ctx.objective = None
            ctx.saved = (xin, acts, pooled)
            return tuple(out[k] for k in range(5))
        label, weights, divisor = objective
        label = label.detach().to(xin.device).contiguous().float()
        if label.numel() != n * h * w:
            raise ValueError("objective label must be [N,1,H,W] like the output maps")
        weights = tuple(float(v) for v in weights)
        out, sums, losses = ops.tail_fwd(pqs, m.fuse.bias.detach(), n, h, w, label=label, loss_weights=weights,
                                         divisor=divisor)
        ctx.objective = (out, label, sums, weights, float(divisor))
        ctx.saved = (xin, acts, pooled)
        maps = tuple(out[k] for k in range(5))
        total = losses[5:6].reshape(())          # 0-dim view of the weighted total
        per_map = losses[0:5]
        ctx.mark_non_differentiable(*maps, per_map)
        return maps + (total, per_map)

    @staticmethod
    def backward(ctx, *grads):
        engine = ctx.engine
        m = engine.m
        fast = ctx.fast
        xin, acts, pooled = ctx.saved
        n, h, w = ctx.dims
        convs = _trunk_convs(m)
        pg = {}                                   # parameter -> gradient tensor
        obj = ctx.objective
        if obj is not None:
            g_total = grads[5]
            weights = obj[3]
            grads = tuple((True if weights[k] != 0.0 else None) for k in range(5)) if g_total is not None else (None,) * 5
        if all(g is None for g in grads):
            return (None, None, None) + tuple(None for _ in engine._param_list())
        # ---- weight-gradient plumbing: ONE zeroed arena for all tensor-core wgrad workspaces, ONE finish launch at
        # the end.  In direct mode (engine.accumulate_param_grads_in_place, set by the package's training loops) the
        # finish adds straight into an existing p.grad and the bias column sums are accumulated into p.grad by the
        # dgrad epilogues - what autograd's AccumulateGrad would otherwise do with one add kernel per parameter.
        direct = bool(getattr(engine, "accumulate_param_grads_in_place", False))

        def grad_target(p):
            g = p.grad
            if direct and p.requires_grad and g is not None and g.is_cuda and g.dtype == torch.float32 \
                    and g.is_contiguous() and g.device == xin.device:
                return g
            return None
        wconvs = [c for stage in convs for c in stage][1:]
        ws_sizes = [ops.wgrad_workspace_floats(c.out_channels, c.in_channels) for c in wconvs]
        # side branch: G [18 C + 2] per scale (rounded up to 16 bytes) behind the wgrad workspaces
        g_sizes = [(ops.side_folded_wgrad_floats(sp.in_channels) + 3) // 4 * 4 for sp in m.side_prep]
        arena = torch.zeros(sum(ws_sizes) + sum(g_sizes), dtype=torch.float32, device=xin.device)
        ws_of, off = {}, 0
        for c, sz in zip(wconvs, ws_sizes):
            ws_of[c] = arena[off:off + sz]
            off += sz
        g_of = []
        for sz in g_sizes:
            g_of.append(arena[off:off + sz])
            off += sz
        fresh = [c for c in wconvs if grad_target(c.weight) is None]
        fresh_buf = torch.empty(sum(c.weight.numel() for c in fresh), dtype=torch.float32, device=xin.device)
        finish_items, off = [], 0

        def wgrad(conv, inp, dz_act):
            nonlocal off
            it = ops.conv3x3_wgrad(inp, dz_act, conv.out_channels, fast=fast, deferred_ws=ws_of[conv])
            tgt = grad_target(conv.weight)
            if tgt is not None:
                it["dw"], it["accumulate"] = tgt, True
                pg[conv.weight] = None
            else:
                nel = conv.weight.numel()
                it["dw"], it["accumulate"] = fresh_buf[off:off + nel].view(conv.weight.shape), False
                off += nel
                pg[conv.weight] = it["dw"]
            finish_items.append(it)

        if obj is not None:
            # tail + loss backward in ONE launch: dL/dlogit is formed on the fly (never written), d fuse.bias comes from
            # the forward's sums
            out, label, sums, weights, divisor = obj
            dpq, fb = ops.tail_loss_bwd(out, label, sums, weights, divisor, g_total.detach().contiguous().float(),
                                        n, h, w, want_fuse_bias=grads[4] is not None)
            if fb is not None:
                pg[m.fuse.bias] = fb.reshape(m.fuse.bias.shape)
        else:
            dpq = ops.tail_bwd(list(grads), n, h, w)
            if grads[4] is not None:
                pg[m.fuse.bias] = ops.sum_f32(grads[4]).reshape(m.fuse.bias.shape)
        # one zeroed buffer for all 13 trunk bias gradients; the dgrad / unpool epilogues accumulate into its slices
        flat_convs = [c for stage in convs for c in stage]
        bias_buf = torch.zeros(sum(c.out_channels for c in flat_convs), dtype=torch.float32, device=xin.device)
        bias_slices, bias_direct, boff = {}, set(), 0
        for c in flat_convs:
            tgt = grad_target(c.bias)
            if tgt is not None:
                bias_direct.add(c)
            bias_slices[c] = tgt if tgt is not None else bias_buf[boff:boff + c.out_channels]
            boff += c.out_channels

        def bias_grad(c):                              # None: already accumulated into c.bias.grad
            return None if c in bias_direct else bias_slices[c]
        # every parameter gradient of the side branch from G[t][o][c] = sum_px dpq[px - t][o] x[px][c] (one pass over
        # the stage output per scale) and one finish launch for the four scales
        fold = engine._folded_side_all()
        side_params = [m.fuse.weight] if grads[4] is not None else []
        for i in range(4):
            side_params += [m.side_prep[i].weight, m.side_prep[i].bias]
            if grads[i] is not None:
                side_params += [m.score_dsn[i].weight, m.score_dsn[i].bias]
        in_place = all(grad_target(p) is not None for p in side_params)
        if not in_place:
            fresh_small = torch.zeros(4 * 34 + 64, dtype=torch.float32, device=xin.device)
            fresh_side = torch.empty(sum(sp.weight.numel() for sp in m.side_prep), dtype=torch.float32,
                                     device=xin.device)
        ops.side_folded_wgrad_multi([acts[i + 1][-1] for i in range(4)], dpq, g_of)      # G of the four scales, one launch
        entries, off_side = [], 0
        for i in range(4):
            sp, sd = m.side_prep[i], m.score_dsn[i]
            e = {"g": g_of[i], "side_w": sp.weight.detach(), "side_b": sp.bias.detach(), "proj_w": engine._proj(i),
                 "c": sp.in_channels}
            if in_place:
                e["d_side_w"], e["d_side_b"] = sp.weight.grad, sp.bias.grad
                pg[sp.weight] = pg[sp.bias] = None
                if grads[i] is not None:
                    e["d_score_w"], e["d_score_b"] = sd.weight.grad, sd.bias.grad
                    pg[sd.weight] = pg[sd.bias] = None
                if grads[4] is not None:
                    e["d_fuse_w"] = m.fuse.weight.grad.view(-1)[16 * i:16 * i + 16]
                    pg[m.fuse.weight] = None
            else:
                nel = sp.weight.numel()
                e["d_side_w"] = fresh_side[off_side:off_side + nel].view(sp.weight.shape)
                off_side += nel
                small = fresh_small[34 * i:34 * i + 34]
                e["d_side_b"] = small[0:16]
                pg[sp.weight], pg[sp.bias] = e["d_side_w"], e["d_side_b"]
                if grads[i] is not None:
                    e["d_score_w"], e["d_score_b"] = small[16:32], small[32:33]
                    pg[sd.weight] = e["d_score_w"].view(sd.weight.shape)
                    pg[sd.bias] = e["d_score_b"].view(sd.bias.shape)
                if grads[4] is not None:
                    e["d_fuse_w"] = fresh_small[136 + 16 * i:136 + 16 * i + 16]
            entries.append(e)
        if not in_place and grads[4] is not None:
            pg[m.fuse.weight] = fresh_small[136:200].view(m.fuse.weight.shape)
        ops.side_grads_finish(entries, accumulate=in_place)
        dpool = None
        for i in range(4, 0, -1):
            s_out = acts[i][-1]
            last_bias = bias_slices[convs[i][-1]]
            # ReLU'(x) * (unpool(dpool) + side gradient), the latter formed on the fly from dpq and the fp32 folded weights
            # (18 FMAs per element); deepest stage: dpool None, the side branch is the only consumer
            dz = ops.unpool_side_mask(dpool, s_out, dpq[i - 1], fold[i - 1][2], colsum=last_bias)
            for j in range(len(convs[i]) - 1, -1, -1):
                conv = convs[i][j]
                inp = acts[i][j - 1] if j > 0 else pooled[i]
                wgrad(conv, inp, dz)
                pg[conv.bias] = bias_grad(conv)
                wt = engine._packed(conv, f"s{i}c{j}", transpose_flip=True)
                if j > 0:
                    dz, _, _ = ops.conv3x3(dz, wt, None, conv.in_channels, fast=fast, mask=inp.hi,
                                           colsum=bias_slices[convs[i][j - 1]])
                else:
                    dpool, _, _ = ops.conv3x3(dz, wt, None, conv.in_channels, fast=fast)
        # stage 1 (no side branch)
        c12, c11 = convs[0][1], convs[0][0]
        dz = ops.unpool_add_mask(dpool, acts[0][1], None, colsum=bias_slices[c12])
        wgrad(c12, acts[0][0], dz)
        pg[c12.bias] = bias_grad(c12)
        dz, _, _ = ops.conv3x3(dz, engine._packed(c12, "s0c1", transpose_flip=True), None, 64, fast=fast,
                               mask=acts[0][0].hi, colsum=bias_slices[c11])
        dw0, dx = ops.conv_first_bwd(xin, dz, c11.weight.detach(), ctx.needs_input_grad[1])
        pg[c11.weight] = dw0
        pg[c11.bias] = bias_grad(c11)
        ops.wgrad_finish(finish_items)
        ctx.saved = None
        ctx.objective = None
        return (None, dx, None) + tuple(pg.get(p) for p in engine._param_list())


def osvos_apply(engine, x):
    params = engine._param_list()
    outs = _OSVOSFunction.apply(engine, x, None, *params)
    return list(outs)


def osvos_apply_objective(engine, x, label, loss_weights, divisor):
    """Forward + the weighted class-balanced BCE objective as one autograd node.
    -> (maps: list of 5 [N,1,H,W] logit tensors (detached), total: 0-dim differentiable loss, per_map: [5] losses)."""
    params = engine._param_list()
    res = _OSVOSFunction.apply(engine, x, (label, loss_weights, divisor), *params)
    return list(res[:5]), res[5], res[6]
