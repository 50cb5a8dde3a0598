"""Execution engine behind ``OSVOS.forward``: walks the module's parameter
containers and enqueues the native kernels (include/osvos_b200.h).  Python here
is plumbing - packing caches keyed on parameter versions, buffer allocation,
stream handling; all arithmetic is in csrc/.

Call graph replaced: reference networks/vgg_osvos.py:59-74 (forward) and, in
training, the autograd graph PyTorch builds for it.
"""
import os

import torch
import torch.nn as nn

from . import ops
from .layers.osvos_layers import bilinear_deconv_weight


class OSVOSEngine:
    def __init__(self, module):
        # no reference cycle through nn.Module registration: keep a plain attribute
        object.__setattr__(self, "m", module)
        self._pack_cache = {}
        self._deconv_checked = {}
        # CUDA-graph cache for the inference path: (shape, device, precision, parameter versions) -> captured step.
        # One frame is ~22 kernel launches; replaying a graph removes the Python / launch overhead (OSVOS_CUDA_GRAPH=0
        # disables it).
        self._graphs = {}
        self._buffers_seen = {}          # (graph key, input address) -> calls seen (engine._forward_graphed)
        self._graphs_pver = None         # parameter versions the captured graphs belong to
        # graphs bound to input buffers: at most this many (each pins its activation pool); once reached, further
        # buffers go through the generic entry (one input copy) - no eviction, hence no capture / evict churn
        self.max_direct_graphs = 12
        self.use_cuda_graph = os.environ.get("OSVOS_CUDA_GRAPH", "1") != "0"
        # training loops of this package set this: backward adds weight / trunk-bias gradients straight into an
        # existing p.grad (and hands autograd None for them) instead of returning tensors for AccumulateGrad
        self.accumulate_param_grads_in_place = False
        # inference: fold side_prep with its two 1x1 projections into one 3x3 conv C -> 2 (training always does; OSVOS_FOLD_SIDE=0, read per
        # eager pass, switches it off for A/B runs)
        self.fold_side_branch = True

    def direct_grad_accumulation(self):
        """Context manager enabling in-place gradient accumulation for the backward passes run inside it."""
        eng = self

        class _Ctx:
            def __enter__(self):
                self.prev = eng.accumulate_param_grads_in_place
                eng.accumulate_param_grads_in_place = True

            def __exit__(self, *exc):
                eng.accumulate_param_grads_in_place = self.prev
                return False
        return _Ctx()

    # ------------------------------------------------------------ weight caches
    def _cached(self, key, params, make):
        ver = tuple((p.data_ptr(), p._version) for p in params)
        hit = self._pack_cache.get(key)
        if hit is not None and hit[0] == ver:
            return hit[1]
        val = make()
        self._pack_cache[key] = (ver, val)
        return val

    def _packed(self, conv, key, transpose_flip=False, col_pad=64):
        return self._cached((key, transpose_flip, col_pad), [conv.weight],
                            lambda: ops.pack_conv3x3_weights(conv.weight, transpose_flip, col_pad))

    def _tensor_core_convs(self):
        """(conv module, cache key) of every trunk 3x3 conv that runs on the packed tensor-core path (conv1_1 takes its
        OIHW weights directly; side_prep is folded with its 1x1 projections instead)."""
        m = self.m
        out = []
        for i in range(5):
            convs = [c for c in m.stages[i] if isinstance(c, nn.Conv2d)]
            for j, conv in enumerate(convs):
                if i == 0 and j == 0:
                    continue
                out.append((conv, f"s{i}c{j}"))
        return out                # (side_prep runs folded with its projections: engine._folded_side_all)

    def packed_weight_table(self):
        """[(weight Parameter, forward layout, transposed+flipped layout)] of the tensor-core convs, packed now if
        stale.  optim.FusedSGD rewrites these buffers in place from the updated weights."""
        return [(conv.weight, self._packed(conv, key), self._packed(conv, key, transpose_flip=True))
                for conv, key in self._tensor_core_convs()]

    def restamp_packed(self):
        """Declare the cached layouts current for the present parameter versions (called by optim.FusedSGD after it
        has updated the weights AND their packed layouts in one kernel)."""
        for conv, key in self._tensor_core_convs():
            ver = ((conv.weight.data_ptr(), conv.weight._version),)
            for flip in (False, True):
                hit = self._pack_cache.get((key, flip, 64))
                if hit is not None:
                    self._pack_cache[(key, flip, 64)] = (ver, hit[1])

    def drop_derived_caches(self, keep_packed=False):
        """Forget cached derived tensors; with keep_packed the packed conv layouts (static buffers a captured graph
        may point at) are kept."""
        if not keep_packed:
            self._pack_cache.clear()
            return
        for k in [k for k in self._pack_cache if not (isinstance(k, tuple) and len(k) == 3 and k[2] == 64
                                                      and isinstance(k[1], bool))]:
            del self._pack_cache[k]

    def _param_list(self):
        """Parameters the native path differentiates (everything except the fixed deconvolution taps)."""
        m = self.m
        return ([p for p in m.stages.parameters()] + [p for p in m.side_prep.parameters()]
                + [p for p in m.score_dsn.parameters()] + [m.fuse.weight, m.fuse.bias])

    def _proj(self, i):
        m = self.m
        sd, fu = m.score_dsn[i], m.fuse
        return self._cached(("proj", i), [sd.weight, fu.weight],
                            lambda: torch.cat([sd.weight.detach().reshape(16),
                                               fu.weight.detach().reshape(64)[16 * i:16 * i + 16]]).float().contiguous())

    def _folded_side(self, i):
        """(packed [2,C,3,3] operand, bias2) of side_prep[i] folded with score_dsn[i] and fuse's slice."""
        return self._folded_side_all()[i][:2]

    def _folded_side_all(self):
        """[(packed operand, bias2, fp32 W' [9,2,C])] of the four side scales, folded by ONE launch and cached on the
        versions of every parameter that enters (training re-folds after each optimizer step)."""
        m = self.m
        deps = [m.fuse.weight]
        for i in range(4):
            deps += [m.side_prep[i].weight, m.side_prep[i].bias, m.score_dsn[i].weight, m.score_dsn[i].bias]
        return self._cached(("fold_all",), deps, lambda: ops.fold_side_weights_multi(
            [(m.side_prep[i].weight, m.side_prep[i].bias.detach(), self._proj(i), m.score_dsn[i].bias.detach())
             for i in range(4)]))

    def _check_deconvs(self):
        """The native tail implements the bilinear deconvolution in closed form; both entry points of
        the reference keep these weights fixed (lr = 0, train_online.py:84-85, train_parent.py:99-100).
        Anything else is refused loudly rather than computed wrongly."""
        m = self.m
        for name, lst in (("upscale", m.upscale), ("upscale_", m.upscale_)):
            for i, lay in enumerate(lst):
                w = lay.weight
                key = (name, i)
                ver = (w.data_ptr(), w._version)
                if self._deconv_checked.get(key) == ver:
                    continue
                ref = bilinear_deconv_weight(w.shape[0], w.shape[1], w.shape[2]).to(w.device)
                if tuple(w.shape[2:]) != (2 ** (i + 2),) * 2 or not torch.equal(w.detach().float(), ref):
                    raise NotImplementedError(
                        f"{name}.{i}.weight is not the fixed bilinear interpolation kernel written by interp_surgery; "
                        "the B200 path only implements that (reference layers/osvos_layers.py:72-85)")
                self._deconv_checked[key] = ver

    # ----------------------------------------------------------------- forward
    def forward(self, x, fresh_outputs=True):
        if not isinstance(x, torch.Tensor) or x.dim() != 4 or x.size(1) != 3:
            raise ValueError("OSVOS.forward expects a [N, 3, H, W] tensor")
        if not x.is_cuda:
            raise RuntimeError("osvos_pytorch_b200.OSVOS runs on CUDA (sm_100a) only: move the module and the input "
                               "to the GPU.  There is no CPU fallback for the hot path.")
        needs_grad = torch.is_grad_enabled() and (x.requires_grad or any(p.requires_grad for p in self.m.parameters()))
        if x.device != torch.device("cuda", torch.cuda.current_device()):
            with torch.cuda.device(x.device):       # kernels are enqueued on the current stream of x's device
                return self.forward(x, fresh_outputs)
        if needs_grad:
            from .autograd import osvos_apply
            return osvos_apply(self, x)
        if self.use_cuda_graph and not torch.cuda.is_current_stream_capturing():
            return self._forward_graphed(x, fresh_outputs)
        return self.forward_inference(x)

    def forward_objective(self, x, gts, loss_weights, size_average=False, batch_average=True):
        """See OSVOS.forward_objective."""
        if not isinstance(x, torch.Tensor) or x.dim() != 4 or x.size(1) != 3:
            raise ValueError("OSVOS.forward_objective expects a [N, 3, H, W] tensor")
        if not x.is_cuda:
            raise RuntimeError("osvos_pytorch_b200.OSVOS runs on CUDA (sm_100a) only; there is no CPU fallback")
        if len(loss_weights) != 5:
            raise ValueError("loss_weights: one weight per output map (5)")
        if size_average:
            divisor = float(gts.numel())
        elif batch_average:
            divisor = float(gts.size(0))
        else:
            divisor = 1.0
        if x.device != torch.device("cuda", torch.cuda.current_device()):
            with torch.cuda.device(x.device):
                return self.forward_objective(x, gts, loss_weights, size_average, batch_average)
        from .autograd import osvos_apply_objective
        return osvos_apply_objective(self, x, gts, loss_weights, divisor)

    def _forward_graphed(self, x, fresh_outputs=True):
        """Inference through a captured CUDA graph.  Two kinds of entry:
        * generic: the frame is copied into the graph's static input, the graph replayed (any input tensor);
        * direct: an input BUFFER that comes back (same address, shape, contiguous fp32 - the second call on) gets a graph
          captured on that buffer itself: no input copy.  A video pipeline feeds a small ring of device buffers
          (inference.SequenceSegmenter does), so in steady state every replay is direct.  At most `max_direct_graphs`
          of them; all graphs are dropped when a parameter changes.
        The five maps are handed back as fresh tensors (one device copy) unless `fresh_outputs` is False: then they are
        views of the entry's static output, valid until the SAME entry is replayed again (the sequence pipeline reads them
        straight into its pinned host buffers).  Re-captured when shapes or parameters change."""
        m = self.m
        pver = tuple((p.data_ptr(), p._version) for p in m.parameters())
        if pver != self._graphs_pver:                            # parameters changed: every captured graph is stale
            self._graphs.clear()
            self._buffers_seen.clear()
            self._graphs_pver = pver
        pkey = (tuple(x.shape), x.device.index, m.precision)
        direct_ok = x.dtype == torch.float32 and x.is_contiguous() and not x.requires_grad
        entry = None
        if direct_ok:
            dkey = pkey + (x.data_ptr(),)
            entry = self._graphs.get(dkey)
            if entry is None and sum(1 for k in self._graphs if len(k) == 4) < self.max_direct_graphs:
                seen = self._buffers_seen.get(dkey, 0) + 1
                if len(self._buffers_seen) > 64:
                    self._buffers_seen.clear()
                self._buffers_seen[dkey] = seen
                if seen >= 2:                                    # the buffer came back: give it its own graph
                    entry = self._capture(dkey, x, direct=True)
        if entry is None:
            entry = self._graphs.get(pkey)
            if entry is None:
                generic = [k for k in self._graphs if len(k) == 3]
                if len(generic) >= 4:                            # bounded: each entry pins its activation pool
                    self._graphs.pop(generic[0])
                entry = self._capture(pkey, x, direct=False)
        graph, static_x, outs, base = entry
        if static_x is not None:
            static_x.copy_(x, non_blocking=True)
        graph.replay()
        if not fresh_outputs:
            return list(outs)
        if base is not None:
            fresh = base.clone()
            n, _, h, w = (int(v) for v in x.shape)
            return [fresh[k, :n * h * w].view(n, 1, h, w) for k in range(5)]
        return [o.clone() for o in outs]

    def _capture(self, key, x, direct):
        """Capture the inference pass for `key`; direct: on the caller's buffer itself (no static input)."""
        self.forward_inference(x)                                # eager warm-up: packs weights, sets kernel attributes
        static_x = None if direct else x.detach().contiguous().float().clone()
        torch.cuda.synchronize(x.device)
        graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(graph):
            outs = self.forward_inference(x if direct else static_x)
        base = outs[0]._base if outs[0]._base is not None else None
        entry = (graph, static_x, outs, base)
        self._graphs[key] = entry
        return entry

    @torch.no_grad()
    def forward_inference(self, x, simt=False, return_intermediates=False):
        """The inference pass (eager, or while a CUDA graph captures it).  Programmatic dependent launch is switched on
        for its kernels (every one of them waits before touching its predecessor's data, so only prologues overlap);
        OSVOS_PDL_INFER=0 keeps plain stream order."""
        from . import _native as nat
        lib = nat.load()
        prev = lib.osvos_set_pdl(1) if os.environ.get("OSVOS_PDL_INFER", "1") != "0" else None
        try:
            return self._forward_inference(x, simt, return_intermediates)
        finally:
            if prev is not None:
                lib.osvos_set_pdl(prev)

    def _forward_inference(self, x, simt=False, return_intermediates=False):
        m = self.m
        fast = m.precision == "fast"
        self._check_deconvs()
        x = x.detach().contiguous().float()
        n, _, h, w = (int(v) for v in x.shape)
        inter = {}
        convs0 = [c for c in m.stages[0] if isinstance(c, nn.Conv2d)]
        if not fast and not simt and os.environ.get("OSVOS_FUSE_STAGE1", "1") != "0":
            # stage 1 as one kernel: conv1_1 is computed inside conv1_2's kernel on its halo patch (no 105 MB round trip)
            full, a = ops.stage1_fused(x, convs0[0].weight.detach().contiguous().float(), convs0[0].bias.detach(),
                                       self._packed(convs0[1], "s0c1"), convs0[1].bias.detach(), pool=True,
                                       out_act=return_intermediates)
        else:
            a = ops.conv_first(x, convs0[0].weight.detach(), convs0[0].bias.detach(), relu=True, fast=fast)
            # conv1_2 with the 2x2 max pool fused into its epilogue; the full-resolution map is only kept on request
            full, a = ops.conv3x3(a, self._packed(convs0[1], "s0c1"), convs0[1].bias.detach(), convs0[1].out_channels,
                                  relu=True, fast=fast, simt=False, pool=True, out_act=return_intermediates)
        if return_intermediates:
            inter["stage0"] = full
        pqs = []
        fold = not simt and not return_intermediates and self.fold_side_branch and \
            os.environ.get("OSVOS_FOLD_SIDE", "1") != "0"
        # folded side branches of the four scales in ONE launch after the last trunk conv (OSVOS_SIDE_MULTI=0: one launch
        # per scale, right after its stage)
        multi = fold and os.environ.get("OSVOS_SIDE_MULTI", "1") != "0"
        stage_outs = []
        for i in range(1, 5):
            convs = [c for c in m.stages[i] if isinstance(c, nn.Conv2d)]
            for j, conv in enumerate(convs):
                if j == len(convs) - 1 and i < 4 and not simt:
                    full, a = ops.conv3x3(a, self._packed(conv, f"s{i}c{j}"), conv.bias.detach(), conv.out_channels,
                                          relu=True, fast=fast, pool=True)
                else:
                    a, _, _ = ops.conv3x3(a, self._packed(conv, f"s{i}c{j}"), conv.bias.detach(), conv.out_channels,
                                          relu=True, fast=fast, simt=simt)
                    full = a
            if simt and i < 4:
                a = ops.maxpool2x2(full)
            if return_intermediates:
                inter[f"stage{i}"] = full
            sp = m.side_prep[i - 1]
            if simt:
                _, feat, _ = ops.conv3x3(full, self._packed(sp, f"sp{i}"), sp.bias.detach(), 16, relu=False, fast=fast,
                                         out_act=False, out_f32=True, simt=True)
                pq = ops.side_project(feat, self._proj(i - 1), m.score_dsn[i - 1].bias.detach())
            elif multi:
                stage_outs.append(full)
                continue
            elif fold:
                # inference: side_prep o (score_dsn, fuse slice) folded into one 3x3 conv C -> 2 (include/osvos_b200.h)
                feat = None
                pq = ops.side_folded(full, *self._folded_side(i - 1), fast=fast)
            else:
                _, feat, pq = ops.conv3x3(full, self._packed(sp, f"sp{i}"), sp.bias.detach(), 16, relu=False, fast=fast,
                                          out_act=False, out_f32=return_intermediates, proj_w=self._proj(i - 1),
                                          proj_b=m.score_dsn[i - 1].bias.detach())
            if return_intermediates:
                inter[f"side{i}"] = feat
                inter[f"pq{i}"] = pq
            pqs.append(pq)
        if multi:
            pqs = ops.side_folded_multi(stage_outs, self._folded_side_all(), fast=fast)
        out, _ = ops.tail_fwd(pqs, m.fuse.bias.detach(), n, h, w)
        outs = [out[k] for k in range(5)]
        if return_intermediates:
            return outs, inter
        return outs
