"""Fused momentum-SGD step for the OSVOS parameter groups.

Replaces ``torch.optim.SGD.step()`` (+ ``zero_grad()``) as the reference uses it (train_online.py:77-88,147-148;
train_parent.py:85-103,170-171): ONE kernel launch updates every trainable tensor with its group's lr / weight
decay / momentum, optionally zeroes the gradients in the same pass, and re-emits the tensor-core operand layouts
of the 3x3 conv weights (forward and transposed+flipped) from the updated values, so the engine does not repack
after the step.  Same constructor and ``param_groups`` / ``state`` layout as ``torch.optim.SGD`` (state key
``momentum_buffer``), so checkpoints of one load into the other.  dampening = 0, nesterov = False only.
"""
import ctypes

import torch

from . import _native as nat
from . import ops


class FusedSGD(torch.optim.Optimizer):
    def __init__(self, params, lr=1e-3, momentum=0.0, weight_decay=0.0, dampening=0.0, nesterov=False, engine=None):
        if dampening != 0.0 or nesterov:
            raise NotImplementedError("FusedSGD implements the reference's configuration: dampening 0, no nesterov")
        super().__init__(params, dict(lr=lr, momentum=momentum, weight_decay=weight_decay))
        self._engine = engine                       # OSVOSEngine whose packed layouts are re-emitted, or None
        self._table = None
        self._table_key = None
        self._total_items = 0
        self._entries = []

    # ------------------------------------------------------------------ descriptor table
    def _build(self):
        lib = nat.load()
        packed = {}
        if self._engine is not None:
            for w, fwd, flip in self._engine.packed_weight_table():
                packed[id(w)] = (fwd, flip)
        entries, key = [], []
        for group in self.param_groups:
            for p in group["params"]:
                if p.grad is None or not p.requires_grad:
                    continue                           # as torch.optim.SGD: tensors without a gradient are skipped
                if not (p.is_cuda and p.grad.is_cuda and p.dtype == torch.float32 and p.is_contiguous()
                        and p.grad.is_contiguous()):
                    raise RuntimeError("FusedSGD: parameters and gradients must be contiguous fp32 CUDA tensors; "
                                       "there is no CPU fallback")
                st = self.state[p]
                if st.get("momentum_buffer") is None:
                    st["momentum_buffer"] = torch.zeros_like(p, memory_format=torch.contiguous_format)
                pk = packed.get(id(p))
                entries.append((p, group, st["momentum_buffer"], pk))
                key.append((p.data_ptr(), p.grad.data_ptr(), st["momentum_buffer"].data_ptr(), float(group["lr"]),
                            float(group["weight_decay"]), float(group["momentum"]),
                            None if pk is None else (pk[0].data_ptr(), pk[1].data_ptr())))
        key = tuple(key)
        if key == self._table_key:
            return
        if not entries:
            self._entries, self._table, self._table_key, self._total_items = [], None, key, 0
            return
        if len(entries) > nat.SGD_MAX_SEGMENTS:
            raise RuntimeError(f"FusedSGD: {len(entries)} tensors > OSVOS_SGD_MAX_SEGMENTS")
        arr = (nat.SgdSegment * len(entries))()
        total = 0
        for seg, (p, group, buf, pk) in zip(arr, entries):
            seg.param, seg.grad, seg.momentum = p.data_ptr(), p.grad.data_ptr(), buf.data_ptr()
            seg.numel = p.numel()
            seg.lr, seg.weight_decay, seg.momentum_coef = group["lr"], group["weight_decay"], group["momentum"]
            if pk is not None:
                cout, cin = int(p.shape[0]), int(p.shape[1])
                seg.cout, seg.cin = cout, cin
                seg.colp_fwd, seg.colp_flip = (cin + 63) // 64 * 64, (cout + 63) // 64 * 64
                seg.packed_fwd, seg.packed_flip = pk[0].data_ptr(), pk[1].data_ptr()
                seg.work_items = lib.osvos_sgd_work_items(p.numel(), cout, cin)
                if seg.work_items == 0:
                    raise RuntimeError(f"FusedSGD: conv weight {tuple(p.shape)} does not tile (cout % 16, cin % 32)")
            else:
                seg.work_items = lib.osvos_sgd_work_items(p.numel(), 0, 0)
            total += seg.work_items
        raw = torch.frombuffer(bytearray(ctypes.string_at(ctypes.addressof(arr), ctypes.sizeof(arr))), dtype=torch.uint8)
        self._table = raw.to(entries[0][0].device)
        self._entries, self._table_key, self._total_items = entries, key, total

    # ------------------------------------------------------------------ step
    @torch.no_grad()
    def step(self, closure=None, zero_grad=False):
        """One SGD update.  ``zero_grad=True`` also clears the gradients in the same kernel (the reference calls
        ``optimizer.zero_grad()`` right after ``step()``)."""
        loss = None
        if closure is not None:
            with torch.enable_grad():
                loss = closure()
        self._build()
        if not self._entries:
            return loss
        lib = nat.load()
        dev = self._entries[0][0].device
        with torch.cuda.device(dev):
            ops._count(1)
            nat.check(lib.osvos_sgd_step(self._table.data_ptr(), len(self._entries), self._total_items,
                                         1 if zero_grad else 0, torch.cuda.current_stream().cuda_stream),
                      "osvos_sgd_step")
        params = [e[0] for e in self._entries]
        torch.autograd.graph.increment_version(params)      # caches keyed on Tensor._version see the update
        if zero_grad:
            torch.autograd.graph.increment_version([p.grad for p in params])
        if self._engine is not None:
            self._engine.restamp_packed()
        return loss
