"""Test-time path of the reference (train_online.py:170-189) as a device pipeline.

The reference loops over the sequence one frame at a time: ``img.to(device)`` -> ``net.forward`` ->
``outputs[-1].cpu()`` -> numpy sigmoid -> ``scipy.misc.imsave``; every stage waits for the previous one.
``SequenceSegmenter`` keeps the same per-frame semantics but overlaps the three legs on separate CUDA streams with
a small ring of buffers: while frame i is in the network, frame i+1 is crossing PCIe host->device and the result
of frame i-1 is crossing device->host.  The result is either the fused logit map (fp32, what ``outputs[-1]`` holds)
or the 8-bit map produced on the device by ``ops.logits_to_u8`` (``bytescale`` = the PNG payload the reference
writes, ``prob``, ``mask``), which cuts the device->host bytes by 4.
"""
import torch

from . import ops


class SequenceSegmenter:
    """``for result in SequenceSegmenter(net)(frames): ...`` - ``frames`` yields host tensors [N,3,H,W] fp32 (pinned
    for real overlap); ``result`` is a pinned host tensor [N,1,H,W] (fp32 logits or uint8) that stays valid until
    ``depth`` further frames have been consumed.  All frames of one call must share a shape."""

    def __init__(self, net, output="logits", depth=3):
        if output not in ("logits", "bytescale", "prob", "mask"):
            raise ValueError("output must be one of logits / bytescale / prob / mask")
        self.net, self.output, self.depth = net, output, max(2, int(depth))
        self._shape = None

    def _allocate(self, shape, device):
        n, _, h, w = shape
        out_dtype = torch.float32 if self.output == "logits" else torch.uint8
        self._dev_in = [torch.empty(shape, dtype=torch.float32, device=device) for _ in range(self.depth)]
        self._dev_out = [torch.empty((n, 1, h, w), dtype=out_dtype, device=device) for _ in range(self.depth)]
        self._host_out = [torch.empty((n, 1, h, w), dtype=out_dtype).pin_memory() for _ in range(self.depth)]
        self._s_in, self._s_out = torch.cuda.Stream(device), torch.cuda.Stream(device)
        mk = lambda: [torch.cuda.Event() for _ in range(self.depth)]
        self._ev_loaded, self._ev_consumed, self._ev_done, self._ev_host = mk(), mk(), mk(), mk()
        self._shape = tuple(shape)
        self.h2d_bytes_per_frame = n * 3 * h * w * 4
        self.d2h_bytes_per_frame = n * h * w * (4 if self.output == "logits" else 1)

    def _submit(self, i, frame, device):
        k = i % self.depth
        cur = torch.cuda.current_stream(device)
        with torch.cuda.stream(self._s_in):
            if i >= self.depth:
                self._s_in.wait_event(self._ev_consumed[k])     # the network has read the previous tenant
            self._dev_in[k].copy_(frame, non_blocking=True)
            self._ev_loaded[k].record(self._s_in)
        cur.wait_event(self._ev_loaded[k])
        if i >= self.depth:
            cur.wait_event(self._ev_host[k])                    # previous result of this slot is on the host
        with torch.no_grad():
            eng = getattr(self.net, "_engine", None)
            if eng is not None:
                # the ring slot is a buffer that comes back: from its second frame on the engine replays a graph captured on
                # the slot itself (no input copy), and the fused map is read out of the graph's static output right here on
                # the same stream (no copy of the five maps into fresh tensors)
                fused = eng.forward(self._dev_in[k], fresh_outputs=False)[-1]
            else:
                fused = self.net(self._dev_in[k])[-1]
            self._ev_consumed[k].record(cur)
            if self.output == "logits":
                self._dev_out[k].copy_(fused)
            else:
                ops.logits_to_u8(fused, self.output, out=self._dev_out[k])
        self._ev_done[k].record(cur)
        with torch.cuda.stream(self._s_out):
            self._s_out.wait_event(self._ev_done[k])
            self._host_out[k].copy_(self._dev_out[k], non_blocking=True)
            self._ev_host[k].record(self._s_out)

    def __call__(self, frames):
        device = next(self.net.parameters()).device
        if device.type != "cuda":
            raise RuntimeError("SequenceSegmenter runs on CUDA only; there is no CPU fallback for the OSVOS hot path")
        submitted = 0
        with torch.cuda.device(device):
            for frame in frames:
                if frame.dim() == 3:
                    frame = frame.unsqueeze(0)
                if self._shape != tuple(frame.shape):
                    if submitted:
                        raise ValueError("all frames of one sequence must share a shape")
                    self._allocate(tuple(frame.shape), device)
                self._submit(submitted, frame, device)
                submitted += 1
                ready = submitted - self.depth + 1              # keep depth-1 frames in flight
                if ready >= 1:
                    k = (ready - 1) % self.depth
                    self._ev_host[k].synchronize()
                    yield self._host_out[k]
            for j in range(max(0, submitted - self.depth + 1), submitted):
                k = j % self.depth
                self._ev_host[k].synchronize()
                yield self._host_out[k]

    def join_current_stream(self):
        """Make the current stream wait for every copy issued so far (lets CUDA events on the current stream bracket
        the whole pipeline, as bench.py does)."""
        if self._shape is not None:
            cur = torch.cuda.current_stream()
            cur.wait_stream(self._s_in)
            cur.wait_stream(self._s_out)
