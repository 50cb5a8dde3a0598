"""Validates the halo-reuse conv kernel variants against the CUDA-core cross-check (one variant per process)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import math, torch
from osvos_pytorch_b200 import ops
torch.set_grad_enabled(False)
dev = torch.device("cuda:0")
g = torch.Generator().manual_seed(1)
worst = 0.0
for (n, h, w, cin, cout) in [(1, 16, 8, 64, 64), (1, 20, 13, 64, 64), (2, 17, 9, 128, 128), (1, 33, 45, 64, 128),
                             (1, 9, 11, 256, 256), (1, 3, 5, 512, 512), (1, 30, 27, 128, 16), (1, 64, 96, 64, 64)]:
    for fast in (False, True):
        x = torch.randn(n, cin, h, w, generator=g) * 3
        wt = torch.randn(cout, cin, 3, 3, generator=g) * math.sqrt(2.0 / (9 * cin))
        b = torch.randn(cout, generator=g) * 0.1
        a = ops.nchw_to_act(x.to(dev), fast)
        wp = ops.pack_conv3x3_weights(wt.to(dev))
        _, y, _ = ops.conv3x3(a, wp, b.to(dev), cout, relu=True, fast=fast, out_act=False, out_f32=True)
        _, ys, _ = ops.conv3x3(a, wp, b.to(dev), cout, relu=True, fast=fast, out_act=False, out_f32=True, simt=True)
        torch.cuda.synchronize()
        err = float((y - ys).abs().max() / ys.abs().max())
        worst = max(worst, err)
        print(f"  {n}x{h}x{w} {cin}->{cout} fast={int(fast)}: maxrel vs simt {err:.2e}")
print(f"VARIANT impl={os.environ.get('OSVOS_CONV_IMPL')} pitch={os.environ.get('OSVOS_HALO_PITCH')} bo={os.environ.get('OSVOS_HALO_BO')} worst={worst:.3e} {'OK' if worst < 1e-4 else 'WRONG'}")
