"""A/B of a per-launch environment switch of the library inside ONE process - development aid.

    python scripts/ab_env.py OSVOS_SPLITACC128 1 0 [H W]

For each value: the engine's CUDA graphs are dropped and re-captured, 480x854 inference is replayed 200 times over
four rotating frames (CUDA events), and the five output maps are compared with the first value's.
"""
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
torch.set_grad_enabled(False)
from oracle import osvos_oracle as oc
from osvos_pytorch_b200.networks.vgg_osvos import OSVOS, he_init_

var, values = sys.argv[1], sys.argv[2:4]
h, w = (int(sys.argv[4]), int(sys.argv[5])) if len(sys.argv) > 5 else (480, 854)
net = he_init_(OSVOS(pretrained=0, verbose=False)).cuda().eval()
xs = [oc.synthetic_frame(1, h, w, 1234 + i)[0].cuda() for i in range(4)]
ref = None
for rnd in range(2):                      # two rounds: the second repeats the measurement in reverse order
    for v in (values if rnd == 0 else values[::-1]):
        os.environ[var] = v
        net._engine._graphs.clear()
        for i in range(8):
            net(xs[i % 4])
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(200):
            net(xs[i % 4])
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 200
        outs = [o.clone() for o in net(xs[0])]
        if ref is None:
            ref = outs
        err = max(float((a - b).abs().max() / b.abs().max()) for a, b in zip(outs, ref))
        print(f"{var}={v}: {ms:.4f} ms/frame = {1000 / ms:.1f} fps   max-rel difference to the first variant {err:.2e}")
