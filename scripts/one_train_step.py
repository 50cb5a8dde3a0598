"""One eager 480x854 fwd + online loss + bwd (no CUDA graph) - the process ncu wraps for per-kernel captures of the
training kernels (wgrad, dgrad, unpool, tail / side backward).

    ncu --section SourceCounters --section SpeedOfLight --import-source on --clock-control none \
        -k regex:"wgrad_tc|unpool|conv_first_wgrad" -c 12 -o out python scripts/one_train_step.py
"""
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from oracle import osvos_oracle as oc
from osvos_pytorch_b200.layers.osvos_layers import class_balanced_cross_entropy_loss as cbce
from osvos_pytorch_b200.networks.vgg_osvos import OSVOS, he_init_

argv = [a for a in sys.argv[1:] if not a.startswith("--")]
h, w = (int(argv[0]), int(argv[1])) if len(argv) >= 2 else (480, 854)
net = he_init_(OSVOS(pretrained=0, verbose=False)).cuda().train()
x, gt = oc.synthetic_frame(1, h, w, 1234)
if "--unfused" in sys.argv:      # the reference's call sequence: forward, separate loss, autograd
    loss = cbce(net(x.cuda())[-1], gt.cuda(), size_average=False)
else:                            # the package's fused objective (tail + loss one kernel each way)
    _, loss, _ = net.forward_objective(x.cuda(), gt.cuda(), (0.0, 0.0, 0.0, 0.0, 1.0))
loss.backward()
torch.cuda.synchronize()
print("ok", float(loss))
