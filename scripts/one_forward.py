"""One eager 480x854 forward (no CUDA graph) - the process ncu wraps for per-kernel captures.

    ncu --set full --import-source on --clock-control none -k regex:conv3x3_halo -c 12 -o out python scripts/one_forward.py
"""
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
torch.set_grad_enabled(False)
from oracle import osvos_oracle as oc
from osvos_pytorch_b200.networks.vgg_osvos import OSVOS, he_init_

h, w = (int(sys.argv[1]), int(sys.argv[2])) if len(sys.argv) > 2 else (480, 854)
prec = sys.argv[3] if len(sys.argv) > 3 else "exact"
net = he_init_(OSVOS(pretrained=0, verbose=False, precision=prec)).cuda().eval()
net._engine.use_cuda_graph = False
x, _ = oc.synthetic_frame(1, h, w, 1234)
out = net(x.cuda())
torch.cuda.synchronize()
print("ok", float(out[-1].abs().max()))
