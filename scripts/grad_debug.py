"""Development aid: where does the whole-network gradient error come from?
(a) native fwd + native loss + native bwd vs oracle; (b) native bwd fed with the ORACLE's dL/dlogit."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from oracle import osvos_oracle as oc
from osvos_pytorch_b200.networks.vgg_osvos import OSVOS
from osvos_pytorch_b200.layers.osvos_layers import class_balanced_cross_entropy_loss as cbce

h, w = (int(sys.argv[1]), int(sys.argv[2])) if len(sys.argv) > 2 else (40, 56)
params = oc.he_params(seed=0)
net = OSVOS(pretrained=0, verbose=False)
net.load_state_dict(params, strict=False)
net = net.cuda().train()
x, gt = oc.synthetic_frame(1, h, w, 21)
ref_loss, ref_outs, og = oc.forward_backward(params, x, gt, objective="online")

def rel(a, b):
    a, b = a.detach().double().cpu(), b.detach().double().cpu()
    return float((a - b).norm() / b.norm())

net.zero_grad()
outs = net(x.cuda())
cbce(outs[-1], gt.cuda(), size_average=False).backward()
ga = {n: p.grad.clone() for n, p in net.named_parameters() if p.grad is not None}
net.zero_grad()
outs = net(x.cuda())
g_or = oc.class_balanced_cross_entropy_grad(ref_outs[-1], gt, size_average=False)
torch.autograd.backward([outs[-1]], [g_or.cuda()])
gb = {n: p.grad.clone() for n, p in net.named_parameters() if p.grad is not None}
print(f"{h}x{w}: logit err {rel(outs[-1], ref_outs[-1]):.2e}")
print(f"{'param':28s} {'(a) full':>10s} {'(b) oracle dL/dx':>16s}")
for n in og:
    if n in ga:
        print(f"{n:28s} {rel(ga[n], og[n]):10.2e} {rel(gb[n], og[n]):16.2e}")

# ReLU-mask agreement between the native forward activations and the oracle's (explains step-wise jumps above)
from osvos_pytorch_b200 import ops
with torch.no_grad():
    stages = oc.trunk_forward(params, x)
    _, inter = net._engine.forward_inference(x.cuda(), return_intermediates=True)
for i in range(5):
    got = ops.act_to_nchw(inter[f"stage{i}"]).cpu()
    flips = int(((got > 0) != (stages[i] > 0)).sum())
    print(f"stage {i} output: {flips} ReLU-mask flips of {got.numel()} (active {int((stages[i] > 0).sum())})")
