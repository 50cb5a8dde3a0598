"""BASELINE.json configs[4]: resolution sweep 240p / 480p / 720p / 1080p, batch 1, fwd and fwd+bwd:
fps, algorithmic conv TFLOP/s and fraction of the measured bf16 peak (development report, not the bench)."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from oracle import osvos_oracle as oc
from osvos_pytorch_b200.networks.vgg_osvos import OSVOS, he_init_
from osvos_pytorch_b200.layers.osvos_layers import class_balanced_cross_entropy_loss as cbce

peak = 1455.4
try:
    peak = json.load(open(os.path.join(os.path.dirname(__file__), "..", "MEASURED_PEAKS.json")))["bf16_tflops_sustained"]
except Exception:
    pass
prec = sys.argv[1] if len(sys.argv) > 1 else "exact"
net = he_init_(OSVOS(pretrained=0, verbose=False, precision=prec)).cuda()

def timed(fn, reps):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps

print(f"precision={prec}; peak = measured sustained bf16 {peak} TFLOP/s; conv flops = 2MNK over the 17 3x3 convs")
print(f"{'res':>10s} {'fwd ms':>8s} {'fwd fps':>8s} {'TF/s':>7s} {'frac':>6s} | {'f+b ms':>8s} {'f+b fps':>8s} {'TF/s':>7s} {'frac':>6s}")
for (h, w) in ((240, 427), (480, 854), (720, 1280), (1080, 1920)):
    x, gt = oc.synthetic_frame(1, h, w, 1234)
    x, gt = x.cuda(), gt.cuda()
    fl = oc.conv_flops(h, w)
    net.eval()
    with torch.no_grad():
        ms_f = timed(lambda: net(x), 30)
    net.train()
    # fwd + online objective + bwd as the replayed CUDA graph the training loops use (training.GraphedTrainStep); eager
    # launches are host-bound below ~720p (50 launches + autograd per step) and would time the Python side
    from osvos_pytorch_b200.training import GraphedTrainStep, ONLINE_WEIGHTS
    net._engine.drop_derived_caches()
    gstep = GraphedTrainStep(net, ONLINE_WEIGHTS, {"image": x, "gt": gt})
    ms_b = timed(lambda: gstep(), 20)
    del gstep
    net._engine.drop_derived_caches()
    tf_f, tf_b = fl / ms_f / 1e9, 3 * fl / ms_b / 1e9
    print(f"{h:>4d}x{w:<5d} {ms_f:8.3f} {1000/ms_f:8.1f} {tf_f:7.1f} {tf_f/peak:6.3f} | {ms_b:8.3f} {1000/ms_b:8.1f} {tf_b:7.1f} {tf_b/peak:6.3f}")
