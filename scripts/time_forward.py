"""Per-kernel CUDA-event timing of one 480x854 forward (development aid, not the bench)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
torch.set_grad_enabled(False)
from oracle import osvos_oracle as oc
from osvos_pytorch_b200.networks.vgg_osvos import OSVOS, he_init_

h, w = (int(sys.argv[1]), int(sys.argv[2])) if len(sys.argv) > 2 else (480, 854)
prec = sys.argv[3] if len(sys.argv) > 3 else "exact"
net = he_init_(OSVOS(pretrained=0, verbose=False, precision=prec)).cuda().eval()
x, _ = oc.synthetic_frame(1, h, w, 1234)
x = x.cuda()
for _ in range(3):
    net(x)
torch.cuda.synchronize()
# whole forward
ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
reps = 20
ev[0].record()
for _ in range(reps):
    net(x)
ev[1].record()
torch.cuda.synchronize()
ms = ev[0].elapsed_time(ev[1]) / reps
print(f"forward {h}x{w} {prec}: {ms:.3f} ms/frame = {1000/ms:.1f} fps; conv TFLOP/s (algorithmic) {oc.conv_flops(h, w)/ms/1e9:.1f}")

net._engine.use_cuda_graph = False
for _ in range(3):
    net(x)
torch.cuda.synchronize()
ev[0].record()
for _ in range(reps):
    net(x)
ev[1].record()
torch.cuda.synchronize()
ms_e = ev[0].elapsed_time(ev[1]) / reps
print(f"  (eager, no CUDA graph: {ms_e:.3f} ms/frame = {1000/ms_e:.1f} fps)")
# per-op timing by wrapping ops
import osvos_pytorch_b200.ops as O
rec = []
def wrap(name):
    f = getattr(O, name)
    def g(*a, **k):
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record(); r = f(*a, **k); e.record()
        desc = name
        if name == "conv3x3":
            desc += " " + "x".join(str(v) for v in a[0].shape) + f"->{a[3]}"
        rec.append((desc, s, e))
        return r
    setattr(O, name, g)
for nme in ("conv_first", "conv3x3", "maxpool2x2", "tail_fwd"):
    wrap(nme)
net(x)
torch.cuda.synchronize()
tot = 0
for d, s, e in rec:
    t = s.elapsed_time(e); tot += t
    extra = ""
    if d.startswith("conv3x3"):
        n_, hh, ww, ci = (int(v) for v in d.split(" ")[1].split("->")[0].split("x")); co = int(d.split("->")[1])
        fl = 2.0 * n_ * hh * ww * co * 9 * ci
        extra = f"  {fl/t/1e9:8.1f} TFLOP/s"
    print(f"  {d:45s} {t*1000:9.1f} us{extra}")
print(f"  sum {tot:.3f} ms")
