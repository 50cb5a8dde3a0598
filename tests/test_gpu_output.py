"""SURVEY.md 8(f) item 4 - test-time output path: on-device 8-bit maps and the pipelined sequence loop, against the
oracle's restatement of the reference's numpy sigmoid + scipy.misc.imsave (train_online.py:181-187)."""
import numpy as np
import pytest
import torch

from oracle import osvos_oracle as oc

pytestmark = pytest.mark.gpu


def _net(precision="exact"):
    from osvos_pytorch_b200.networks.vgg_osvos import OSVOS, he_init_
    return he_init_(OSVOS(pretrained=0, verbose=False, precision=precision), seed=0).cuda().eval()


@pytest.mark.parametrize("shape", [(1, 1, 48, 70), (3, 1, 33, 45), (2, 1, 480, 854), (1, 1, 7, 5)])
def test_logits_to_u8_matches_reference_png_payload(shape):
    from osvos_pytorch_b200 import ops
    g = torch.Generator().manual_seed(5)
    x = (torch.randn(shape, generator=g) * 6.0).float()
    xc = x.cuda()
    got = ops.logits_to_u8(xc, "bytescale").cpu().numpy()
    for f in range(shape[0]):
        want = oc.png_payload(x[f, 0].numpy())
        diff = np.abs(got[f, 0].astype(np.int32) - want.astype(np.int32))
        # rounding of (p - pmin)*scale + 0.5 at exact .5 boundaries may differ by one code between expf and np.exp
        assert diff.max() <= 1
        assert (diff != 0).mean() < 2e-3
    prob = ops.logits_to_u8(xc, "prob").cpu().numpy().astype(np.int32)
    want = np.floor(255.0 / (1.0 + np.exp(-x.numpy().astype(np.float64))) + 0.5).astype(np.int32)
    assert np.abs(prob - want).max() <= 1 and (prob != want).mean() < 2e-3
    mask = ops.logits_to_u8(xc, "mask").cpu().numpy()
    assert np.array_equal(mask, np.where(x.numpy() > 0, 255, 0).astype(np.uint8))       # bit-exact


def test_bytescale_edge_cases():
    from osvos_pytorch_b200 import ops
    const = torch.full((1, 1, 8, 12), 1.5).cuda()
    assert int(ops.logits_to_u8(const, "bytescale").max()) == 0                        # cmax == cmin -> all zero
    two = torch.tensor([[-3.0, 4.0, 4.0, -3.0, 0.5]]).view(1, 1, 1, 5).cuda()
    got = ops.logits_to_u8(two, "bytescale").cpu().numpy().ravel()
    assert np.array_equal(got, oc.png_payload(two.cpu().numpy()[0, 0]).ravel())
    assert got[0] == 0 and got[1] == 255
    neg = (-torch.rand(2, 1, 16, 16) * 50 - 1).cuda()                                   # all-negative logits, per-frame extrema
    got = ops.logits_to_u8(neg, "bytescale").cpu().numpy()
    for f in range(2):
        want = oc.png_payload(neg[f, 0].cpu().numpy())
        assert np.abs(got[f, 0].astype(int) - want.astype(int)).max() <= 1
        assert got[f, 0].max() == 255 and got[f, 0].min() == 0


@pytest.mark.parametrize("depth,count", [(2, 1), (2, 5), (3, 7), (4, 3)])
def test_sequence_segmenter_equals_per_frame_forward(depth, count):
    from osvos_pytorch_b200.inference import SequenceSegmenter
    net = _net()
    frames = [oc.synthetic_frame(1, 40, 56, 100 + i)[0].pin_memory() for i in range(count)]
    with torch.no_grad():
        want = [net(f.cuda())[-1].cpu() for f in frames]
    seg = SequenceSegmenter(net, output="logits", depth=depth)
    got = [r.clone() for r in seg(iter(frames))]
    assert len(got) == count
    for a, b in zip(got, want):
        assert torch.equal(a, b)                       # same kernels, same inputs -> identical, and in order
    # a second sequence through the same object (buffers reused)
    got2 = [r.clone() for r in seg(iter(frames[::-1]))]
    for a, b in zip(got2, want[::-1]):
        assert torch.equal(a, b)


def test_sequence_segmenter_u8_outputs_and_errors():
    from osvos_pytorch_b200.inference import SequenceSegmenter
    net = _net()
    frames = [oc.synthetic_frame(1, 33, 45, 7 + i)[0] for i in range(4)]
    with torch.no_grad():
        logits = [net(f.cuda())[-1].cpu() for f in frames]
    got = [r.clone() for r in SequenceSegmenter(net, output="bytescale")(frames)]
    for a, l in zip(got, logits):
        want = oc.png_payload(l[0, 0].numpy())
        assert a.dtype == torch.uint8 and a.shape == (1, 1, 33, 45)
        assert np.abs(a[0, 0].numpy().astype(int) - want.astype(int)).max() <= 1
    masks = [r.clone() for r in SequenceSegmenter(net, output="mask")(frames)]
    for a, l in zip(masks, logits):
        assert torch.equal(a > 0, l > 0)
    with pytest.raises(ValueError):
        SequenceSegmenter(net, output="jpeg")
    bad = [frames[0], oc.synthetic_frame(1, 40, 56, 1)[0]]
    with pytest.raises(ValueError):
        list(SequenceSegmenter(net)(bad))
