"""TEST INFRASTRUCTURE - loads the UNMODIFIED reference modules from oracle/_ref/ (built by oracle/make_ref.sh).

The reference imports its own pieces by top-level names (`from layers.osvos_layers import ...`, `from mypath import Path`:
/root/reference/networks/vgg_osvos.py:12-13) and this repo ships drop-in shims under the very same names (`networks/`,
`layers/`, `mypath.py`), so the copy cannot simply be put on sys.path.  `load()` imports it with the colliding
`sys.modules` entries set aside and puts them back afterwards; the returned module objects keep working because the
reference binds what it needs at import time.

Only tests/, bench.py's reference / cpu_baseline / gpu_reference legs and __graft_entry__.smoke() may call this.
"""
import contextlib
import importlib
import io
import os
import sys
from types import SimpleNamespace

REF_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "_ref")
_NAMES = ("networks", "networks.vgg_osvos", "layers", "layers.osvos_layers", "mypath", "util", "util.path_abstract")
_cache = None


def available() -> bool:
    return os.path.exists(os.path.join(REF_DIR, "networks", "vgg_osvos.py"))


def load():
    """-> namespace(net=<reference networks.vgg_osvos>, layers=<reference layers.osvos_layers>)."""
    global _cache
    if _cache is not None:
        return _cache
    if not available():
        raise FileNotFoundError(f"{REF_DIR} missing: run `bash oracle/make_ref.sh` in the build container")
    saved = {n: sys.modules.pop(n) for n in _NAMES if n in sys.modules}
    sys.path.insert(0, REF_DIR)
    try:
        importlib.invalidate_caches()
        with contextlib.redirect_stdout(io.StringIO()):
            net = importlib.import_module("networks.vgg_osvos")
            lay = importlib.import_module("layers.osvos_layers")
        assert os.path.abspath(net.__file__).startswith(REF_DIR), net.__file__
        assert os.path.abspath(lay.__file__).startswith(REF_DIR), lay.__file__
    finally:
        sys.path.remove(REF_DIR)
        for n in _NAMES:
            sys.modules.pop(n, None)
        sys.modules.update(saved)
        importlib.invalidate_caches()
    _cache = SimpleNamespace(net=net, layers=lay)
    return _cache


def build_reference(params, device="cpu"):
    """The reference's own `OSVOS(pretrained=0)` with the oracle's seeded weights loaded (ConvT weights stay the
    reference's `interp_surgery` output)."""
    ref = load()
    with contextlib.redirect_stdout(io.StringIO()):
        net = ref.net.OSVOS(pretrained=0)
    sd = net.state_dict()
    for k, v in params.items():
        assert sd[k].shape == v.shape, k
        sd[k] = v.clone()
    net.load_state_dict(sd)
    return net.to(device)
