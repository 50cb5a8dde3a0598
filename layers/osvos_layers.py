"""Import-path shim for ``from layers.osvos_layers import class_balanced_cross_entropy_loss`` etc.
(reference train_online.py:22, train_parent.py:21)."""
from osvos_pytorch_b200.layers.osvos_layers import (  # noqa: F401
    center_crop, class_balanced_cross_entropy_loss, interp_surgery, logit, sigmoid_np, upsample_filt)
