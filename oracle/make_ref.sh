#!/bin/bash
# TEST INFRASTRUCTURE - builds oracle/_ref/: an UNMODIFIED copy of the reference's hot-path modules, made from the sources
# where they lie under /root/reference (nothing of it is committed: oracle/_ref/ is git-ignored, but NOT gpurun-ignored, so it
# travels to the GPU box like a built .so).  The reference is pure Python, so "compiling" it is copying these files:
#   networks/vgg_osvos.py   OSVOS module, make_layers_osvos, VGG loaders       
#   layers/osvos_layers.py  class_balanced_cross_entropy_loss, center_crop ... (a10, a12)
#   mypath.py, util/path_abstract.py  imported by networks/vgg_osvos.py
# Used ONLY by tests/, bench.py --impl reference / cpu_baseline / gpu_reference legs and __graft_entry__.smoke() as the
# checker and the timed baseline - never by the product package (tests/test_abi.py asserts that).
set -e
SRC=${1:-/root/reference}
DST="$(cd "$(dirname "$0")" && pwd)/_ref"
if [ ! -d "$SRC/networks" ]; then
  echo "make_ref: $SRC not present (GPU box): keeping the prebuilt $DST" >&2
  [ -f "$DST/networks/vgg_osvos.py" ]
  exit $?
fi
rm -rf "$DST"
mkdir -p "$DST/networks" "$DST/layers" "$DST/util"
for f in networks/__init__.py networks/vgg_osvos.py layers/__init__.py layers/osvos_layers.py mypath.py util/__init__.py util/path_abstract.py; do
  cp "$SRC/$f" "$DST/$f"
done
( cd "$SRC" && sha256sum networks/vgg_osvos.py layers/osvos_layers.py mypath.py util/path_abstract.py ) > "$DST/SHA256SUMS"
echo "make_ref: $DST built from $SRC"
