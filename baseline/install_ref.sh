#!/bin/bash
# Installs the UNMODIFIED reference (bgavran/autodiff) into baseline/_ref (git-ignored; travels to the GPU box with
# the gpurun snapshot).  /root/reference is read-only, so the wheel is built from a copy under /tmp.
# --no-deps: the only dependencies are numpy (present) and graphviz (absent; drawing only - baseline/ref_shim.py stubs it).
set -e
HERE="$(cd "$(dirname "$0")" && pwd)"
SRC="${1:-/root/reference}"
[ -d "$SRC" ] || { echo "no reference at $SRC"; exit 1; }
TMP="$(mktemp -d)"
cp -r "$SRC" "$TMP/ref"
rm -rf "$HERE/_ref"
python -m pip install -q --no-index --no-build-isolation --find-links /opt/wheelhouse --no-deps --target "$HERE/_ref" "$TMP/ref"
rm -rf "$TMP"
echo "installed reference into $HERE/_ref"
