#!/bin/sh
# Installs the UNMODIFIED reference (nelpy/ghost, /root/reference) into baseline/_ref/ so that the CPU
# arm of bench.py (`--impl reference`, `cpu_baseline`) can time the reference's own
# ContinuousWaveletTransform.transform on the GPU box's host cores.
# baseline/_ref/ is git-ignored (nothing of the reference enters the history) but travels with gpurun.
# /root/reference is read-only and the build writes into the source tree, hence the copy under /tmp.
set -e
root=$(cd "$(dirname "$0")/.." && pwd)
src=${1:-/root/reference}
[ -d "$src" ] || { echo "no reference at $src"; exit 0; }
tmp=$(mktemp -d)
cp -r "$src" "$tmp/ref"
rm -rf "$root/baseline/_ref"
python -m pip install -q --no-index --no-build-isolation --no-deps --find-links /opt/wheelhouse \
    --target "$root/baseline/_ref" "$tmp/ref"
rm -rf "$tmp"
echo "installed reference into baseline/_ref"
