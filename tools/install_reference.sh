#!/bin/bash
# The reference has no setup.py / pyproject (it is a directory of scripts), so `pip install --target baseline/_ref`
# has nothing to install: this copies the UNMODIFIED hot-path sources where the reference arm of bench.py and
# tools/psnr_check.py import them from.  baseline/_ref/ is git-ignored (never in history) and NOT gpurun-ignored
# (it travels to the GPU box, which has no /root/reference).
set -e
SRC=${1:-/root/reference}
DST="$(cd "$(dirname "$0")/.." && pwd)/baseline/_ref"
[ -f "$SRC/nerf-ours/render.py" ] || { echo "no reference tree at $SRC"; exit 0; }
mkdir -p "$DST/nerf-ours" "$DST/nerf++-ours"
cp "$SRC"/nerf-ours/*.py "$DST/nerf-ours/"
cp -r "$SRC"/nerf-ours/configs "$DST/nerf-ours/"
cp "$SRC"/nerf++-ours/*.py "$DST/nerf++-ours/"
cp -r "$SRC"/nerf++-ours/configs "$DST/nerf++-ours/" 2>/dev/null || true
(cd "$SRC" && find nerf-ours nerf++-ours -maxdepth 1 -name '*.py' | sort | xargs sha256sum) > "$DST/SHA256SUMS"
echo "reference sources copied to $DST ($(wc -l < "$DST/SHA256SUMS") files)"
