#!/bin/sh
# Copy the Python packages of the UNMODIFIED reference that the L4 binding test needs into baseline/_ref/ (git-ignored,
# NOT gpurun-ignored: it travels to the GPU box as scratch, it never enters the history).  /root/reference does not
# exist on the GPU box; with this copy tests/test_binding.py and tools/run_l4_on_gpu.py run the reference's own
# train_* loops there on the CUDA drop-ins.
set -e
SRC=${1:-/root/reference}
DST="$(dirname "$0")/../baseline/_ref"
rm -rf "$DST"; mkdir -p "$DST/Experiments"
for d in FidelityFusion_Models GaussianProcess MFGP_ver2023May MF_BayesianOptimization; do
  (cd "$SRC" && find "$d" -name '*.py' -print0 | cpio -0pdm --quiet "$OLDPWD/$DST") 2>/dev/null || \
  (cd "$SRC" && find "$d" -name '*.py' | while read f; do mkdir -p "$OLDPWD/$DST/$(dirname "$f")"; cp "$f" "$OLDPWD/$DST/$f"; done)
done
cp "$SRC/Experiments/log_debugger.py" "$DST/Experiments/"
echo "staged $(find "$DST" -name '*.py' | wc -l) files under $DST"
