#!/bin/sh
# Copy the Python packages of the UNMODIFIED reference that the L4 binding test and the reference bench arm need into
# baseline/_ref/ (git-ignored, NOT gpurun-ignored: it travels to the GPU box as scratch, it never enters the history).
# /root/reference does not exist on the GPU box; with this copy tests/test_binding.py and tools/run_l4_on_gpu.py run the
# reference's own train_* loops there on the CUDA drop-ins, and bench.py --impl reference times the reference's own cigp.
set -e
SRC=${1:-/root/reference}
ROOT=$(cd "$(dirname "$0")/.." && pwd)
DST="$ROOT/baseline/_ref"
rm -rf "$DST"
mkdir -p "$DST/Experiments"
cd "$SRC"
for d in FidelityFusion_Models GaussianProcess MFGP_ver2023May MF_BayesianOptimization Bayesian_optimization; do
  find "$d" -name '*.py' | while read -r f; do
    mkdir -p "$DST/$(dirname "$f")"
    cp "$f" "$DST/$f"
  done
done
cp "$SRC/Experiments/log_debugger.py" "$DST/Experiments/"
echo "staged $(find "$DST" -name '*.py' | wc -l) files under $DST"
