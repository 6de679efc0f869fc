mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -15
python tools/bench_kron.py --json gpurun_out/kron_bench.json 2>&1 | tee gpurun_out/kron_bench.txt
