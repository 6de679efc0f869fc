mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_kron.py -m gpu -q -x 2>&1 | tail -5
python tools/bench_kron.py --json gpurun_out/kron_bench.json 2>&1 | grep -v "^ \|^$\|ncalls\|Ordered\|List red\|function calls" | tee gpurun_out/kron_bench.txt
