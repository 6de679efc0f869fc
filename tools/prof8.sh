# restored-state check: GPU tests, the default bench (with cpu baseline), reference arm, launch list of the bench command, phases
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q -x 2>&1 | tail -3
python bench.py 2>gpurun_out/bench_err.log | tee gpurun_out/bench_line.json | cut -c1-1500
tail -3 gpurun_out/bench_err.log
python tools/phase_times.py | tee gpurun_out/phase_c2.txt
python tools/phase_times.py --n 512 --d 8 --batch 1024 | tee gpurun_out/phase_c5.txt
python tools/profile_gemm.py --reps 5 | tee gpurun_out/gemm_shapes.txt
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_bench.csv python bench.py --steps 2 --warmup 3 --skip-batched --skip-cpu-baseline > gpurun_out/bench_ncu.log 2>&1
python tools/summarize_launches.py gpurun_out/launches_bench.csv --last-frac 0.2 | head -14
