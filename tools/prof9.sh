# per-kernel time + DRAM traffic + pipe utilisation for one C5 chunk (1024 x N=512) and one C2 eval; Kronecker-side HBM bench
mkdir -p gpurun_out
M=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active,sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active,sm__throughput.avg.pct_of_peak_sustained_elapsed
python tools/bench_kron.py --json gpurun_out/kron_bench.json 2>&1 | tee gpurun_out/kron_bench.txt
ncu --metrics $M --clock-control none --csv --log-file gpurun_out/metrics_c5.csv python tools/profile_c2.py --evals 2 --n 512 --d 8 --batch 1024 > gpurun_out/prof_c5.log 2>&1
tail -2 gpurun_out/prof_c5.log
ncu --metrics $M --clock-control none --csv --log-file gpurun_out/metrics_c2.csv python tools/profile_c2.py --evals 2 > gpurun_out/prof_c2.log 2>&1
tail -2 gpurun_out/prof_c2.log
python tools/summarize_metrics.py gpurun_out/metrics_c5.csv --agg | head -30
