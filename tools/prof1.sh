mkdir -p gpurun_out
timeout 300 python -m pytest tests -m gpu -q 2>&1 | tail -5
python tools/profile_gemm.py --n 8192 --reps 3 | tee gpurun_out/gemm_shapes.txt
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_c2.csv python tools/profile_c2.py --evals 2 > gpurun_out/prof_c2.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:gemm_dmma -c 1 -o gpurun_out/gemm_lauum python tools/profile_gemm.py --n 8192 --reps 1 --which lauum > gpurun_out/prof_gemm.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_c5.csv python tools/profile_c2.py --evals 2 --n 512 --d 8 --batch 1024 > gpurun_out/prof_c5.log 2>&1
ls -la gpurun_out
