mkdir -p gpurun_out
ncu --set full --clock-control none --import-source on -k regex:potrf_trtri_base -s 6 -c 1 -o gpurun_out/base_c2 python tools/profile_c2.py --evals 1 --n 2048 > gpurun_out/prof_base.log 2>&1
ncu -i gpurun_out/base_c2.ncu-rep --page source --csv > gpurun_out/base_c2_sass.csv 2>/dev/null
ncu -i gpurun_out/base_c2.ncu-rep --page raw --csv > gpurun_out/base_c2_raw.csv 2>/dev/null
ls -la gpurun_out | tail -5
