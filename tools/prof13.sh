mkdir -p gpurun_out
ncu --set full --clock-control none --import-source on -k regex:"mode_dot_small|mode_gram_stream" -o gpurun_out/mode_dot_v3 python tools/prof_mode_dot.py > gpurun_out/prof_mode_dot.log 2>&1
tail -3 gpurun_out/prof_mode_dot.log
ncu -i gpurun_out/mode_dot_v3.ncu-rep --page details --csv > gpurun_out/mode_dot_v3_details.csv 2>/dev/null
ncu -i gpurun_out/mode_dot_v3.ncu-rep --page raw --csv > gpurun_out/mode_dot_v3_raw.csv 2>/dev/null
ls -la gpurun_out | tail -5
