# ncu --set full of the 128x128 DMMA GEMM in the LAUUM and TRSM shapes at N=8192 (source-level stall reasons)
mkdir -p gpurun_out
ncu --set full --clock-control none --import-source on -k regex:gemm_dmma -s 1 -c 1 -o gpurun_out/gemm_lauum python tools/profile_gemm.py --which lauum --reps 1 > gpurun_out/prof_gemm_lauum.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:gemm_dmma -s 1 -c 1 -o gpurun_out/gemm_trsm python tools/profile_gemm.py --which trsm --reps 1 > gpurun_out/prof_gemm_trsm.log 2>&1
for w in lauum trsm; do
ncu -i gpurun_out/gemm_$w.ncu-rep --page raw --csv > gpurun_out/gemm_${w}_raw.csv 2>/dev/null
ncu -i gpurun_out/gemm_$w.ncu-rep --page source --csv > gpurun_out/gemm_${w}_sass.csv 2>/dev/null
done
python tools/profile_gemm.py --reps 5
ls -la gpurun_out
