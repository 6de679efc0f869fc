mkdir -p gpurun_out
timeout 400 python -m pytest tests -m gpu -q -x 2>&1 | tail -3
python bench.py --steps 10 --warmup 3 --skip-cpu-baseline 2>&1 | python -c "
import sys,json
for l in sys.stdin:
    if l.startswith('{'):
        j=json.loads(l); print('ms/eval',j['ms_per_step'],'TF',j['roofline']['achieved'],'frac',j['roofline']['frac'],'launches',j['gpu_launches'],'batched',j['batched'])
    else: print(l.strip()[-300:])
"
ncu --set full --clock-control none --import-source on -k regex:potrf_trtri_base -s 2 -c 1 -o gpurun_out/base_kernel_b python tools/profile_c2.py --evals 1 --n 512 --d 8 --batch 296 > gpurun_out/prof_base.log 2>&1
ncu -i gpurun_out/base_kernel_b.ncu-rep --page source --print-source cuda --csv > gpurun_out/base_src_cuda.csv 2>gpurun_out/base_src_err.txt
ncu -i gpurun_out/base_kernel_b.ncu-rep --page source --csv > gpurun_out/base_src_sass.csv 2>>gpurun_out/base_src_err.txt
ls -la gpurun_out | head -20
