mkdir -p gpurun_out
timeout 400 python -m pytest tests -m gpu -q 2>&1 | tail -5
for nb in 128 256 512; do echo "FFGP_NB=$nb"; FFGP_NB=$nb python bench.py --steps 10 --warmup 3 --skip-cpu-baseline 2>&1 | python -c "
import sys,json
for l in sys.stdin:
    if l.startswith('{'):
        j=json.loads(l); print('ms/eval',j['ms_per_step'],'TF',j['roofline']['achieved'],'frac',j['roofline']['frac'],'launches',j['gpu_launches'],'batched GPs/s',j['batched']['value'],'nll',j['config']['nll'])
    else: print(l.strip()[-300:])
"; done
FFGP_NB=256 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_c2.csv python tools/profile_c2.py --evals 2 > gpurun_out/prof_c2.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:potrf_trtri_base -s 4 -c 1 -o gpurun_out/base_kernel python tools/profile_c2.py --evals 1 --n 2048 > gpurun_out/prof_base.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_c5.csv python tools/profile_c2.py --evals 2 --n 512 --d 8 --batch 1024 > gpurun_out/prof_c5.log 2>&1
ls gpurun_out
