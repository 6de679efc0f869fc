mkdir -p gpurun_out
timeout 400 python -m pytest tests -m gpu -q -x 2>&1 | tail -3
for nb in 256 512; do FFGP_NB=$nb python bench.py --steps 10 --warmup 3 --skip-cpu-baseline 2>&1 | python -c "
import sys,json
for l in sys.stdin:
    if l.startswith('{'):
        j=json.loads(l); print('ms/eval',j['ms_per_step'],'TF',j['roofline']['achieved'],'frac',j['roofline']['frac'],'launches',j['gpu_launches'],'batched',j['batched']['value'], j['batched']['roofline_frac'])
    else: print(l.strip()[-300:])
"; done
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_c2.csv python tools/profile_c2.py --evals 2 > gpurun_out/prof_c2.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_c5.csv python tools/profile_c2.py --evals 2 --n 512 --d 8 --batch 1024 > gpurun_out/prof_c5.log 2>&1
