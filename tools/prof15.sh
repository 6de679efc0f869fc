mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_dense.py -m gpu -q -x 2>&1 | tail -3
python bench.py --steps 10 --warmup 3 --skip-cpu-baseline --skip-batched 2>&1 | python -c "
import sys,json
for l in sys.stdin:
    if l.startswith('{'):
        j=json.loads(l); print('ms/eval',j['ms_per_step'],'TF',j['roofline']['achieved'],'frac',j['roofline']['frac'],'launches',j['gpu_launches'],'nll',j['config']['nll'])
    else: print(l.strip()[-300:])
"
python tools/phase_times.py
FFGP_TRACE=1 python tools/timeline_c2.py > gpurun_out/timeline_c2.txt 2>&1
