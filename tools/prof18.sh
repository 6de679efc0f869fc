# full GPU suite + bench (no CPU baseline) + C5/C2 phase times after the triangular-fragment-skipping GEMM
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -3
python bench.py --steps 10 --warmup 3 --skip-cpu-baseline --skip-kron 2>&1 | python -c "
import sys,json
for l in sys.stdin:
    if l.startswith('{'):
        j=json.loads(l); print('ms/eval',j['ms_per_step'],'TF',j['roofline']['achieved'],'frac',j['roofline']['frac'],'launches',j['gpu_launches'],'nll',j['config']['nll'], 'kern', j['roofline']['kernel']['achieved']); print(j['batched']); print(j['clocks'])
    else: print(l.strip()[-300:])
"
python tools/phase_times.py --n 512 --d 8 --batch 1024
python tools/phase_times.py
