mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -5
python tools/bench_kron.py --json gpurun_out/kron_bench.json 2>&1 | grep -v "^ \|^$\|ncalls\|Ordered\|List red\|function calls\|torch tensordot\|transposed" | tee gpurun_out/kron_bench.txt
for cap in 0 120; do echo "FFGP_BG_CAP=$cap"; FFGP_BG_CAP=$cap python bench.py --steps 10 --warmup 3 --skip-cpu-baseline --skip-batched 2>&1 | python -c "
import sys,json
for l in sys.stdin:
    if l.startswith('{'):
        j=json.loads(l); print('ms/eval',j['ms_per_step'],'TF',j['roofline']['achieved'],'frac',j['roofline']['frac'],'launches',j['gpu_launches'],'nll',j['config']['nll'])
    else: print(l.strip()[-300:])
"; done
python tools/phase_times.py
