"""Brief of one `ncu --set full` report: duration, pipe utilisation, occupancy limits, stall samples, DRAM bytes and the
hottest source lines.   python tools/ncu_brief.py report.ncu-rep [kernel-substring]"""
import csv, subprocess, sys
rep = sys.argv[1]
raw = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
h, v = rows[0], rows[-1]
keys = ['gpu__time_duration.sum', 'sm__throughput.avg.pct', 'sm__inst_executed_pipe_fp64.avg.pct', 'sm__warps_active.avg.pct',
        'launch__occupancy_limit', 'launch__registers_per_thread', 'smsp__pcsamp_warps_issue_stalled', 'dram__bytes_read.sum',
        'dram__bytes_write.sum', 'smsp__issue_active.avg.pct', 'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum',
        'dram__throughput.avg.pct', 'lts__t_sector_hit_rate.pct', 'sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_fp64_op_dmma', 'launch__grid_size', 'launch__shared_mem_per_block_dynamic']
for i, k in enumerate(h):
    if any(k.startswith(x) for x in keys) and '_not_issued' not in k:
        print(f'{k:78s} {v[i]}')
src = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv', '--print-source', 'cuda'], capture_output=True, text=True).stdout
try:
    rows = list(csv.reader(src.splitlines()))
    hi = [i for i, r in enumerate(rows) if 'Source' in r and any('Sampl' in c for c in r)][0]
    hdr = rows[hi]
    si = hdr.index('Source'); ci = [i for i, c in enumerate(hdr) if c.startswith('# Samples') or c == 'Warp Stall Sampling (All Samples)'][0]
    body = [r for r in rows[hi + 1:] if len(r) == len(hdr)]
    tot = sum(float(r[ci] or 0) for r in body) or 1
    print('--- hottest source lines (% of stall samples) ---')
    for r in sorted(body, key=lambda r: -float(r[ci] or 0))[:14]:
        print(f'{100 * float(r[ci] or 0) / tot:5.1f}%  {r[si].strip()[:130]}')
except Exception as e:
    print('source page unavailable:', e)
