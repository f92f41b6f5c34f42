"""Development aid: per-source-line instruction / stall-sample shares from an ncu report (--import-source on).
usage: python scripts/ncu_lines.py report.ncu-rep [top_n]"""
import csv
import io
import subprocess
import sys

rep, top = __import__('os').path.abspath(sys.argv[1]), int(sys.argv[2]) if len(sys.argv) > 2 else 40
raw = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True, cwd='/tmp').stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr = rows[0]
want = ['Kernel Name', 'gpu__time_duration.sum', 'smsp__inst_executed.sum', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
        'smsp__thread_inst_executed_per_inst_executed.ratio', 'smsp__warps_active.avg.per_cycle_active',
        'launch__registers_per_thread', 'launch__grid_size', 'launch__block_size', 'launch__shared_mem_per_block_dynamic',
        'sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_fmaheavy.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active',
        'dram__bytes_read.sum', 'dram__bytes_write.sum']
for r in rows[2:]:
    for w in want:
        if w in hdr:
            print(f'{w}: {r[hdr.index(w)]} {rows[1][hdr.index(w)]}')
    stalls = [(float(r[i] or 0), h) for i, h in enumerate(hdr)
              if h.startswith('smsp__average_warp') and 'issue_stalled' in h and h.endswith('.ratio') and 'not_issued' not in h]
    for v, h in sorted(stalls, reverse=True)[:7]:
        print(f'   stall {h.split("issue_stalled_")[1].split("_per")[0]}: {v:.2f}')
src = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv', '--print-source', 'cuda,sass'], capture_output=True,
                     text=True, cwd='/tmp').stdout
cur, agg, total = None, {}, 0
for r in csv.reader(io.StringIO(src)):
    if len(r) >= 2 and r[0] == 'File Path':
        cur = r[1].split('/')[-1]
        continue
    if len(r) < 8 or not r[0].isdigit():
        continue
    try:
        inst, samp = int(r[7]), int(r[6])
    except ValueError:
        continue
    k = (cur, int(r[0]))
    a = agg.get(k, (0, 0, r[1].strip()[:100]))
    agg[k] = (a[0] + inst, a[1] + samp, a[2])
    total += inst
tot_s = sum(v[1] for v in agg.values()) or 1
print('total warp-inst', total, 'samples', tot_s)
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1])[:top]:
    print(f'{k[0]}:{k[1]:4d} inst {100 * v[0] / max(total, 1):5.1f}% samp {100 * v[1] / tot_s:5.1f}%  {v[2]}')
if len(sys.argv) > 3:      # line ranges "file:lo-hi,..." -> share per region
    for spec in sys.argv[3].split(','):
        f, rng = spec.split(':')
        lo, hi = (int(x) for x in rng.split('-'))
        i = sum(v[0] for k, v in agg.items() if k[0] == f and lo <= k[1] <= hi)
        sm = sum(v[1] for k, v in agg.items() if k[0] == f and lo <= k[1] <= hi)
        print(f'region {spec}: inst {100 * i / max(total, 1):.1f}% samp {100 * sm / tot_s:.1f}%')
