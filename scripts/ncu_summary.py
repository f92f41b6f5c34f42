"""Turn the captures of scripts/profile_round.sh (gpurun_out/<tag>_*) into the tracked summaries under profiles/:
<tag>_kernels_ncu_summary.md (launch list shares + per-kernel ncu metrics + stall reasons), <tag>_launches_bench.csv and
<tag>_dram_traffic.json (DRAM bytes per launch of each captured kernel, read by bench.py for roofline.traffic).
usage: python scripts/ncu_summary.py r01"""
import csv
import io
import json
import os
import shutil
import subprocess
import sys
from collections import OrderedDict

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
tag = sys.argv[1] if len(sys.argv) > 1 else 'r01'
out, prof = os.path.join(ROOT, 'gpurun_out'), os.path.join(ROOT, 'profiles')
METRICS = ['gpu__time_duration.sum', 'smsp__inst_executed.sum', 'smsp__thread_inst_executed.sum',
           'smsp__issue_active.avg.pct_of_peak_sustained_active', 'smsp__thread_inst_executed_per_inst_executed.ratio',
           'smsp__warps_active.avg.per_cycle_active', 'launch__registers_per_thread', 'launch__grid_size', 'launch__block_size',
           'launch__shared_mem_per_block_dynamic', 'sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active',
           'sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active',
           'sm__inst_executed_pipe_fmaheavy.avg.pct_of_peak_sustained_active',
           'sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active',
           'sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active',
           'sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active',
           'smsp__sass_thread_inst_executed_op_fadd_pred_on.sum', 'smsp__sass_thread_inst_executed_op_fmul_pred_on.sum',
           'smsp__sass_thread_inst_executed_op_ffma_pred_on.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum']


def short(name):
    return name.replace('void ', '').replace('<unnamed>::', '').split('(')[0]


def raw_rows(rep):
    txt = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True, cwd='/tmp').stdout
    rows = list(csv.reader(io.StringIO(txt)))
    return rows[0], rows[1], rows[2:]


lines = [f'# Round {tag} - ncu summaries of the kernels of `python bench.py --steps 6 --warmup 3 --no-cpu-baseline`', '',
         '`--set full --clock-control none --import-source on`, one B200, captured after the same command had exited 0 '
         'without ncu (recipe: `scripts/profile_round.sh`, summary: `scripts/ncu_summary.py`).  Per-launch values; launch '
         f'list of the same command: `{tag}_launches_bench.csv`.', '']
# ---- launch list ----
lcsv = os.path.join(out, f'{tag}_launches.csv')
if os.path.exists(lcsv):
    shutil.copy(lcsv, os.path.join(prof, f'{tag}_launches_bench.csv'))
    agg = OrderedDict()
    with open(lcsv) as f:
        rows = [r for r in csv.reader(f) if len(r) > 14 and r[0].isdigit()]
    for r in rows:
        k = short(r[4])[:70]
        a = agg.setdefault(k, [0, 0.0])
        a[0] += 1
        a[1] += float(r[14]) / 1e6
    tot = sum(v[1] for v in agg.values())
    lines += ['## Launch list (gpu__time_duration, cold-cache, serialised: shares only)', '',
              '| kernel | launches | total ms | avg us | share |', '|---|---|---|---|---|']
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        lines.append(f'| `{k}` | {v[0]} | {v[1]:.3f} | {v[1] / v[0] * 1e3:.1f} | {100 * v[1] / tot:.1f}% |')
    lines.append('')
# ---- full captures ----
traffic = {}
for rep in (f'{tag}_prof_bench.ncu-rep', f'{tag}_prof_hb_flat.ncu-rep', f'{tag}_prof_lazy.ncu-rep', f'{tag}_prof_knownfn.ncu-rep', f'{tag}_prof_nnlookup.ncu-rep',
            f'{tag}_prof_tpe.ncu-rep'):
    path = os.path.join(out, rep)
    if not os.path.exists(path):
        continue
    hdr, units, rows = raw_rows(path)
    by_kernel = OrderedDict()
    for r in rows:
        by_kernel.setdefault(short(r[hdr.index('Kernel Name')]), []).append(r)
    for k, rs in by_kernel.items():
        r = rs[-1]
        lines += [f'## `{k}`  ({len(rs)} launches captured, last one shown)', '', '| metric | unit | value |', '|---|---|---|']
        for m in METRICS:
            if m in hdr:
                lines.append(f'| `{m}` | {units[hdr.index(m)]} | {r[hdr.index(m)]} |')
        stalls = [(float(r[i] or 0), h.split('issue_stalled_')[1].split('_per')[0]) for i, h in enumerate(hdr)
                  if h.startswith('smsp__average_warp') and 'issue_stalled' in h and h.endswith('.ratio') and 'not_issued' not in h]
        tot_s = sum(v for v, _ in stalls) or 1.0
        lines += ['', 'Warp-stall reasons (share of sampled warp latency per issued instruction): ' +
                  ', '.join(f'{n} {100 * v / tot_s:.1f}%' for v, n in sorted(stalls, reverse=True)[:7]), '']
        rd, wr = hdr.index('dram__bytes_read.sum'), hdr.index('dram__bytes_write.sum')

        def tobytes(x, unit):
            return float(x) * {'byte': 1, 'Kbyte': 1e3, 'Mbyte': 1e6, 'Gbyte': 1e9}.get(unit, 1)
        traffic[k] = {'dram_bytes_per_launch_avg': sum(tobytes(q[rd], units[rd]) + tobytes(q[wr], units[wr]) for q in rs) / len(rs),
                      'launches_captured': len(rs)}
# ---- SASS-counted FP32 work of one simulation launch (separate --metrics pass) ----
scsv = os.path.join(out, f'{tag}_sass_flops.csv')
if os.path.exists(scsv):
    vals = {}
    with open(scsv) as f:
        for r in csv.reader(f):
            if len(r) > 14 and r[0].isdigit():
                vals[r[12]] = float(r[14])
                kname = short(r[4])
    fadd, fmul, ffma = (vals.get(f'smsp__sass_thread_inst_executed_op_{k}_pred_on.sum', 0.0) for k in ('fadd', 'fmul', 'ffma'))
    flop = fadd + fmul + 2 * ffma
    lines += [f'## SASS-counted FP32 work of one `{kname}` launch', '', '| metric | value |', '|---|---|'] + \
             [f'| `{k}` | {v:.0f} |' for k, v in vals.items()] + \
             [f'| FP32 flop (fadd + fmul + 2 ffma) | {flop:.0f} |', '']
    traffic.setdefault(kname, {})['sass_fp32_flop_per_launch'] = flop
    traffic[kname]['thread_inst_per_launch'] = vals.get('smsp__thread_inst_executed.sum')
traffic['_source'] = f'ncu --set full --clock-control none, profiles/{tag}_kernels_ncu_summary.md'
with open(os.path.join(prof, f'{tag}_kernels_ncu_summary.md'), 'w') as f:
    f.write('\n'.join(lines) + '\n')
with open(os.path.join(prof, f'{tag}_dram_traffic.json'), 'w') as f:
    json.dump(traffic, f, indent=1)
print('\n'.join(lines))
