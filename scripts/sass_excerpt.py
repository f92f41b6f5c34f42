"""Write the SASS of the simulation kernel's attempt loops (lean variant: hb_sim_kernel<float, PHILOX, no smoothing>) from
the built library to profiles/<tag>_hb_sim_attempt_loop.sass, with an opcode histogram per loop copy.
usage: python scripts/sass_excerpt.py r02"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, 'autotune-v2_b200', 'lib', 'libat2_b200.so')
tag = sys.argv[1] if len(sys.argv) > 1 else 'r02'
elf = subprocess.run(['cuobjdump', '-elf', LIB], capture_output=True, text=True).stdout
sym = sorted(set(re.findall(r'_Z\w*hb_sim_kernelIfLi0ELb0ELi1ELb0EEEvNS_6ParamsIT_EE\b', elf)))[0]
sass = subprocess.run(['cuobjdump', '-sass', '-fun', sym, LIB], capture_output=True, text=True).stdout
ins = []
for line in sass.splitlines():
    m = re.match(r'\s+/\*([0-9a-f]{4,5})\*/\s+(.*?);\s*/\*', line)
    if m:
        ins.append((int(m.group(1), 16), m.group(2).strip()))
# the two innermost trajectory loops: backward branches spanning 120-260 instructions
loops = []
for a, t in ins:
    m = re.search(r'\bBRA(?:\.U)?\s+(?:!?U?P\d,\s*)?0x([0-9a-f]+)', t)
    if m and 120 <= (a - int(m.group(1), 16)) // 16 + 1 <= 260:
        loops.append((int(m.group(1), 16), a))
out = [f'SASS of the trajectory loops of {sym}',
       '(cuobjdump -sass of autotune-v2_b200/lib/libat2_b200.so, sm_100a; scripts/sass_excerpt.py).  One trip = one',
       'Philox4x32-10 block (the NEXT trip\'s, issued ahead) + two Box-Muller pairs -> four Marsaglia-Tsang attempts + recurrence',
       'steps (csrc/sim_core.cuh philox_trajectory).  Loop 1 captures checkpoints (predicated STS + row-pointer bump); loop 2 is',
       'the capture-free copy that runs after the second-to-last checkpoint.', '']
for k, (lo, hi) in enumerate(loops[:2]):
    body = [(a, t) for a, t in ins if lo <= a <= hi]
    hist = collections.Counter(re.sub(r'^@!?U?P\d+\s+', '', t).split()[0].split('.')[0] for _, t in body)
    out.append(f'== loop {k + 1}: 0x{lo:04x}..0x{hi:04x}, {len(body)} instructions per 4 attempts = {len(body) / 4:.1f} per attempt')
    out.append('   ' + ', '.join(f'{op} {n}' for op, n in hist.most_common()))
    pipes = {'fma (FFMA FMUL FADD IMAD)': sum(hist[o] for o in ('FFMA', 'FMUL', 'FADD', 'IMAD')),
             'alu (LOP3 I2FP PRMT FSETP FSEL SEL VIADD ISETP IADD3 SHF)': sum(hist[o] for o in ('LOP3', 'I2FP', 'PRMT', 'FSETP', 'FSEL', 'SEL', 'VIADD', 'ISETP', 'IADD3', 'SHF')),
             'xu (MUFU)': hist['MUFU'], 'lsu (LDS STS)': hist['LDS'] + hist['STS']}
    out.append('   per pipe: ' + ', '.join(f'{p} {n}' for p, n in pipes.items()))
    out.append('')
    out += [f'        /*{a:04x}*/  {t} ;' for a, t in body]
    out.append('')
path = os.path.join(ROOT, 'profiles', f'{tag}_hb_sim_attempt_loop.sass')
with open(path, 'w') as f:
    f.write('\n'.join(out) + '\n')
print('\n'.join(out[:6] + [l for l in out if l.startswith('==') or l.startswith('   ')]))
