#!/usr/bin/env python
"""Install the UNMODIFIED reference under baseline/_ref/ (git-ignored, travels to the GPU box with the snapshot) so that
`bench.py --impl reference` and the `cpu_baseline` leg can time the reference's own classes on the box's host cores.

1. the contract's recipe: pip install --no-index --no-build-isolation --find-links /opt/wheelhouse --target baseline/_ref
   <reference> - the reference has no setup.py / pyproject.toml, so pip refuses it ("neither 'setup.py' nor
   'pyproject.toml' found"); the outcome is recorded in baseline/_ref/INSTALL.txt and DESIGN.md;
2. fallback: the reference is pure Python, so "installing" it is placing its package directory on the path - the
   `autotune/` tree is copied verbatim (byte for byte; a SHA-256 over the tree is recorded) into baseline/_ref/.
Nothing under baseline/_ref is tracked, imported by the product, or edited.
"""
import hashlib
import os
import shutil
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.environ.get('AT2_REFERENCE', '/root/reference')
DST = os.path.join(ROOT, 'baseline', '_ref')


def tree_digest(path):
    h = hashlib.sha256()
    for d, dirs, files in sorted(os.walk(path)):
        dirs.sort()
        for f in sorted(files):
            if f.endswith('.pyc'):
                continue
            p = os.path.join(d, f)
            h.update(os.path.relpath(p, path).encode())
            with open(p, 'rb') as fh:
                h.update(fh.read())
    return h.hexdigest()


def install(verbose=True):
    if not os.path.isdir(os.path.join(REF, 'autotune')):
        if verbose:
            print(f'{REF} not present: keeping whatever baseline/_ref holds')
        return os.path.isdir(os.path.join(DST, 'autotune'))
    os.makedirs(DST, exist_ok=True)
    want = tree_digest(os.path.join(REF, 'autotune'))
    stamp = os.path.join(DST, 'INSTALL.txt')
    if os.path.isdir(os.path.join(DST, 'autotune')) and os.path.exists(stamp) and want in open(stamp).read():
        return True
    pip = subprocess.run([sys.executable, '-m', 'pip', 'install', '--no-index', '--no-build-isolation', '--find-links',
                          '/opt/wheelhouse', '--target', DST, REF], capture_output=True, text=True)
    note = [f'pip install --target baseline/_ref {REF}: rc={pip.returncode}',
            (pip.stderr.strip().splitlines() or ['(no stderr)'])[-1]]
    if pip.returncode != 0 or not os.path.isdir(os.path.join(DST, 'autotune')):
        shutil.rmtree(os.path.join(DST, 'autotune'), ignore_errors=True)
        shutil.copytree(os.path.join(REF, 'autotune'), os.path.join(DST, 'autotune'),
                        ignore=shutil.ignore_patterns('__pycache__', '*.pyc'))
        note.append('fallback: autotune/ copied verbatim (pure-Python package, nothing to build)')
    note.append(f'sha256(autotune tree) = {tree_digest(os.path.join(DST, "autotune"))} (reference: {want})')
    with open(stamp, 'w') as f:
        f.write('\n'.join(note) + '\n')
    if verbose:
        print('\n'.join(note))
    return True


if __name__ == '__main__':
    sys.exit(0 if install() else 1)
