#!/usr/bin/env python
"""Stage the reference's UNMODIFIED caller files where the GPU box can see them.

``tests/test_reference_callers_gpu.py`` drives the reference's own ``main.py`` / ``algorithm.py`` /
``eval_policy.py`` against this repo's ``gym_AO`` on a B200.  ``/root/reference`` does not exist on the GPU box, so
the caller files are copied -- byte for byte -- into ``baseline/_ref/callers/`` (git-ignored: reference sources
never enter this repository's history; not gpurun-ignored: the directory travels with the snapshot).  The SHA-256
of every file is checked by the test against ``tests/golden/reference_callers.sha256`` (hashes only).

    python tools/stage_reference_callers.py [--reference /root/reference] [--write-manifest]
"""
import argparse
import hashlib
import os
import shutil

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
FILES = ['main.py', 'algorithm.py', 'network.py', 'replay_buffer.py', 'eval_policy.py', 'arguments.py']


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--reference', default='/root/reference')
    ap.add_argument('--write-manifest', action='store_true')
    a = ap.parse_args()
    dst = os.path.join(ROOT, 'baseline', '_ref', 'callers')
    os.makedirs(dst, exist_ok=True)
    lines = []
    for f in FILES:
        src = os.path.join(a.reference, f)
        shutil.copyfile(src, os.path.join(dst, f))
        lines.append(f'{hashlib.sha256(open(src, "rb").read()).hexdigest()}  {f}')
    if a.write_manifest:
        with open(os.path.join(ROOT, 'tests', 'golden', 'reference_callers.sha256'), 'w') as fh:
            fh.write('\n'.join(lines) + '\n')
    print('\n'.join(lines))


if __name__ == '__main__':
    main()
