#!/usr/bin/env python
"""Stage the UNMODIFIED reference under the git-ignored baseline/_ref/relightable-nr/ so that it travels to the GPU box
(gpurun ships the work tree minus .git / .gpurunignore; /root/reference itself does not exist there).

    python tools/stage_reference.py            # build container only (needs /root/reference)

What is staged: the reference's Python sources (scripts + modules, byte for byte) and its one data fixture
(sphere_samples_4096.mat).  Nothing is copied into tracked paths.  Used by
  * tests/test_scripts_gpu.py      -- runs the unchanged scripts through `python -m relightable_nr_b200.run`,
  * bench.py --impl reference      -- times the reference's own PyTorch modules on the host cores (cpu_baseline.kind "reference"),
  * tests/golden/ref_import.py     -- imports the real modules as the checker when /root/reference is absent.
`pip install --target baseline/_ref /root/reference` does not apply: the reference is a directory of scripts with no
installable package (its only setup.py builds the torch-1.1 CUDA extension, which does not compile against torch 2.11).
"""
import os
import shutil
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = '/root/reference'
DST = os.path.join(ROOT, 'baseline', '_ref', 'relightable-nr')
KEEP_EXT = ('.py', '.mat', '.sh', '.md', '.yml', '.txt')


def stage(verbose=True):
    if not os.path.isdir(SRC):
        if verbose:
            print('stage_reference: %s absent (GPU box) -- using whatever is already staged under %s' % (SRC, DST))
        return os.path.isdir(DST)
    n = 0
    for dirpath, dirnames, filenames in os.walk(SRC):
        dirnames[:] = [d for d in dirnames if d not in ('.git', '__pycache__', 'build')]
        rel = os.path.relpath(dirpath, SRC)
        for fn in filenames:
            if not fn.endswith(KEEP_EXT) and fn != 'LICENSE':
                continue
            dst_dir = os.path.join(DST, rel) if rel != '.' else DST
            os.makedirs(dst_dir, exist_ok=True)
            s, d = os.path.join(dirpath, fn), os.path.join(dst_dir, fn)
            if not os.path.exists(d) or os.path.getmtime(s) > os.path.getmtime(d) or os.path.getsize(s) != os.path.getsize(d):
                shutil.copy2(s, d)
            n += 1
    if verbose:
        print('stage_reference: %d files staged under %s' % (n, DST))
    return True


if __name__ == '__main__':
    sys.exit(0 if stage() else 1)
