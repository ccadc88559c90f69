"""Copies the UNMODIFIED reference tree into the git-ignored `baseline/_ref/` so that it travels to the GPU box
with the repository snapshot (gpurun ships the working tree, not /root/reference).

    python baseline/install_reference.py [--src /root/reference]

What is copied: `lib/` (the model zoo and the modules it imports; pure Python + the upfirdn2d plugin sources,
lib/model_zoo/stylegan_utils/upfirdn2d.{cpp,cu,h}) and `configs/`.  Nothing is edited.  `baseline/_ref/` is listed in
.gitignore: no reference source enters the history.  It is used only by
  * `bench.py --impl reference` (the reference arm: the reference's own generator, CPU or `--ref-device cuda`),
  * the `-m gpu` tests that compare the CUDA path with the reference running on the same GPU
    (tests/test_gpu_reference.py; skipped when the tree is absent),
through `tests/golden/ref_import.py`, which stubs the two unused third-party imports (matplotlib, easydict).
The reference has no setup.py / pyproject, so `pip install /root/reference` is not applicable (recorded in DESIGN.md).
"""
import argparse
import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
DST = os.path.join(HERE, '_ref')


def install(src='/root/reference', quiet=False):
    if not os.path.isdir(os.path.join(src, 'lib', 'model_zoo')):
        if not quiet:
            print(f'reference tree not found at {src}; keeping whatever is in {DST}')
        return os.path.isdir(os.path.join(DST, 'lib', 'model_zoo'))
    for sub in ('lib', 'configs'):
        d = os.path.join(DST, sub)
        if os.path.isdir(d):
            shutil.rmtree(d)
        shutil.copytree(os.path.join(src, sub), d, ignore=shutil.ignore_patterns('__pycache__', '*.pyc'))
    n = sum(len(fs) for _, _, fs in os.walk(DST))
    if not quiet:
        print(f'installed {n} reference files into {DST}')
    return True


if __name__ == '__main__':
    ap = argparse.ArgumentParser()
    ap.add_argument('--src', default='/root/reference')
    args = ap.parse_args()
    sys.exit(0 if install(args.src) else 1)
