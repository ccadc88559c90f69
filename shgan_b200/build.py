"""Builds the C-ABI CUDA library `shgan_b200/lib/libshgan_b200.so` for sm_100a with nvcc.

    python -m shgan_b200.build [--force]

nvcc cross-compiles without a GPU.  The library is kept in-tree (git-ignored) so that it travels to
the GPU box with the repository snapshot.  One object per .cu file, rebuilt only when the source or
a header is newer than the object.
"""
import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, 'csrc')
LIBDIR = os.path.join(HERE, 'lib')
OBJDIR = os.path.join(LIBDIR, 'obj')
LIB = os.path.join(LIBDIR, 'libshgan_b200.so')
# test-only: the fp32 FMA cross-check convolution (csrc/check/*.cu).  Never loaded by the product path.
CHECK_LIB = os.path.join(LIBDIR, 'libshgan_b200_check.so')
CHECK_SRC = os.path.join(CSRC, 'check')
INCLUDE = os.path.join(os.path.dirname(HERE), 'include')

NVCC = os.environ.get('NVCC', '/usr/local/cuda/bin/nvcc')
FLAGS = ['-gencode', 'arch=compute_100a,code=sm_100a', '-lineinfo', '-O3', '-std=c++17',
         '-Xcompiler', '-fPIC', '--expt-relaxed-constexpr'] + os.environ.get('SHGAN_NVCC_FLAGS', '').split()


def _sources():
    return sorted(f for f in os.listdir(CSRC) if f.endswith('.cu'))


def _headers_mtime():
    hs = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith('.cuh')]
    hs += [os.path.join(INCLUDE, f) for f in os.listdir(INCLUDE) if f.endswith('.h')]
    return max(os.path.getmtime(h) for h in hs)


def _compile(src, verbose):
    obj = os.path.join(OBJDIR, os.path.basename(src)[:-3] + '.o')
    cmd = [NVCC] + FLAGS + ['-c', os.path.join(CSRC, src), '-o', obj]
    if verbose:
        print(' '.join(cmd), flush=True)
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError(f'nvcc failed on {src}:\n{r.stdout}\n{r.stderr}')
    return obj


def build(force=False, verbose=False):
    """Compile every CUDA source for sm_100a and link the shared library.  Returns the library path."""
    os.makedirs(OBJDIR, exist_ok=True)
    hm = _headers_mtime()
    todo, objs, check_objs = [], [], []
    check_sources = [os.path.join('check', f) for f in sorted(os.listdir(CHECK_SRC)) if f.endswith('.cu')]
    for src in _sources() + check_sources:
        obj = os.path.join(OBJDIR, os.path.basename(src)[:-3] + '.o')
        (check_objs if src in check_sources else objs).append(obj)
        sm = max(os.path.getmtime(os.path.join(CSRC, src)), hm)
        if force or not os.path.exists(obj) or os.path.getmtime(obj) < sm:
            todo.append(src)
    if todo:
        with ThreadPoolExecutor(max_workers=min(8, len(todo))) as ex:
            list(ex.map(lambda s: _compile(s, verbose), todo))
    if todo or not os.path.exists(LIB):
        cmd = [NVCC, '-shared', '-o', LIB] + objs + ['-gencode', 'arch=compute_100a,code=sm_100a', '-cudart', 'static']
        if verbose:
            print(' '.join(cmd), flush=True)
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError(f'link failed:\n{r.stdout}\n{r.stderr}')
    if todo or not os.path.exists(CHECK_LIB):
        cmd = [NVCC, '-shared', '-o', CHECK_LIB] + check_objs + [os.path.join(OBJDIR, 'api.o'), '-gencode',
                                                               'arch=compute_100a,code=sm_100a', '-cudart', 'static']
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError(f'link of the test-only check library failed:\n{r.stdout}\n{r.stderr}')
    return LIB


if __name__ == '__main__':
    print(build(force='--force' in sys.argv, verbose=True))
