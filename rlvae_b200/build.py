"""In-tree build of librlvae_b200.so for sm_100a (explicit nvcc, no JIT cache).

The built library stays next to the sources (rlvae_b200/lib/) so that it travels
to the GPU box with the repo snapshot.  cudart is linked statically and libcuda is
never linked (TMA descriptors are encoded through cudaGetDriverEntryPoint), so the
library also loads on a machine without a GPU driver.
"""
from __future__ import annotations

import os
import shutil
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, 'csrc')
LIB_DIR = os.path.join(HERE, 'lib')
LIB = os.path.join(LIB_DIR, 'librlvae_b200.so')
SOURCES = ['rlvae_capi.cu', 'rlvae_direct.cu', 'rlvae_perpoint.cu', 'rlvae_hmc.cu', 'rlvae_tc.cu', 'rlvae_tc16.cu', 'rlvae_tc64.cu', 'rlvae_build.cu']
NVCC_FLAGS = ['-gencode', 'arch=compute_100a,code=sm_100a', '-O3', '-lineinfo', '-std=c++17',
              '-Xcompiler', '-fPIC', '-shared', '-cudart', 'static']


def _nvcc() -> str:
    for cand in (os.environ.get('NVCC'), shutil.which('nvcc'), '/usr/local/cuda/bin/nvcc'):
        if cand and os.path.isfile(cand):
            return cand
    raise RuntimeError('nvcc not found (set NVCC=/path/to/nvcc)')


def is_stale() -> bool:
    if not os.path.isfile(LIB):
        return True
    t = os.path.getmtime(LIB)
    deps = [os.path.join(CSRC, f) for f in os.listdir(CSRC)] + \
           [os.path.join(os.path.dirname(HERE), 'include', 'rlvae_b200.h')]
    return any(os.path.getmtime(d) > t for d in deps)


def build(force: bool = False, verbose: bool = False) -> str:
    """Compile every translation unit (in parallel: the tcgen05 kernels take ~1.5 minutes each) and link."""
    if not force and not is_stale():
        return LIB
    from concurrent.futures import ThreadPoolExecutor
    os.makedirs(LIB_DIR, exist_ok=True)
    obj_dir = os.path.join(LIB_DIR, 'obj')           # *.o is git-ignored
    os.makedirs(obj_dir, exist_ok=True)
    extra = os.environ.get('RLVAE_NVCC_EXTRA', '').split()     # e.g. -DRLVAE_TC_PROFILE (debugging aid)
    compile_flags = [f for f in NVCC_FLAGS if f not in ('-shared', '-cudart', 'static')]
    log = []

    def compile_one(src: str) -> str:
        obj = os.path.join(obj_dir, os.path.splitext(src)[0] + '.o')
        cmd = [_nvcc()] + compile_flags + extra + (['-Xptxas', '-v'] if verbose else []) + \
              ['-c', os.path.join(CSRC, src), '-o', obj]
        res = subprocess.run(cmd, capture_output=True, text=True)
        if res.returncode != 0:
            raise RuntimeError('nvcc failed:\n' + ' '.join(cmd) + '\n' + res.stdout + res.stderr)
        log.append(res.stderr)
        return obj

    with ThreadPoolExecutor(max_workers=min(len(SOURCES), os.cpu_count() or 1)) as pool:
        objs = list(pool.map(compile_one, SOURCES))
    cmd = [_nvcc()] + NVCC_FLAGS + ['-o', LIB + '.tmp'] + objs
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError('nvcc (link) failed:\n' + ' '.join(cmd) + '\n' + res.stdout + res.stderr)
    os.replace(LIB + '.tmp', LIB)
    if verbose:
        print('\n'.join(log))
    return LIB


if __name__ == '__main__':
    print(build(force=True, verbose=True))
