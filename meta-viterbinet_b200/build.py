"""Builds the C-ABI CUDA library in-tree with nvcc for sm_100a (cross-compiles without a GPU).

    python meta-viterbinet_b200/build.py [--force]

The resulting ``libmvn_b200.so`` sits next to this file; it is git-ignored but travels to the
GPU box with the gpurun snapshot.
"""
import hashlib
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, 'csrc')
LIB = os.path.join(HERE, 'libmvn_b200.so')
STAMP = os.path.join(HERE, '.libmvn_b200.stamp')
NVCC_FLAGS = ['-gencode', 'arch=compute_100a,code=sm_100a', '-O3', '-lineinfo', '-std=c++17',
              '--shared', '-Xcompiler', '-fPIC', '-Xptxas', '-v']


def _sources():
    return sorted(os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith('.cu'))


def _digest():
    h = hashlib.sha256()
    inc = os.path.join(os.path.dirname(HERE), 'include')
    files = [os.path.join(CSRC, f) for f in sorted(os.listdir(CSRC))]
    files += [os.path.join(inc, f) for f in sorted(os.listdir(inc))]
    for f in files:
        h.update(f.encode())
        with open(f, 'rb') as fh:
            h.update(fh.read())
    h.update(' '.join(NVCC_FLAGS).encode())
    return h.hexdigest()


def build(force: bool = False, verbose: bool = False) -> str:
    digest = _digest()
    if not force and os.path.exists(LIB) and os.path.exists(STAMP) and open(STAMP).read().strip() == digest:
        return LIB
    nvcc = shutil.which('nvcc') or '/usr/local/cuda/bin/nvcc'
    if not os.path.exists(nvcc):
        raise RuntimeError('nvcc not found: cannot build libmvn_b200.so')
    cmd = [nvcc] + NVCC_FLAGS + ['-o', LIB] + _sources()
    res = subprocess.run(cmd, capture_output=True, text=True)
    log = res.stdout + res.stderr
    with open(os.path.join(HERE, 'build.log'), 'w') as f:
        f.write(' '.join(cmd) + '\n' + log)
    if res.returncode != 0:
        raise RuntimeError('nvcc failed:\n' + log[-6000:])
    if verbose:
        print(log)
    with open(STAMP, 'w') as f:
        f.write(digest)
    return LIB


if __name__ == '__main__':
    print(build(force='--force' in sys.argv, verbose='-v' in sys.argv))
