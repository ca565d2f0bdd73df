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


def build(force: bool = False, verbose: bool = False, out: str = LIB, extra_flags=()) -> str:
    """`out` / `extra_flags`: tuning builds of the same sources (e.g. tools/libmvn_trace.so with -DMVN_TC_TRACE)."""
    digest = _digest()
    if out == LIB and not force and os.path.exists(LIB) and os.path.exists(STAMP) and open(STAMP).read().strip() == digest:
        return LIB
    nvcc = shutil.which('nvcc') or '/usr/local/cuda/bin/nvcc'
    if not os.path.exists(nvcc):
        raise RuntimeError('nvcc not found: cannot build libmvn_b200.so')
    # one nvcc per source in parallel (the fused kernels take minutes each), then one link
    from concurrent.futures import ThreadPoolExecutor
    objdir = os.path.join(HERE, '.obj')
    os.makedirs(objdir, exist_ok=True)
    cflags = [f for f in NVCC_FLAGS if f != '--shared'] + list(extra_flags)

    def compile_one(src):
        obj = os.path.join(objdir, os.path.basename(src)[:-3] + '.o')
        cmd = [nvcc] + cflags + ['-c', '-o', obj, src]
        res = subprocess.run(cmd, capture_output=True, text=True)
        return obj, ' '.join(cmd) + '\n' + res.stdout + res.stderr, res.returncode

    with ThreadPoolExecutor(max_workers=min(8, os.cpu_count() or 1)) as ex:
        results = list(ex.map(compile_one, _sources()))
    log = ''.join(r[1] for r in results)
    rc = max(r[2] for r in results)
    if rc == 0:
        cmd = [nvcc, '--shared', '-o', out] + [r[0] for r in results]
        res = subprocess.run(cmd, capture_output=True, text=True)
        log += ' '.join(cmd) + '\n' + res.stdout + res.stderr
        rc = res.returncode
    with open(os.path.join(HERE, 'build.log' if out == LIB else 'build_variant.log'), 'w') as f:
        f.write(log)
    if rc != 0:
        raise RuntimeError('nvcc failed:\n' + log[-6000:])
    if verbose:
        print(log)
    if out == LIB:
        with open(STAMP, 'w') as f:
            f.write(digest)
    return out


if __name__ == '__main__':
    # python build.py [--force] [-v] [--out PATH -DFLAG ...]
    out = sys.argv[sys.argv.index('--out') + 1] if '--out' in sys.argv else LIB
    print(build(force='--force' in sys.argv, verbose='-v' in sys.argv, out=out, extra_flags=[a for a in sys.argv[1:] if a.startswith('-D')]))
