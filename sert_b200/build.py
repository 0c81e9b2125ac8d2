"""In-tree build of libsert_b200.so (nvcc, sm_100a only).

`python -m sert_b200.build` or `sert_b200.build.build()`; objects are cached by source mtime.
The shared library is written next to this file so it travels with the repository snapshot.
"""
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, 'csrc')
OBJ = os.path.join(HERE, 'build')
LIB = os.path.join(HERE, 'libsert_b200.so')

NVCC_FLAGS = [
    '-std=c++17', '-O3', '-lineinfo',
    '-gencode', 'arch=compute_100a,code=sm_100a',
    '-Xcompiler', '-fPIC', '-Xcompiler', '-fvisibility=hidden',
    '--expt-relaxed-constexpr',
]


def _nvcc():
    nvcc = shutil.which('nvcc') or '/usr/local/cuda/bin/nvcc'
    if not os.path.exists(nvcc):
        raise RuntimeError('nvcc not found: libsert_b200 cannot be built')
    return nvcc


def sources():
    return sorted(os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith('.cu'))


def _headers_mtime():
    hdrs = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith(('.cuh', '.h'))]
    hdrs.append(os.path.join(HERE, '..', 'include', 'sert_b200.h'))
    return max(os.path.getmtime(h) for h in hdrs)


def build(force=False, verbose=False, extra_flags=()):
    nvcc = _nvcc()
    os.makedirs(OBJ, exist_ok=True)
    hm = _headers_mtime()
    objs, rebuilt = [], False
    procs = []
    for src in sources():
        obj = os.path.join(OBJ, os.path.basename(src)[:-3] + '.o')
        objs.append(obj)
        stale = (force or not os.path.exists(obj) or
                 os.path.getmtime(obj) < max(os.path.getmtime(src), hm))
        if stale:
            cmd = [nvcc] + NVCC_FLAGS + list(extra_flags) + ['-c', src, '-o', obj]
            if verbose:
                print(' '.join(cmd))
            procs.append((src, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT)))
            rebuilt = True
    for src, p in procs:
        out, _ = p.communicate()
        if p.returncode != 0:
            raise RuntimeError('nvcc failed on %s:\n%s' % (src, out.decode()))
        if verbose and out:
            print(out.decode())
    if rebuilt or not os.path.exists(LIB):
        cmd = [nvcc, '-shared', '-o', LIB] + objs + ['-gencode', 'arch=compute_100a,code=sm_100a']
        out = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT)
        if out.returncode != 0:
            raise RuntimeError('link failed:\n%s' % out.stdout.decode())
    return LIB


if __name__ == '__main__':
    print(build(force='--force' in sys.argv, verbose=True))
