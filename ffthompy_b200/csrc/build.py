"""Build libffthom_b200.so in-tree with nvcc for sm_100a (no torch extension machinery:
the library is a plain C-ABI shared object loaded through ctypes)."""
import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
PKG = os.path.dirname(HERE)
SOURCES = ['fh_fft.cu', 'fh_pointwise.cu', 'fh_fused.cu', 'fh_reg3.cu', 'fh_mid2.cu', 'fh_slab2.cu', 'fh_material.cu', 'fh_mid512.cu', 'fh_host.cu', 'fh_odd.cu']
HEADERS = sorted(f for f in os.listdir(HERE) if f.endswith('.cuh') or f.endswith('.h')) + [os.path.join('..', '..', 'include', 'ffthom_b200.h')]
LIB = os.path.join(PKG, 'libffthom_b200.so')
NVCC = os.environ.get('NVCC', '/usr/local/cuda/bin/nvcc')
FLAGS = ['-gencode', 'arch=compute_100a,code=sm_100a', '-O3', '-lineinfo', '-std=c++17',
         '-Xcompiler', '-fPIC', '-Xptxas', '-v' if os.environ.get('FH_PTXAS_V') else '-O3']
# development builds: FH_SPLIT_COMPILE=1 adds nvcc's -split-compile 0 (fh_fused.cu 6.5 -> 4 min on 8 cores).  It changes
# instruction selection of some kernels, so shipped / measured builds do not use it.
if os.environ.get('FH_SPLIT_COMPILE'):
    FLAGS += ['-split-compile', '0']


def _stale(target, deps):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(d) > t for d in deps)


def build(force=False, verbose=False):
    os.makedirs(os.path.join(HERE, 'build'), exist_ok=True)
    hdrs = [os.path.join(HERE, h) for h in HEADERS] + [os.path.abspath(__file__)]
    objs = []
    jobs = []
    for src in SOURCES:
        s = os.path.join(HERE, src)
        o = os.path.join(HERE, 'build', src.replace('.cu', '.o'))
        objs.append(o)
        # fh_reg3.cuh (kernel bodies) is included by fh_reg3.cu only; the other units see fh_reg3.h
        # kernel-body headers that only some units include (keeps the 7-minute fh_fused.cu out of their edit cycle)
        private = {'fh_reg3.cuh': ('fh_reg3.cu',), 'fh_mid2.cuh': ('fh_mid2.cu', 'fh_mid512.cu'), 'fh_mid3.cuh': ('fh_mid2.cu',),
                   'fh_mid512.h': ('fh_reg3.cu', 'fh_mid512.cu'), 'fh_odd.h': ('fh_fused.cu', 'fh_odd.cu')}
        deps = [h for h in hdrs if os.path.basename(h) not in private or src in private[os.path.basename(h)]]
        if force or _stale(o, [s] + deps):
            jobs.append([NVCC] + FLAGS + ['-c', s, '-o', o])

    def run(cmd):
        if verbose:
            print(' '.join(cmd), flush=True)
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError('nvcc failed:\n' + ' '.join(cmd) + '\n' + r.stdout + r.stderr)
        return r.stdout + r.stderr

    with ThreadPoolExecutor(max_workers=4) as ex:
        outs = list(ex.map(run, jobs))
    if verbose:
        for o in outs:
            if o.strip():
                print(o)
    if jobs or force or _stale(LIB, objs):
        run([NVCC, '-shared', '-o', LIB] + objs + ['-gencode', 'arch=compute_100a,code=sm_100a'])
    return LIB


if __name__ == '__main__':
    print(build(force='--force' in sys.argv, verbose=True))
