"""Compiles the CUDA sources in csrc/ into libtimbretrap_b200.so (sm_100a only, in-tree)."""

import glob
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, 'csrc')
OUT = os.path.join(HERE, 'libtimbretrap_b200.so')

NVCC_FLAGS = ['-gencode', 'arch=compute_100a,code=sm_100a', '-lineinfo', '-O3', '-std=c++17',
              '--expt-relaxed-constexpr', '-Xcompiler', '-fPIC'] + os.environ.get('TT_NVCC_EXTRA', '').split()   # e.g. -D switches of A/B builds


def sources():
    return sorted(glob.glob(os.path.join(CSRC, '*.cu')))


def needs_build():
    if not os.path.exists(OUT):
        return True
    t = os.path.getmtime(OUT)
    deps = sources() + glob.glob(os.path.join(CSRC, '*.cuh')) + glob.glob(os.path.join(HERE, '..', 'include', '*.h'))
    return any(os.path.getmtime(d) > t for d in deps)


def build(force=False, verbose=False):
    if not force and not needs_build():
        return OUT
    nvcc = os.environ.get('NVCC', 'nvcc')
    objs = []
    os.makedirs(os.path.join(HERE, 'build'), exist_ok=True)
    procs = []
    for src in sources():
        obj = os.path.join(HERE, 'build', os.path.basename(src)[:-3] + '.o')
        cmd = [nvcc] + NVCC_FLAGS + (['-Xptxas', '-v'] if verbose else []) + ['-c', src, '-o', obj]
        procs.append((cmd, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
        objs.append(obj)
    for cmd, p in procs:
        out, _ = p.communicate()
        if verbose or p.returncode:
            print(out)
        if p.returncode:
            raise RuntimeError('nvcc failed: ' + ' '.join(cmd))
    link = [nvcc, '-gencode', 'arch=compute_100a,code=sm_100a', '-shared', '-o', OUT] + objs + ['-lcuda']
    subprocess.check_call(link)
    return OUT


SELFTEST_SRC = os.path.join(HERE, '..', 'tests', 'csrc', 'umma_probe.cu')
SELFTEST_OUT = os.path.join(HERE, '..', 'tests', 'csrc', 'libtt_selftest.so')


def build_selftest(force=False):
    """tests/csrc/libtt_selftest.so: the tcgen05 / TMEM plumbing self-test (test infrastructure, not part of the product ABI)."""
    deps = [SELFTEST_SRC] + glob.glob(os.path.join(CSRC, '*.cuh'))
    if not force and os.path.exists(SELFTEST_OUT) and all(os.path.getmtime(d) <= os.path.getmtime(SELFTEST_OUT) for d in deps):
        return SELFTEST_OUT
    nvcc = os.environ.get('NVCC', 'nvcc')
    subprocess.check_call([nvcc] + NVCC_FLAGS + ['-shared', SELFTEST_SRC, '-o', SELFTEST_OUT, '-lcuda'])
    return SELFTEST_OUT


if __name__ == '__main__':
    print(build(force='--force' in sys.argv, verbose='-v' in sys.argv))
    print(build_selftest(force='--force' in sys.argv))
