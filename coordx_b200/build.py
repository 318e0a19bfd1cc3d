"""Build libcoordx_b200.so (the C-ABI library of include/coordx_b200.h) in-tree with nvcc
for sm_100a.  `python -m coordx_b200.build` or `__graft_entry__.build()`."""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, 'csrc', 'coordx_abi.cu')
LIB = os.path.join(HERE, 'libcoordx_b200.so')
DEPS = [os.path.join(HERE, 'csrc', f) for f in os.listdir(os.path.join(HERE, 'csrc'))] + \
       [os.path.join(os.path.dirname(HERE), 'include', 'coordx_b200.h')]

NVCC_FLAGS = ['-gencode', 'arch=compute_100a,code=sm_100a', '-lineinfo', '-O3', '-std=c++17',
              '-Xcompiler', '-fPIC', '-shared', '-Xptxas', '-v', '--expt-relaxed-constexpr']


def needs_build():
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    return any(os.path.getmtime(d) > t for d in DEPS)


def build(force=False, verbose=False):
    if not force and not needs_build():
        return LIB
    nvcc = os.environ.get('NVCC', '/usr/local/cuda/bin/nvcc')
    cmd = [nvcc] + NVCC_FLAGS + ['-o', LIB, SRC]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        sys.stderr.write(r.stdout + r.stderr)
        raise RuntimeError('nvcc failed building libcoordx_b200.so')
    log = os.path.join(HERE, 'csrc', 'ptxas.log')
    with open(log, 'w') as f:
        f.write(r.stderr)
    if verbose:
        print(r.stderr)
    return LIB


if __name__ == '__main__':
    print(build(force='--force' in sys.argv, verbose=True))
