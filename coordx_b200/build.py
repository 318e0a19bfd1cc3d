"""Build libcoordx_b200.so (the C-ABI library of include/coordx_b200.h) in-tree with nvcc
for sm_100a.  `python -m coordx_b200.build` or `__graft_entry__.build()`.

Translation units compiled in parallel and linked into one shared library:
  coordx_abi.cu     C ABI + the FP32 (FFMA) kernels
  trunk_fwd_tc.cu   tcgen05 fused fusion+trunk forward (CTA pair); compiled FWD_COPIES times (-DCOORDX_FWD_COPY=k):
                    each copy of the kernel owns one __constant__ image of the epilogue constants
  trunk_bwd_tc.cu   tcgen05 fused trunk backward: host side, recompute (fallback) passes
  trunk_bwd_tc2.cu  tcgen05 trunk backward: dW pass (W_h / dZ^T in tensor memory) and the merged feature-gradient pass
"""
import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, 'csrc')
OBJ = os.path.join(HERE, '_build')
LIB = os.path.join(HERE, 'libcoordx_b200.so')
FWD_COPIES = 4          # keep in step with EPI_COPIES in trunk_fwd_tc.cu
# (source, object stem, extra nvcc flags)
UNITS = [('coordx_abi.cu', 'coordx_abi', []), ('trunk_bwd_tc.cu', 'trunk_bwd_tc', []), ('trunk_bwd_tc2.cu', 'trunk_bwd_tc2', [])] + \
        [('trunk_fwd_tc.cu', 'trunk_fwd_tc' if k == 0 else f'trunk_fwd_tc_copy{k}', [f'-DCOORDX_FWD_COPY={k}']) for k in range(FWD_COPIES)]
HEADERS = [os.path.join(CSRC, f) for f in sorted(os.listdir(CSRC)) if f.endswith(('.cuh', '.h'))] + \
          [os.path.join(os.path.dirname(HERE), 'include', 'coordx_b200.h')]

NVCC_FLAGS = ['-gencode', 'arch=compute_100a,code=sm_100a', '-lineinfo', '-O3', '-std=c++17',
              '-Xcompiler', '-fPIC', '-Xptxas', '-v', '--expt-relaxed-constexpr']


def _stale(target, deps):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(d) > t for d in deps)


def needs_build():
    return any(_stale(os.path.join(OBJ, stem + '.o'), [os.path.join(CSRC, src)] + HEADERS) for src, stem, _ in UNITS) or \
        _stale(LIB, [os.path.join(OBJ, stem + '.o') for _, stem, _ in UNITS if os.path.exists(os.path.join(OBJ, stem + '.o'))])


def _compile(nvcc, unit):
    src, obj = os.path.join(CSRC, unit[0]), os.path.join(OBJ, unit[1] + '.o')
    if not _stale(obj, [src] + HEADERS):
        return unit[1], 0, ''
    r = subprocess.run([nvcc] + NVCC_FLAGS + unit[2] + ['-c', '-o', obj, src], capture_output=True, text=True)
    return unit[1], r.returncode, r.stdout + r.stderr


def build(force=False, verbose=False):
    if not force and not needs_build():
        return LIB
    nvcc = os.environ.get('NVCC', '/usr/local/cuda/bin/nvcc')
    os.makedirs(OBJ, exist_ok=True)
    if force:
        for u in UNITS:
            o = os.path.join(OBJ, u[1] + '.o')
            if os.path.exists(o):
                os.remove(o)
    with ThreadPoolExecutor(len(UNITS)) as ex:
        results = list(ex.map(lambda u: _compile(nvcc, u), UNITS))
    log = []
    for unit, rc, out in results:
        log.append(f'==== {unit}\n{out}')
        if rc != 0:
            sys.stderr.write(out)
            raise RuntimeError(f'nvcc failed on {unit}')
    objs = [os.path.join(OBJ, u[1] + '.o') for u in UNITS]
    r = subprocess.run([nvcc, '-gencode', 'arch=compute_100a,code=sm_100a', '-shared', '-o', LIB] + objs,
                       capture_output=True, text=True)
    if r.returncode != 0:
        sys.stderr.write(r.stdout + r.stderr)
        raise RuntimeError('nvcc failed linking libcoordx_b200.so')
    with open(os.path.join(OBJ, 'ptxas.log'), 'w') as f:
        f.write('\n'.join(log))
    if verbose:
        print('\n'.join(log))
    return LIB


if __name__ == '__main__':
    print(build(force='--force' in sys.argv, verbose='-v' in sys.argv))
