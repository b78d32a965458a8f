"""Build the C-ABI shared library (morpheus_b200/libmorpheus_b200.so) with nvcc for sm_100a.

In-tree, no torch dependency: the .so only needs libcudart.  `python -m morpheus_b200.build` or
`__graft_entry__.build()`.  Object files are rebuilt only when their sources changed.
"""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, 'csrc')
OBJ = os.path.join(HERE, 'build')
LIB = os.path.join(HERE, 'libmorpheus_b200.so')
SOURCES = ['api.cu', 'grid_encode.cu', 'composite.cu', 'sampler.cu', 'field_fwd.cu', 'field_bwd.cu', 'field_fwd_tc.cu', 'field_bwd_tc.cu', 'field_bwd_sdf_tc.cu', 'field_bwd_fd_tc.cu', 'field_fd_reg_tc.cu', 'conv_tc.cu', 'glue.cu']
NVCC = os.environ.get('NVCC', '/usr/local/cuda/bin/nvcc')
FLAGS = ['-O3', '-std=c++17', '-gencode', 'arch=compute_100a,code=sm_100a', '-lineinfo', '-Xcompiler', '-fPIC',
         '--expt-relaxed-constexpr', '-Xptxas', '-v'] + os.environ.get('MB_NVCC_EXTRA', '').split()


def _newer(src_list, target):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(s) > t for s in src_list)


def build(verbose=False, force=False):
    os.makedirs(OBJ, exist_ok=True)
    headers = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith(('.cuh', '.h'))]
    headers.append(os.path.join(os.path.dirname(HERE), 'include', 'morpheus_b200.h'))
    objs = []
    procs = []
    for s in SOURCES:
        src = os.path.join(CSRC, s)
        obj = os.path.join(OBJ, s.replace('.cu', '.o'))
        objs.append(obj)
        if force or _newer([src] + headers, obj):
            cmd = [NVCC] + FLAGS + ['-c', src, '-o', obj]
            procs.append((s, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
    log = []
    for s, pr in procs:
        out, _ = pr.communicate()
        log.append(f'==== {s}\n{out}')
        if pr.returncode != 0:
            sys.stderr.write(out)
            raise RuntimeError(f'nvcc failed on {s}')
    if procs or not os.path.exists(LIB):
        cmd = [NVCC, '-shared', '-o', LIB] + objs + ['-gencode', 'arch=compute_100a,code=sm_100a', '-lcudart']
        r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
        if r.returncode != 0:
            sys.stderr.write(r.stdout)
            raise RuntimeError('link failed')
    with open(os.path.join(OBJ, 'ptxas.log'), 'a' if not force else 'w') as f:
        f.write('\n'.join(log))
    if verbose:
        print('\n'.join(log))
    return LIB


if __name__ == '__main__':
    print(build(verbose='-v' in sys.argv, force='-f' in sys.argv))
