"""Build libconmech_b200.so (sm_100a only) in-tree with nvcc.

    python -m conmech_b200.build [--force]

The library is a plain C-ABI shared object (include/conmech_b200.h); it links the CUDA runtime
statically, so loading it needs neither torch nor a GPU.
"""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, 'csrc')
LIBDIR = os.path.join(HERE, 'lib')
LIB = os.path.join(LIBDIR, 'libconmech_b200.so')
NVCC = os.environ.get('NVCC', '/usr/local/cuda/bin/nvcc')

ARCH = ['-gencode', 'arch=compute_100a,code=sm_100a']
COMMON = ['-O3', '-std=c++17', '-lineinfo', '-Xcompiler', '-fPIC', '-Xptxas', '-v']
# per-file extra flags: the element kernel mirrors Python float expressions, no FMA contraction
SOURCES = [
    ('cm_api.cu', []),
    ('k1_element.cu', ['-fmad=false']),
    ('k2_dofmap_assemble.cu', []),
    ('k3_factor.cu', []),
    ('k3_factor_flow.cu', []),
    ('k4_solve_recover.cu', []),
]


def _newer(a, b):
    return (not os.path.exists(b)) or os.path.getmtime(a) > os.path.getmtime(b)


def build_host_packer(force=False, logs=None):
    """_cmpack: CPython extension (plain C, gcc) that packs Python id lists into subset bit masks."""
    import sysconfig
    src = os.path.join(HERE, 'csrc_host', 'cmpack.c')
    out = os.path.join(HERE, '_cmpack' + sysconfig.get_config_var('EXT_SUFFIX'))
    if force or _newer(src, out):
        cmd = [os.environ.get('CC', 'gcc'), '-O2', '-shared', '-fPIC', '-I' + sysconfig.get_paths()['include'], src, '-o', out]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if logs is not None:
            logs.append('$ ' + ' '.join(cmd) + '\n' + r.stdout + r.stderr)
        if r.returncode != 0:
            sys.stderr.write(r.stdout + r.stderr)
            raise RuntimeError('gcc failed on cmpack.c')
    return out


def build(force=False, verbose=False):
    os.makedirs(LIBDIR, exist_ok=True)
    headers = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith('.cuh')]
    headers.append(os.path.join(HERE, '..', 'include', 'conmech_b200.h'))
    objs = []
    relink = force or not os.path.exists(LIB)
    logs = []
    for src, extra in SOURCES:
        s = os.path.join(CSRC, src)
        o = os.path.join(LIBDIR, src.replace('.cu', '.o'))
        objs.append(o)
        if force or _newer(s, o) or any(_newer(h, o) for h in headers):
            cmd = [NVCC] + ARCH + COMMON + extra + ['-c', s, '-o', o]
            r = subprocess.run(cmd, capture_output=True, text=True)
            logs.append('$ ' + ' '.join(cmd) + '\n' + r.stdout + r.stderr)
            if r.returncode != 0:
                sys.stderr.write(logs[-1])
                raise RuntimeError('nvcc failed on ' + src)
            relink = True
    if relink:
        cmd = [NVCC] + ARCH + ['-shared', '-Xcompiler', '-fPIC', '-o', LIB] + objs
        r = subprocess.run(cmd, capture_output=True, text=True)
        logs.append('$ ' + ' '.join(cmd) + '\n' + r.stdout + r.stderr)
        if r.returncode != 0:
            sys.stderr.write(logs[-1])
            raise RuntimeError('link failed')
    build_host_packer(force, logs)
    if logs:
        with open(os.path.join(LIBDIR, 'build.log'), 'w') as f:
            f.write('\n'.join(logs))
        if verbose:
            print('\n'.join(logs))
    return LIB


if __name__ == '__main__':
    print(build(force='--force' in sys.argv, verbose=True))
