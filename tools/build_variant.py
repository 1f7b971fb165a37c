"""Build a timing variant of the library: tools/build_variant.py NAME -DCM_EXP=1 ...
Recompiles the factorisation kernel with the given defines, links it with the production objects
into gpurun_scratch/lib_NAME.so (selected at run time with CONMECH_B200_LIB=...)."""
import os
import subprocess
import sys

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
from conmech_b200 import build as b      # noqa: E402

name, defs = sys.argv[1], sys.argv[2:]
b.build()
out = os.path.join(REPO, 'gpurun_scratch')
os.makedirs(out, exist_ok=True)
files = [f for f in os.environ.get('CM_VARIANT_FILES', 'k3_factor_flow.cu').split(',')]
objs = []
for src, extra in b.SOURCES:
    o = os.path.join(b.LIBDIR, src.replace('.cu', '.o'))
    if src in files:
        o = os.path.join(out, name + '_' + src.replace('.cu', '.o'))
        cmd = [b.NVCC] + b.ARCH + b.COMMON + extra + defs + ['-c', os.path.join(b.CSRC, src), '-o', o]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            sys.exit(r.stdout + r.stderr)
        print('\n'.join(l for l in (r.stdout + r.stderr).splitlines() if 'k3_factor_flowILi4ELb0' in l or 'spill' in l)[:2000])
    objs.append(o)
lib = os.path.join(out, 'lib_' + name + '.so')
subprocess.check_call([b.NVCC] + b.ARCH + ['-shared', '-Xcompiler', '-fPIC', '-o', lib] + objs)
print(lib)
