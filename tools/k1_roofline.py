"""Roofline of the element kernel K1: GB/s on a large synthetic batch of elements
through the C ABI entry cm_element_kernel_dev (device buffers).  Algorithmic bytes per element:
152 B in (6 coordinates, 7 properties, 6 springs) + 153*8 B out (K_G 144, R 9)."""
import ctypes as C
import os
import sys

import numpy as np
import torch

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
from conmech_b200 import _capi      # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 20
lib = _capi.load()
g = torch.Generator(device='cuda').manual_seed(0)
xu = torch.rand(n, 3, dtype=torch.float64, device='cuda', generator=g)
xv = xu + 0.5 + torch.rand(n, 3, dtype=torch.float64, device='cuda', generator=g)
props = torch.rand(n, 7, dtype=torch.float64, device='cuda', generator=g) + 0.5
props[:, 4] *= 2.1e8
cr = torch.full((n, 6), float('inf'), dtype=torch.float64, device='cuda')
Kg = torch.empty(n, 144, dtype=torch.float64, device='cuda')
R3 = torch.empty(n, 9, dtype=torch.float64, device='cuda')
st = torch.cuda.current_stream().cuda_stream
call = lambda: _capi.check(lib.cm_element_kernel_dev(xu.data_ptr(), xv.data_ptr(), props.data_ptr(), cr.data_ptr(),
                                                     n, Kg.data_ptr(), R3.data_ptr(), st))
for _ in range(3):
    call()
torch.cuda.synchronize()
best = 1e30
for _ in range(5):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); call(); e1.record(); torch.cuda.synchronize()
    best = min(best, e0.elapsed_time(e1))
bytes_alg = n * (152 + 153 * 8)
print('K1 element kernel: n={} {:.3f} ms -> {:.1f} GB/s algorithmic ({:.1f} % of the measured 6549.8 GB/s copy rate)'.format(
    n, best, bytes_alg / best / 1e6, 100 * bytes_alg / best / 1e6 / 6549.8))
