#!/usr/bin/env python
"""Host<->device copy ceilings of this box: pinned 1-D copies each way, both ways at once, and pitched 2-D copies."""
import time, torch, ctypes as C, sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
n = 1 << 30
h = torch.empty(n, dtype=torch.uint8).pin_memory(); h2 = torch.empty(n, dtype=torch.uint8).pin_memory()
d = torch.empty(n, dtype=torch.uint8, device='cuda'); d2 = torch.empty(n, dtype=torch.uint8, device='cuda')
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
def t(fn, reps=5):
    fn(); torch.cuda.synchronize(); t0 = time.time()
    for _ in range(reps): fn()
    torch.cuda.synchronize(); return (time.time() - t0) / reps
a = t(lambda: d.copy_(h, non_blocking=True)); print('H2D 1D  %.1f GB/s' % (n / a / 1e9))
b = t(lambda: h2.copy_(d2, non_blocking=True)); print('D2H 1D  %.1f GB/s' % (n / b / 1e9))
def both():
    with torch.cuda.stream(s1): d.copy_(h, non_blocking=True)
    with torch.cuda.stream(s2): h2.copy_(d2, non_blocking=True)
c = t(both); print('both    %.1f GB/s each (%.1f total)' % (n / c / 1e9, 2 * n / c / 1e9))
from archive_pdf_tools_b200 import _lib
L = _lib.lib()
row, pitch = 7650, 7664; rows = n // pitch
st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
e = t(lambda: L.b200mrc_copy2d(C.c_void_p(d.data_ptr()), pitch, C.c_void_p(h.data_ptr()), row, row, rows, 1, st)); print('H2D 2D  %.1f GB/s' % (row * rows / e / 1e9))
f = t(lambda: L.b200mrc_copy2d(C.c_void_p(h2.data_ptr()), row, C.c_void_p(d2.data_ptr()), pitch, row, rows, 2, st)); print('D2H 2D  %.1f GB/s' % (row * rows / f / 1e9))
