"""output_to_stream on the device (csrc/rt_ppm.cu) against the host writer, at a frame size.
    python profiles/bench_ppm.py [nx ny]          (ncu launch list: ncu --metrics gpu__time_duration.sum ...)"""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

import __graft_entry__ as entry

nx = int(sys.argv[1]) if len(sys.argv) > 1 else 3840
ny = int(sys.argv[2]) if len(sys.argv) > 2 else 2160
pkg = entry.load_package()
rt = pkg.RayTracer(0)
rt.set_stream(torch.cuda.current_stream().cuda_stream)
fb = torch.rand((ny, nx, 3), device="cuda").sqrt()
host = fb.cpu().numpy()
import ctypes as C
n = C.c_size_t()
ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
best = 1e9
for k in range(5):
    ev[0].record()
    rt._ck(rt.L.rt_ppm_format(rt._ctx, C.c_void_p(fb.data_ptr()), nx, ny, C.byref(n)), "rt_ppm_format")
    ev[1].record()
    torch.cuda.synchronize()
    best = min(best, ev[0].elapsed_time(ev[1]))
buf = torch.empty(n.value, dtype=torch.uint8).pin_memory()
t0 = time.perf_counter()
rt._ck(rt.L.rt_ppm_read(rt._ctx, C.c_void_p(buf.data_ptr()), n.value), "rt_ppm_read")
t_read = time.perf_counter() - t0
t0 = time.perf_counter()
want = pkg.format_ppm(host)
t_host = time.perf_counter() - t0
same = bytes(buf.numpy().tobytes()) == want
alg = nx * ny * 24 + n.value          # two reads of the float frame + the text
print(f"ppm {nx}x{ny}: text {n.value} B, device format {best:.3f} ms (len + scan + write, incl. one 8-byte readback) = {alg / best / 1e6:.0f} GB/s algorithmic, "
      f"text D2H to pinned {t_read * 1e3:.2f} ms, host writer (size + fill passes) {t_host * 1e3:.0f} ms, identical {same}")
rt.close()
