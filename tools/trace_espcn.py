"""In-kernel timeline of espcn_fused_kernel's roles on CTA 0 (development tool).  Build the trace library here with
ml_super_resolution_b200.build.build_trace_library(), then on the GPU box:  python tools/trace_espcn.py [first_row n_rows]"""
import ctypes as C
import os
import sys

import numpy as np
import torch

sys.path.insert(0, "/root/repo")
from ml_super_resolution_b200 import _ffi  # noqa: E402

TRACE_LIB = os.path.join(os.path.dirname(_ffi.LIB_PATH), "build", "libsrk_trace.so")
_ffi.LIB_PATH = TRACE_LIB
from ml_super_resolution_b200.espcn.model_espcn import EspcnNet  # noqa: E402

v0 = int(sys.argv[1]) if len(sys.argv) > 1 else 100
nv = int(sys.argv[2]) if len(sys.argv) > 2 else 12
channels = int(sys.argv[3]) if len(sys.argv) > 3 else 1
net = EspcnNet(None, 3, channels)
g = torch.Generator(device="cuda").manual_seed(0)
x = torch.rand((4, 1080, 1920, channels), device="cuda", generator=g) * 2 - 1
out = torch.empty((4, 3240, 5760, channels), device="cuda")
for _ in range(3):
    net.forward_fused(x, out=out)
torch.cuda.synchronize()
buf = np.zeros(24 * 256, dtype=np.uint64)
raw = C.CDLL(TRACE_LIB)
assert raw.srk_debug_ef_trace_read(buf.ctypes.data_as(C.c_void_p)) == 0
T = buf.reshape(24, 256).astype(np.int64)
names = ["g:rows", "g:Afree", "g:built", "M1:go", "M1:iss", "E1:full", "E1:drain", "E1:slot", "E1:done", "M2:go", "M2:iss", "E2:full", "E2:drain",
         "E2:slot", "E2:done", "M3:go", "M3:iss", "E3:full", "E3:drain", "E3:store", "E3:done", "M1:cmpl", "M2:cmpl", "M3:cmpl"]
t0 = T[3, v0]
print(f"cycles relative to MMA1 go of row {v0}; rows {v0}..{v0 + nv - 1}; period over 200 rows: {(T[3, 220] - T[3, 20]) / 200:.0f} cycles/row")
print("row   " + " ".join(f"{n:>9s}" for n in names))
for v in range(v0, v0 + nv):
    print(f"{v:4d}  " + " ".join(f"{T[e, v] - t0:9d}" for e in range(24)))
