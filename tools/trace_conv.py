"""In-kernel timeline of conv_tc_kernel's warp roles on CTA 0 (development tool).  Build the trace library here with
ml_super_resolution_b200.build.build_trace_library() (-> ml_super_resolution_b200/build/libsrk_trace.so, -DSRK_TRACE), then on the
GPU box:  python tools/trace_conv.py [f2|f3|c64|train]
The product loader knows nothing about the trace build: this tool points _ffi.LIB_PATH at it before the first load.
"""
import ctypes as C
import os
import sys

import numpy as np
import torch

sys.path.insert(0, "/root/repo")
from ml_super_resolution_b200 import _ffi  # noqa: E402

TRACE_LIB = os.path.join(os.path.dirname(_ffi.LIB_PATH), "build", "libsrk_trace.so")
_ffi.LIB_PATH = TRACE_LIB
from ml_super_resolution_b200 import ops  # noqa: E402
from ml_super_resolution_b200.espcn.model_espcn import EspcnNet  # noqa: E402
from ml_super_resolution_b200.tiling import plan_tiles  # noqa: E402

which = sys.argv[1] if len(sys.argv) > 1 else "f2"
sets = 2 if which in ("c64", "f1", "train") else 4
g = torch.Generator(device="cuda").manual_seed(0)
if which in ("c64", "train"):
    N, H, W = (64, 240, 240) if which == "c64" else (64, 41, 41)
    x = ops.fpa_empty(N, H, W, 64)
    x.data.normal_(generator=g)
    y = ops.fpa_empty(N, H, W, 64)
    w = torch.randn((3, 3, 64, 64), device="cuda", generator=g) * 0.05
    wp = ops.pack_conv_weights(w)
    b = torch.zeros(64, device="cuda")
    run = lambda: ops.conv_tc(x, wp, b, 3, "relu", out=y)  # noqa: E731
else:
    net = EspcnNet(None, 3, 1)
    F = 4
    lr = torch.rand((F, 1080, 1920, 1), device="cuda", generator=g) * 2 - 1
    out = torch.empty((F, 3240, 5760, 1), device="cuda")
    Ht, Wt, tiles = plan_tiles(F, 1080, 1920, 4)
    panels = ops.make_panels([t.as_tuple() for t in tiles])
    t1, t2 = net._get_bufs(len(tiles), Ht, Wt)
    a = net.arena
    ops.conv_first_tc(lr, net.plan.views[net._i1], a.view("f1/bias:0"), 5, "SAME", "tanh", panels=panels, panel_hw=(Ht, Wt), out=t1)
    ops.conv_tc(t1, net.plan.views[net._i2], a.view("f2/bias:0"), 3, "tanh", out=t2)
    if which == "f1":
        run = lambda: ops.conv_first_tc(lr, net.plan.views[net._i1], a.view("f1/bias:0"), 5, "SAME", "tanh", panels=panels, panel_hw=(Ht, Wt), out=t1)  # noqa: E731
    elif which == "f2":
        run = lambda: ops.conv_tc(t1, net.plan.views[net._i2], a.view("f2/bias:0"), 3, "tanh", out=t2)  # noqa: E731
    else:
        run = lambda: ops.conv_tc_last(t2, net.plan.views[net._i3], net.bias3, 3, net.cout3, None, shuffle_r=3, panels=panels,  # noqa: E731
                                       frame_shape=(F, 1080, 1920), out=out)
for _ in range(3):
    run()
torch.cuda.synchronize()
buf = np.zeros(24 * 256, dtype=np.uint64)
raw = C.CDLL(TRACE_LIB)
rc = raw.srk_debug_trace_read(buf.ctypes.data_as(C.c_void_p))
assert rc == 0, rc
T = buf.reshape(24, 256).astype(np.int64)
t0 = T[3, 0]
if which == "train":
    nt = int((T[3] > 0).sum())
    k0 = T[20, 0]
    print(f"train shape: CTA 0 has {nt} tiles; cycles from kernel entry: pdl_wait passed {T[20,1]-k0}, weights+first data -> first MMA issued {T[2,0]-k0}, "
          f"first commit {T[3,0]-k0}, first acc full {T[4,0]-k0}, last commit {T[3,nt-1]-k0}, last tile epilogue done {T[7,nt-1]-k0}, "
          f"last store freed {T[9,nt-1]-k0}, CTA exit {T[20,2]-k0}")
    print("  tile 0 seen ready at", int(T[1, 0] - k0), "; chunk TMA issue times:", [int(T[0, i] - k0) for i in range(10) if T[0, i]])
    print("  per-tile commit times:", [int(T[3, i] - k0) for i in range(nt)])
    print("  per-tile epilogue done:", [int(T[7, i] - k0) for i in range(nt)])
    sys.exit(0)
names = ["load-issued(chunk)", "mma:data-ready", "mma:acc-free", "mma:committed", "epi:acc-full", "epi:tmem-read", "epi:passes-done",
         "epi:tile-done", "store:staged", "store:freed"]
lo, hi = 24, 48
print(f"{which}: cycles relative to first MMA commit; tiles {lo}..{hi - 1}")
print("tile " + " ".join(f"{n.split(':')[-1][:10]:>10}" for n in names[1:]))
for t in range(lo, hi):
    print(f"{t:4d} " + " ".join(f"{(T[e, t] - t0) if T[e, t] else 0:10d}" for e in range(1, 10)))
d = lambda a: float(np.mean(a))  # noqa: E731
r = np.arange(lo, hi)
print("\naverages over these tiles (cycles):")
print(f"  tile period (commit -> commit)           {d(T[3, r] - T[3, r - 1]):8.0f}")
print(f"  mma: commit -> next tile seen ready      {d(T[1, r + 2] - T[3, r]):8.0f}   (same issuer, its next tile)")
print(f"  mma: wait for a free accumulator         {d(T[2, r] - T[1, r]):8.0f}")
print(f"  mma: issue (acc-free -> commit)          {d(T[3, r] - T[2, r]):8.0f}")
print(f"  commit -> epilogue sees acc full         {d(T[4, r] - T[3, r]):8.0f}")
print(f"  epi: acc full -> last tmem read          {d(T[5, r] - T[4, r]):8.0f}")
print(f"  epi: acc full -> passes done             {d(T[6, r] - T[4, r]):8.0f}")
print(f"  epi: passes done -> tile done            {d(T[7, r] - T[6, r]):8.0f}")
print(f"  epi: set idle (tile done -> next full)   {d(T[4, r + sets] - T[7, r]):8.0f}")
if T[10, lo]:
    print(f"  epi pass 0: acc full -> tmem loaded      {d(T[10, r] - T[4, r]):8.0f}")
    print(f"  epi pass 0: -> edge rows published       {d(T[11, r] - T[10, r]):8.0f}")
    print(f"  epi pass 0: -> exchange barrier passed   {d(T[12, r] - T[11, r]):8.0f}")
    print(f"  epi pass 0: -> shuffles + adds done      {d(T[13, r] - T[12, r]):8.0f}")
    print(f"  epi pass 0: -> bias/act/pack done        {d(T[14, r] - T[13, r]):8.0f}")
if T[8, lo]:
    print(f"  store: tile done -> staged seen          {d(T[8, r] - T[7, r]):8.0f}")
    print(f"  store: staged -> freed                   {d(T[9, r] - T[8, r]):8.0f}")
for a_, b_ in ((8, 32), (32, 64), (64, 128), (128, 192), (192, 250)):
    print(f"  tile period over tiles {a_:3d}..{b_:3d}: {(T[3, b_] - T[3, a_]) / (b_ - a_):7.0f} cycles")
if which == "f1":
    print(f"  gather: slot free -> rows written        {d(T[17, r] - T[16, r]):8.0f}")
    print(f"  gather: rows written -> next slot free   {d(T[16, r + 2] - T[17, r]):8.0f}   (same group, its next tile)")
    print(f"  gather done -> issuer sees A full        {d(T[1, r] - T[17, r]):8.0f}")
    sys.exit(0)
cr = np.arange(60, 200)
print(f"  producer per chunk: slot-free wait {d(T[17, cr] - T[16, cr]):6.0f} | ready arrives {d(T[0, cr] - T[17, cr]):6.0f} | expect_tx {d(T[18, cr] - T[0, cr]):6.0f} | "
      f"TMA issue {d(T[19, cr] - T[18, cr]):6.0f} | loop {d(T[16, cr + 1] - T[19, cr]):6.0f}")
ch = T[0]
n_ch = int((ch > 0).sum())
print(f"  chunks loaded by CTA 0: {n_ch}; mean issue period over chunks 40..120: {d(np.diff(ch[40:120])):.0f} cycles")
# how far ahead of the MMA's need does the load get issued?  chunk index needed by tile t is unknown here; print the raw issue times
print("  load issue times (chunks 40..56):", [int(c - t0) for c in ch[40:56]])
