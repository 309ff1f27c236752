"""Development check of srk_espcn_forward (csrc/espcn_fused.cu) on the GPU box: parity against the CPU oracle at small sizes,
against the layer-by-layer kernels at 1080p-LR, row-band sharding bit-identity, and kernel time.  Not a pytest: the
parity cases live in tests/test_models_gpu.py; this prints numbers while the kernel is being tuned."""
import sys
import time

import numpy as np
import torch

sys.path.insert(0, "/root/repo")
import os  # noqa: E402

from ml_super_resolution_b200 import _ffi  # noqa: E402

if os.environ.get("SRK_DEV_LIB"):  # development builds of the library (kernel variants under tuning); the product loader has no override
    _ffi.LIB_PATH = os.environ["SRK_DEV_LIB"]
from ml_super_resolution_b200.espcn.model_espcn import EspcnNet  # noqa: E402
from oracle import models as OM  # noqa: E402
from oracle import ops as O  # noqa: E402


def trained_like(p, scale=5.0, seed=3):
    rng = np.random.default_rng(seed)
    return {k: (v * scale if k.endswith("kernel:0") else v + rng.normal(0, 0.05, v.shape).astype(np.float32)) for k, v in p.items()}


def main():
    quick = len(sys.argv) > 1 and sys.argv[1] == "quick"
    only_time = len(sys.argv) > 1 and sys.argv[1] == "time"
    ok = True
    for channels, r, shape in [] if only_time else [(1, 3, (1, 36, 64)), (3, 3, (2, 17, 17)), (3, 2, (1, 20, 300)), (3, 4, (1, 9, 11)), (1, 3, (2, 130, 250)),
                               (1, 2, (1, 7, 121)), (1, 4, (1, 33, 119))]:
        n, h, w = shape
        params = trained_like(OM.espcn_init(seed=9, scaling_factor=r, channels=channels))
        net = EspcnNet(params, r, channels)
        lr = OM.synthetic_images(5, n, h, w, channels)
        x = torch.from_numpy(lr).cuda()
        ref = OM.espcn_forward(params, lr)
        packed = net.forward(x, shuffle=False).cpu().numpy()
        shuffled = net.forward(x, shuffle=True).cpu().numpy()
        layered = net.forward(x, shuffle=False, fused=False).cpu().numpy()
        u8 = net.forward_fused(x, shuffle=True, uint8=True).cpu().numpy()
        e_f, e_l = np.abs(packed - ref).max(), np.abs(layered - ref).max()
        exact = np.array_equal(shuffled, O.pixel_shuffle(packed, r))
        u8_ref = np.clip(shuffled * np.float32(127.5) + np.float32(127.5), 0, 255).astype(np.uint8)
        u8_ok = np.array_equal(u8, u8_ref)
        # row-band sharding: three ranks write disjoint bands of one buffer
        out = torch.full_like(torch.from_numpy(shuffled), float("nan")).cuda()
        for rk in range(3):
            net.forward_fused(x, shuffle=True, out=out, rank=rk, world=3)
        band_ok = np.array_equal(out.cpu().numpy(), shuffled)
        good = e_f <= 4e-2 and exact and u8_ok and band_ok
        ok &= good
        print(f"C={channels} r={r} {shape}: fused max|err| {e_f:.3e} (layered {e_l:.3e}); shuffle exact {exact}; u8 exact {u8_ok}; bands exact {band_ok}"
              f"  {'OK' if good else 'FAIL'}", flush=True)
    if quick:
        sys.exit(0 if ok else 1)
    # full size: one 1080p-LR frame against the layered kernels and the fp32 oracle
    for channels in () if only_time else (1, 3):
        params = trained_like(OM.espcn_init(seed=11, scaling_factor=3, channels=channels))
        net = EspcnNet(params, 3, channels)
        lr = OM.synthetic_images(7, 1, 1080, 1920, channels)
        x = torch.from_numpy(lr).cuda()
        fused = net.forward(x, shuffle=False).cpu().numpy()
        layered = net.forward(x, shuffle=False, fused=False).cpu().numpy()
        t0 = time.time()
        ref = OM.espcn_forward(params, lr, dtype=np.float32)
        print(f"1080p C={channels}: fused vs oracle(fp32) {np.abs(fused - ref).max():.3e}; layered vs oracle {np.abs(layered - ref).max():.3e}; "
              f"fused vs layered {np.abs(fused - layered).max():.3e}  (oracle {time.time() - t0:.1f} s)", flush=True)
        ok &= np.abs(fused - ref).max() <= 4e-2
    # timing: 4 frames per step, like bench.py
    for channels, u8 in ((1, False), (1, True), (3, False)):
        net = EspcnNet(None, 3, channels)
        F = 4
        g = torch.Generator(device="cuda").manual_seed(0)
        x = torch.rand((F, 1080, 1920, channels), device="cuda", generator=g) * 2 - 1
        out = torch.empty((F, 3240, 5760, channels), dtype=torch.uint8 if u8 else torch.float32, device="cuda")
        for fused in (True, False):
            if u8 and not fused:
                continue
            run = (lambda: net.forward_fused(x, out=out, uint8=u8)) if fused else (lambda: net.forward(x, out=out, fused=False))
            for _ in range(3):
                run()
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(20):
                run()
            e1.record()
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / 20
            print(f"C={channels} {'u8' if u8 else 'f32'} {'fused' if fused else 'layered'}: {ms:.3f} ms per {F} frames = {F * 3240 * 5760 / ms / 1e6:.1f} Gpix/s", flush=True)
    sys.exit(0 if ok else 1)


if __name__ == "__main__":
    main()
