import sys, os
sys.path.insert(0, "/root/repo")
import numpy as np, torch
from ml_super_resolution_b200 import _ffi
_ffi.LIB_PATH = os.environ["SRK_DEV_LIB"]
from ml_super_resolution_b200.espcn.model_espcn import EspcnNet
from oracle import models as OM
params = OM.espcn_init(seed=9, scaling_factor=3, channels=1)
params = {k: (v * 5 if k.endswith("kernel:0") else v + 0.01) for k, v in params.items()}
net = EspcnNet(params, 3, 1)
for shape in [(1, 40, 100), (1, 300, 100)]:
    n, h, w = shape
    lr = OM.synthetic_images(5, n, h, w, 1)
    x = torch.from_numpy(lr).cuda()
    ref = OM.espcn_forward(params, lr)
    got = net.forward(x, shuffle=False).cpu().numpy()
    err = np.abs(got - ref).max(axis=(0, 3))  # [h, w]
    bad_rows = np.where(err.max(axis=1) > 0.05)[0]
    print(shape, "bad rows:", bad_rows[:40], "count", len(bad_rows))
    if len(bad_rows):
        r = bad_rows[0]
        print(" row", r, "bad cols:", np.where(err[r] > 0.05)[0][:40])
