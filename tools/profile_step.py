"""Driver for ncu launch lists: a few VDSR training steps / one ESPCN batch / one VDSR 4K frame."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
which = sys.argv[1] if len(sys.argv) > 1 else "train"
g = torch.Generator(device="cuda").manual_seed(0)
if which == "train":
    from ml_super_resolution_b200.vdsr.model_vdsr import VdsrNet
    net = VdsrNet(None, 20, 3)
    sd = torch.rand((64, 41, 41, 3), device="cuda", generator=g) * 2 - 1
    hd = torch.rand((64, 41, 41, 3), device="cuda", generator=g) * 2 - 1
    for _ in range(3):
        net.train_step(sd, hd, 1e-4)
elif which == "espcn":
    from ml_super_resolution_b200.espcn.model_espcn import EspcnNet
    net = EspcnNet(None, 3, 1)
    lr = torch.rand((4, 1080, 1920, 1), device="cuda", generator=g) * 2 - 1
    out = torch.empty((4, 3240, 5760, 1), device="cuda")
    for _ in range(3):
        net.forward(lr, True, out=out)
else:
    from ml_super_resolution_b200.vdsr.model_vdsr import VdsrNet
    net = VdsrNet(None, 20, 3)
    frame = torch.rand((1, 2160, 3840, 3), device="cuda", generator=g) * 2 - 1
    out = torch.empty_like(frame)
    for _ in range(2):
        net.forward(frame, out=out)
torch.cuda.synchronize()
print("ok")
