"""Runs srk_espcn_forward a few times on 4 x 1080p-LR frames (the bench workload) -- the target of ncu captures."""
import sys

import torch

sys.path.insert(0, "/root/repo")
from ml_super_resolution_b200.espcn.model_espcn import EspcnNet  # noqa: E402

channels = int(sys.argv[1]) if len(sys.argv) > 1 else 1
u8 = len(sys.argv) > 2 and sys.argv[2] == "u8"
net = EspcnNet(None, 3, channels)
g = torch.Generator(device="cuda").manual_seed(0)
x = torch.rand((4, 1080, 1920, channels), device="cuda", generator=g) * 2 - 1
out = torch.empty((4, 3240, 5760, channels), dtype=torch.uint8 if u8 else torch.float32, device="cuda")
for _ in range(3):
    net.forward_fused(x, out=out, uint8=u8)
torch.cuda.synchronize()
print("done")
