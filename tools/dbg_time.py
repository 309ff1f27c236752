import sys, os, torch, numpy as np
sys.path.insert(0, "/root/repo")
from ml_super_resolution_b200 import ops
g = torch.Generator(device="cuda").manual_seed(0)
n,h,w=18,2160,252
x = ops.fpa_empty(n,h,w,64); x.data.normal_(generator=g); y = ops.fpa_empty(n,h,w,64)
wp = ops.pack_conv_weights(torch.randn((3,3,64,64), device="cuda", generator=g)/24); b = torch.zeros(64, device="cuda")
for _ in range(3): ops.conv_tc(x, wp, b, 3, "relu", out=y)
torch.cuda.synchronize()
e0,e1=torch.cuda.Event(enable_timing=True),torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(10): ops.conv_tc(x, wp, b, 3, "relu", out=y)
e1.record(); torch.cuda.synchronize()
ms=e0.elapsed_time(e1)/10
print(f"SRK_DBG={os.environ.get('SRK_DBG','0')}: {ms:.4f} ms  {2.0*n*h*w*576*64/ms/1e9:.0f} TFLOP/s")
