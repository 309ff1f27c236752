import sys, os, torch
sys.path.insert(0, "/root/repo")
from ml_super_resolution_b200 import ops
from ml_super_resolution_b200.espcn.model_espcn import EspcnNet
from ml_super_resolution_b200.tiling import plan_tiles
g = torch.Generator(device="cuda").manual_seed(0)
net = EspcnNet(None, 3, 1)
F = 4
lr = torch.rand((F, 1080, 1920, 1), device="cuda", generator=g) * 2 - 1
out = torch.empty((F, 3240, 5760, 1), device="cuda")
Ht, Wt, tiles = plan_tiles(F, 1080, 1920, 4, max_w=int(os.environ.get('PANEL_W', '254')))
panels = ops.make_panels([t.as_tuple() for t in tiles])
t1, t2 = net._get_bufs(len(tiles), Ht, Wt)
a = net.arena
def tm(fn, n=int(os.environ.get('N_REP', '10'))):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n
f1 = tm(lambda: ops.conv_first_tc(lr, net.plan.views[net._i1], a.view("f1/bias:0"), 5, "SAME", "tanh", panels=panels, panel_hw=(Ht, Wt), out=t1))
f2 = tm(lambda: ops.conv_tc(t1, net.plan.views[net._i2], a.view("f2/bias:0"), 3, "tanh", out=t2))
f3 = tm(lambda: ops.conv_tc_last(t2, net.plan.views[net._i3], net.bias3, 3, net.cout3, None, shuffle_r=3, panels=panels, frame_shape=(F, 1080, 1920), out=out))
print(f"SRK_DBG={os.environ.get('SRK_DBG','0')}: f1 {f1:.3f}  f2 {f2:.3f}  f3 {f3:.3f} ms for {F} frames; panels {len(tiles)//F} of {Ht}x{Wt}")
