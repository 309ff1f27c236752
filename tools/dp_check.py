"""Multi-GPU parity check (run under torchrun, one rank per GPU):
  1. data-parallel training: gradients after the NCCL all-reduce equal the single-GPU gradients of the concatenated batch;
     weights stay identical across ranks after 3 steps;
  2. tile-sharded inference: ranks compute disjoint tile shards; their union equals the single-GPU frame bit for bit."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ml_super_resolution_b200.initializers import vdsr_params  # noqa: E402
from ml_super_resolution_b200.vdsr.model_vdsr import VdsrNet  # noqa: E402


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    L, per = 6, 8
    rng = np.random.default_rng(7)
    params = vdsr_params(3, L, 3)
    for k in params:
        if k.endswith("bias:0"):
            params[k] = (0.05 * rng.standard_normal(params[k].shape)).astype(np.float32)
    sd_all = torch.from_numpy(rng.uniform(-1, 1, (per * world, 41, 41, 3)).astype(np.float32)).cuda()
    hd_all = torch.from_numpy(rng.uniform(-1, 1, (per * world, 41, 41, 3)).astype(np.float32)).cuda()
    sd, hd = sd_all[rank * per:(rank + 1) * per].contiguous(), hd_all[rank * per:(rank + 1) * per].contiguous()
    # --- DP gradients vs single-GPU gradients of the whole batch
    net = VdsrNet(params, L)
    net.forward_backward(sd, hd, numel_total=float(sd_all.numel()))
    dist.all_reduce(net.arena.g)
    ref = VdsrNet(params, L)
    ref.forward_backward(sd_all, hd_all)
    err = float((net.arena.g - ref.arena.g).norm() / ref.arena.g.norm())
    loss_dp = net._train_bufs["loss"][0:1].clone()
    dist.all_reduce(loss_dp)
    lerr = abs(float(loss_dp) - float(ref._train_bufs["loss"][0])) / float(ref._train_bufs["loss"][0])
    # --- 3 DP steps keep the replicas identical
    net2 = VdsrNet(params, L)
    for _ in range(3):
        net2.train_step(sd, hd, lr=1e-3)
    w = net2.arena.w.clone()
    wmax, wmin = w.clone(), w.clone()
    dist.all_reduce(wmax, op=dist.ReduceOp.MAX)
    dist.all_reduce(wmin, op=dist.ReduceOp.MIN)
    spread = float((wmax - wmin).abs().max())
    # --- tile-sharded inference
    frame = torch.from_numpy(rng.uniform(-1, 1, (1, 200, 600, 3)).astype(np.float32)).cuda()
    out = torch.zeros_like(frame)
    net.forward(frame, out=out, tile_rows=80, rank=rank, world=world)
    dist.all_reduce(out)  # disjoint ownership: the sum assembles the frame (test-only collective)
    full = ref.forward(frame, tile_rows=80)
    same = bool(torch.equal(out, full))
    if rank == 0:
        print(f"world={world} dp_grad_rel_err={err:.3e} dp_loss_rel_err={lerr:.3e} replica_weight_spread={spread:.3e} tiled_sharded_equal={same}")
        assert err < 2e-3 and lerr < 1e-4 and spread == 0.0 and same
        print("dp_check OK")
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
