"""Multi-GPU parity check (run under torchrun, one rank per GPU); the check itself lives in bench.py (`dp_check`), which runs
it before the timed region of every N > 1 run and prints it in the JSON line:
  torchrun --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 tools/dp_check.py"""
import json
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    res = bench.dp_check(rank, world)
    if rank == 0:
        print(json.dumps(res))
        print("dp_check OK")
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
