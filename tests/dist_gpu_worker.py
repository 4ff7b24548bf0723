"""Worker for tests/test_gpu_multi.py: one process per GPU (torchrun, NCCL over NVLink).  ONE
operand is split across the GPUs; every rank runs the normal kernels through the C ABI on its
chunk and bee_code_interpreter_b200/parallel.py combines the per-rank partials.
Device memory for the exchanged arrays is torch's (the C ABI takes plain device pointers)."""
import json
import os
import sys
from pathlib import Path

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from bee_code_interpreter_b200 import _lib as L, device as dev  # noqa: E402
from bee_code_interpreter_b200 import parallel as par  # noqa: E402



def main():
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    rank, world = dist.get_rank(), dist.get_world_size()
    dev.init(local)
    from bee_code_interpreter_b200 import split_check

    out = split_check.run(dist, dev, rank, world)
    if rank == 0:
        print("RESULT " + json.dumps(out), flush=True)
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
