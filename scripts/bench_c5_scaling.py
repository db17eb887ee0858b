#!/usr/bin/env python
"""Strong scaling of BASELINE.json's fifth configuration (not the bench.py contract line): 65536 synthetic potentials x
500 energies x lmax 12 through phsh_wil + conphas, the potentials sharded over the GPUs of one box (one process per GPU,
no data-path collective; barrier + max-over-ranks timing on CUDA events like bench.py).

    python scripts/bench_c5_scaling.py                                   # N = 1
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \\
        scripts/bench_c5_scaling.py --steps 5 --warmup 3                  # N = 2, 4, 8
"""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

from phaseshifts_b200 import api, workloads  # noqa: E402
from phaseshifts_b200.distributed import shard_range  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--potentials", type=int, default=65536, help="total over all GPUs (strong scaling)")
    a = ap.parse_args()
    rank, world = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    real_stdout = os.dup(1)
    os.dup2(2, 1)  # library banners go to stderr; the JSON line goes to the real stdout
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    lo, hi = shard_range(a.potentials, rank, world)
    w = workloads.c5_wil(P=hi - lo, first=lo)
    rs, zs = torch.from_numpy(w["rs"]).cuda(), torch.from_numpy(w["zs"]).cuda()

    def step():
        return api.phsh_wil_batch(rs, zs, w["nrows"], w["Z"], w["RT"], w["e"], NL=13, NR=101, unwrap=True)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(a.warmup):
        out = step()
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(a.steps):
        out = step()
    e1.record()
    barrier()
    ms = torch.tensor([e0.elapsed_time(e1) / a.steps], device="cuda")
    chk = torch.tensor([float(out.double().abs().sum())], device="cuda")
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        dist.all_reduce(chk, op=dist.ReduceOp.SUM)
    units = a.potentials * 500 * 13
    if rank == 0:
        line = dict(metric="phase_shifts_per_sec", value=units / (float(ms) * 1e-3), unit="delta_l(E)/s", n_gpus=world,
                    steps=a.steps, warmup=a.warmup, ms_per_step=float(ms), higher_is_better=True, scaling="strong",
                    dtype="f64", data="synthetic",
                    config=dict(workload="C5 scaling stress: %d potentials in total x 500 energies x lmax 12, phsh_wil + "
                                         "conphas, potentials sharded over %d GPU(s), no collective" % (a.potentials, world)),
                    checksum_sum_abs_delta=float(chk))
        os.write(real_stdout, (json.dumps(line) + "\n").encode())
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
