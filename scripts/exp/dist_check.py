"""NCCL check of the sharded sweeps on a multi-GPU box (torchrun): every rank computes its shard of a small relativistic
sweep and of a batch of structures on its own GPU, the blocks are all-gathered over NCCL, and the gathered arrays must be
bit-identical to the same sweep computed on one GPU.  Prints one line per rank 0 check.
  python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 scripts/exp/dist_check.py"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np
import torch
import torch.distributed as dist
from phaseshifts_b200 import api, distributed, workloads

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
w = workloads.c4_sweep(P=37, ne=96, lmax=7, seed=4)  # 37 potentials: uneven shards
g = w["grid"]
full = distributed.phsh_rel_sweep(w["pot"], w["jri"], g["e_l"], 7, energies_l0=g["e_l0"], unwrap=True, gather=True)
one = api.phsh_rel_batch(w["pot"], w["jri"], g["e_l"], 7, energies_l0=g["e_l0"], unwrap=True, device=local)
ok_rel = full.shape == one.shape and np.array_equal(full, np.asarray(one))
c6 = workloads.c6_cavpot(S=21)
b = c6["rela"]
rho0, nx0 = api.cavpot_rela_density(b["rmin"], b["rmax"], b["rs"], b["iprint"], device=local)
rho, nx = rho0[None, :], np.array([nx0], dtype=np.int32)
cg = distributed.cavpot_sweep(c6["structures"], rho, nx, gather=True)
c1 = api.cavpot_batch(c6["structures"], rho, nx, device=local)
ok_cav = np.array_equal(cg["pot"], c1["pot"][:, :421]) and np.array_equal(cg["mtz"], c1["mtz"]) and np.array_equal(cg["npts"], c1["npts"])
flags = torch.tensor([int(ok_rel), int(ok_cav)], device="cuda")
dist.all_reduce(flags, op=dist.ReduceOp.MIN)
if rank == 0:
    print("dist_check world=%d backend=%s: phsh_rel_sweep gathered == single GPU: %s; cavpot_sweep gathered == single GPU: %s"
          % (world, dist.get_backend(), bool(flags[0].item()), bool(flags[1].item())))
dist.destroy_process_group()
sys.exit(0 if flags.min().item() == 1 else 1)
