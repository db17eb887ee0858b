"""world_size=2 gloo test of the multi-GPU plumbing (sharding + gather), with the CPU oracle injected as the
per-rank compute function.  CPU only."""
import os
import socket
import sys

import numpy as np
import pytest
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, P, out_dir):
    sys.path.insert(0, ROOT)
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world),
                      LOCAL_RANK=str(rank))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from oracle import oracle as orc
    from phaseshifts_b200 import distributed as D, workloads
    wl = workloads.c4_sweep(P=P, ne=12, lmax=4)

    def compute(pot, jri, e, lmax, energies_l0=None, **kw):
        return orc.rel_batch(pot, jri, energies_l0 if energies_l0 is not None else e, e, lsm=lmax)["jf"]

    full = D.phsh_rel_sweep(wl["pot"], wl["jri"], wl["grid"]["e_l"], wl["lmax"], energies_l0=wl["grid"]["e_l0"],
                            compute=compute)
    local, (a, b) = D.phsh_rel_sweep(wl["pot"], wl["jri"], wl["grid"]["e_l"], wl["lmax"],
                                     energies_l0=wl["grid"]["e_l0"], compute=compute, gather=False)
    np.save(os.path.join(out_dir, "full_%d.npy" % rank), full)
    np.save(os.path.join(out_dir, "local_%d.npy" % rank), local)
    np.save(os.path.join(out_dir, "range_%d.npy" % rank), np.array([a, b]))
    dist.barrier()
    dist.destroy_process_group()


def test_shard_range_partition():
    from phaseshifts_b200.distributed import shard_range
    for P in (0, 1, 5, 8, 4096, 65537):
        for W in (1, 2, 3, 8):
            r = [shard_range(P, k, W) for k in range(W)]
            assert r[0][0] == 0 and r[-1][1] == P
            assert all(r[k][1] == r[k + 1][0] for k in range(W - 1))
            sizes = [b - a for a, b in r]
            assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        shard_range(4, 2, 2)


@pytest.mark.parametrize("P", [5, 4])
def test_two_rank_sweep_equals_single_rank(tmp_path, P, oracle):
    from phaseshifts_b200 import workloads
    world = 2
    mp.spawn(_worker, args=(world, _free_port(), P, str(tmp_path)), nprocs=world, join=True)
    wl = workloads.c4_sweep(P=P, ne=12, lmax=4)
    ref = oracle.rel_batch(wl["pot"], wl["jri"], wl["grid"]["e_l0"], wl["grid"]["e_l"], lsm=4)["jf"]
    for r in range(world):
        assert np.array_equal(np.load(tmp_path / ("full_%d.npy" % r)), ref)
        a, b = np.load(tmp_path / ("range_%d.npy" % r))
        assert np.array_equal(np.load(tmp_path / ("local_%d.npy" % r)), ref[a:b])
