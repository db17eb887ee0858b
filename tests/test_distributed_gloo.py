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


def _cavpot_worker(rank, world, port, S, out_dir):
    sys.path.insert(0, ROOT)
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world),
                      LOCAL_RANK=str(rank))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from phaseshifts_b200 import distributed as D
    structs, rho, nx = _cavpot_case(S)
    full = D.cavpot_sweep(structs, rho, nx, compute=_oracle_cavpot)
    local, (a, b) = D.cavpot_sweep(structs, rho, nx, compute=_oracle_cavpot, gather=False)
    np.savez(os.path.join(out_dir, "cav_%d.npz" % rank), mtz=full["mtz"], npts=full["npts"], pot=full["pot"],
             type_off=full["type_off"], lmtz=local["mtz"], rng=np.array([a, b]))
    dist.barrier()
    dist.destroy_process_group()


def _cavpot_case(S):
    """S ragged structures (1 or 2 atom types) with a synthetic density; NH small to keep the oracle quick."""
    rng = np.random.default_rng(3)
    x = np.exp(-8.8 + 0.05 * np.arange(250))
    rho = (4 * 8.0 ** 3 * x * x * np.exp(-2 * 2.0 * x))[None, :] * 10.0   # 4 pi r^2 times a 1s-like density, 10 electrons
    nx = np.array([250], dtype=np.int32)
    structs = []
    for s in range(S):
        two = s % 3 != 0
        structs.append(dict(spa=5.0 + 0.1 * s, rc=np.eye(3), nrr=[1, 1] if two else [1], z=[10.0, 10.0] if two else [10.0],
                            zc=[0.0, 0.0] if two else [0.0], rmt=[2.0, 2.1] if two else [2.3],
                            rk=[[0, 0, 0], [0.5, 0.5, 0.5]] if two else [[0, 0, 0]], dens=[0, 0] if two else [0],
                            alpha=0.7, nh=3, nform=2))
    return structs, rho, nx


def _oracle_cavpot(structs, rho, nx):
    from oracle import cavpot_io as cio
    mtz, npts, pot = [], [], []
    for st in structs:
        nr = len(st["nrr"])
        c = dict(spa=st["spa"], rc=np.asarray(st["rc"], dtype=float), nr=nr, nrr=np.asarray(st["nrr"], dtype=np.int32),
                 z=np.asarray(st["z"], dtype=float), zc=np.asarray(st["zc"], dtype=float), rmt=np.asarray(st["rmt"], dtype=float),
                 rk=np.asarray(st["rk"], dtype=float), nform=st["nform"], alpha=st["alpha"], nh=st["nh"])
        d = np.asarray(st["dens"])
        r = cio.cavpot(c, np.asarray(rho)[d], np.asarray(nx)[d], real32=False)
        mtz.append(r["mtz"])
        npts += list(r["npts"])
        pot += [r["v"][i, :421] for i in range(nr)]
    return dict(mtz=np.array(mtz), npts=np.array(npts, dtype=np.int32), pot=np.array(pot).reshape(-1, 421))


@pytest.mark.parametrize("S", [5, 2])
def test_two_rank_cavpot_sweep_equals_single_rank(tmp_path, S):
    world = 2
    mp.spawn(_cavpot_worker, args=(world, _free_port(), S, str(tmp_path)), nprocs=world, join=True)
    structs, rho, nx = _cavpot_case(S)
    ref = _oracle_cavpot(structs, rho, nx)
    for r in range(world):
        z = np.load(tmp_path / ("cav_%d.npz" % r))
        assert np.array_equal(z["mtz"], ref["mtz"]) and np.array_equal(z["npts"], ref["npts"])
        assert np.array_equal(z["pot"], ref["pot"]) and z["type_off"][-1] == ref["pot"].shape[0]
        a, b = z["rng"]
        assert np.array_equal(z["lmtz"], ref["mtz"][a:b])


def test_gather_without_process_group_raises(monkeypatch):
    """WORLD_SIZE>1 in the environment but no init_process_group: a gather must fail loudly instead of returning one
    rank's shard as the whole array (and cavpot_sweep must not crash on a missing group object)."""
    from phaseshifts_b200 import distributed as D
    monkeypatch.setenv("RANK", "0")
    monkeypatch.setenv("WORLD_SIZE", "2")
    pot = np.ones((4, 16))
    calls = []

    def compute(p_, j_, e_, l_, **kw):
        calls.append(p_.shape[0])
        return np.zeros((p_.shape[0], 3, l_ + 1))

    with pytest.raises(RuntimeError, match="not initialised"):
        D.phsh_rel_sweep(pot, 16, np.array([1.0, 2.0, 3.0]), 2, compute=compute)
    with pytest.raises(RuntimeError, match="not initialised"):
        D.cavpot_sweep([dict(nrr=[1])] * 4, None, None, compute=lambda *a: None)
    assert calls == []
    # without a gather the env-derived shard is legitimate
    local, (a, b) = D.phsh_rel_sweep(pot, 16, np.array([1.0, 2.0, 3.0]), 2, compute=compute, gather=False)
    assert (a, b) == (0, 2) and local.shape == (2, 3, 3)
