"""Multi-GPU sharding of a sweep: one process per GPU, ``torch.distributed`` for the plumbing only.

Every (potential, l, kappa) integration is independent and the only coupling runs along the energy axis
of one (potential, l) (LVOR average, conphas), so the sweep is partitioned by potential index: rank r of
W owns the contiguous block ``shard_range(P, r, W)`` and there is NO data-path collective.  The optional
gather at the end (``all_gather`` of equally padded blocks; NCCL over NVLink for device tensors, gloo for
host tensors) only serves a consumer that wants the whole array in one place.
"""
from __future__ import annotations

import os

import numpy as np


def shard_range(P: int, rank: int, world: int):
    """Contiguous block partition of ``range(P)``; the first ``P % world`` ranks get one extra item."""
    if world <= 0 or not 0 <= rank < world:
        raise ValueError("bad rank/world: %r/%r" % (rank, world))
    base, extra = divmod(int(P), world)
    start = rank * base + min(rank, extra)
    return start, start + base + (1 if rank < extra else 0)


def _dist():
    import torch.distributed as dist
    return dist if dist.is_available() and dist.is_initialized() else None


def rank_world():
    d = _dist()
    if d is not None:
        return d.get_rank(), d.get_world_size()
    return int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))


def _require_group(world: int, what: str):
    """A gather over world>1 ranks needs an initialised process group: the RANK/WORLD_SIZE environment alone (e.g. under
    torchrun before ``init_process_group``) would otherwise hand back one rank's shard as if it were the whole array."""
    d = _dist()
    if world > 1 and d is None:
        raise RuntimeError("%s: WORLD_SIZE=%d but torch.distributed is not initialised -- call "
                           "torch.distributed.init_process_group first, or pass gather=False" % (what, world))
    return d


def gather_blocks(local, P: int, group=None):
    """All-gather the per-rank blocks ``local`` [P_r, ...] (torch tensor, cpu or cuda) into [P, ...] on every rank."""
    import torch
    d = _dist()
    if d is None:
        if int(os.environ.get("WORLD_SIZE", "1")) > 1:
            _require_group(int(os.environ["WORLD_SIZE"]), "gather_blocks")
        return local
    rank, world = d.get_rank(group), d.get_world_size(group)
    sizes = [shard_range(P, r, world)[1] - shard_range(P, r, world)[0] for r in range(world)]
    if local.shape[0] != sizes[rank]:
        raise ValueError("rank %d holds %d rows, expected %d" % (rank, local.shape[0], sizes[rank]))
    pad = max(sizes)
    buf = local.new_zeros((pad,) + tuple(local.shape[1:]))
    buf[: local.shape[0]] = local
    parts = [torch.empty_like(buf) for _ in range(world)]
    d.all_gather(parts, buf, group=group)
    return torch.cat([p[:n] for p, n in zip(parts, sizes)], dim=0)


def phsh_rel_sweep(pot, jri, energies_ryd, lmax, *, gather=True, compute=None, group=None, **kwargs):
    """Relativistic phase shifts of ``pot`` [P, J] sharded over the ranks of the current process group.

    Each rank computes ``shard_range(P, rank, world)`` on its own GPU (``LOCAL_RANK``) through
    ``api.phsh_rel_batch`` (or ``compute``, same signature -- the CPU tests inject the oracle there) and, if
    ``gather``, every rank returns the full [P, NE, lmax+1] array; otherwise ``(local_block, (start, stop))``.
    """
    import torch
    rank, world = rank_world()
    if gather:
        _require_group(world, "phsh_rel_sweep(gather=True)")
    pot = np.asarray(pot)
    P = pot.shape[0]
    jri = np.broadcast_to(np.asarray(jri, dtype=np.int32), (P,))
    a, b = shard_range(P, rank, world)
    e = np.asarray(energies_ryd)
    e_loc = e[a:b] if e.ndim == 2 else e
    if "energies_l0" in kwargs and kwargs["energies_l0"] is not None and np.asarray(kwargs["energies_l0"]).ndim == 2:
        kwargs = dict(kwargs, energies_l0=np.asarray(kwargs["energies_l0"])[a:b])
    if compute is None:
        from . import api
        local_rank = int(os.environ.get("LOCAL_RANK", rank))

        def compute(p_, j_, e_, l_, **kw):
            return api.phsh_rel_batch(p_, j_, e_, l_, device=local_rank, **kw)
    NE = e.shape[-1]
    if b > a:
        local = np.asarray(compute(pot[a:b], jri[a:b], e_loc, lmax, **kwargs))
    else:
        local = np.empty((0, NE, lmax + 1))
    if not gather or world == 1:
        return local if world == 1 and gather else (local, (a, b))
    t = torch.from_numpy(np.ascontiguousarray(local))
    d = _dist()
    if d is not None and d.get_backend(group) == "nccl":
        t = t.cuda(int(os.environ.get("LOCAL_RANK", rank)))
    return gather_blocks(t, P, group=group).cpu().numpy()


def _gather_ragged(local, sizes, group=None):
    """All-gather blocks whose row counts ``sizes[rank]`` every rank can work out from the (replicated) inputs."""
    import torch
    d = _require_group(len(sizes), "_gather_ragged")
    if d is None:  # a single rank: nothing to exchange
        return local
    rank, world = d.get_rank(group), d.get_world_size(group)
    if local.shape[0] != sizes[rank]:
        raise ValueError("rank %d holds %d rows, expected %d" % (rank, local.shape[0], sizes[rank]))
    buf = local.new_zeros((max(max(sizes), 1),) + tuple(local.shape[1:]))
    buf[: local.shape[0]] = local
    parts = [torch.empty_like(buf) for _ in range(world)]
    d.all_gather(parts, buf, group=group)
    return torch.cat([p[:n] for p, n in zip(parts, sizes)], dim=0)


def cavpot_sweep(structures, rho, nx, *, gather=True, compute=None, group=None):
    """Muffin-tin potentials of a list of structures sharded over the ranks of the current process group.

    Structures are independent, so rank r of W owns ``shard_range(S, r, W)`` and there is no data-path collective.
    ``compute(structures, rho, nx)`` defaults to ``api.cavpot_batch`` on GPU ``LOCAL_RANK`` (the CPU tests inject the
    oracle).  With ``gather`` every rank returns ``dict(mtz [S], npts [T], pot [T][421], type_off [S+1])`` for the whole
    list (the rows of atom types are ragged per structure; every rank derives the block sizes from the replicated
    input); otherwise ``(local_result, (start, stop))``.
    """
    import torch
    rank, world = rank_world()
    if gather:
        _require_group(world, "cavpot_sweep(gather=True)")
    S = len(structures)
    a, b = shard_range(S, rank, world)
    if compute is None:
        from . import api
        local_rank = int(os.environ.get("LOCAL_RANK", rank))

        def compute(st, rho_, nx_):
            return api.cavpot_batch(st, rho_, nx_, device=local_rank)
    ntypes = np.array([len(np.asarray(st["nrr"]).reshape(-1)) for st in structures], dtype=np.int64)
    type_off = np.concatenate([[0], np.cumsum(ntypes)]).astype(np.int32)
    if b > a:
        res = compute(list(structures[a:b]), rho, nx)
        local = dict(mtz=np.asarray(res["mtz"], dtype=np.float64), npts=np.asarray(res["npts"], dtype=np.int32),
                     pot=np.asarray(res["pot"], dtype=np.float64)[:, :421])
    else:
        local = dict(mtz=np.zeros(0), npts=np.zeros(0, dtype=np.int32), pot=np.zeros((0, 421)))
    if not gather or world == 1:
        if world == 1 and gather:
            return dict(local, type_off=type_off)
        return local, (a, b)
    d = _dist()
    dev = None
    if d is not None and d.get_backend(group) == "nccl":
        dev = int(os.environ.get("LOCAL_RANK", rank))
    ranges = [shard_range(S, r, world) for r in range(world)]
    s_sizes = [hi - lo for lo, hi in ranges]
    t_sizes = [int(type_off[hi] - type_off[lo]) for lo, hi in ranges]
    out = {}
    for key, sizes in (("mtz", s_sizes), ("npts", t_sizes), ("pot", t_sizes)):
        t = torch.from_numpy(np.ascontiguousarray(local[key]))
        if dev is not None:
            t = t.cuda(dev)
        out[key] = _gather_ragged(t, sizes, group=group).cpu().numpy()
    out["type_off"] = type_off
    return out


def phsh_rel_batch_multi(pot, jri, energies_ryd, lmax, *, devices=None, **kwargs):
    """Single-process variant: shard ``pot`` [P, J] (host array) over several GPUs of this process.

    One host thread per device calls the re-entrant host-buffer entry point (ctypes releases the GIL, every call
    owns its streams), each writing its slice of one pinned-or-pageable result array -- the "gather to host" of
    the sweep.  ``devices`` defaults to every visible GPU.  Same keyword arguments as ``api.phsh_rel_batch``.
    """
    import threading
    from . import _lib, api
    pot = np.ascontiguousarray(pot, dtype=np.float64)
    P = pot.shape[0]
    jri = np.ascontiguousarray(np.broadcast_to(np.asarray(jri, dtype=np.int32), (P,)))
    e = np.ascontiguousarray(energies_ryd, dtype=np.float64)
    e0 = kwargs.pop("energies_l0", None)
    e0 = None if e0 is None else np.ascontiguousarray(e0, dtype=np.float64)
    if devices is None:
        devices = list(range(max(1, _lib.device_count())))
    out = kwargs.pop("out", None)
    if out is None:
        out = np.empty((P, e.shape[-1], int(lmax) + 1))
    errors = []

    def work(k, dev):
        a, b = shard_range(P, k, len(devices))
        if b <= a:
            return
        try:
            api.phsh_rel_batch(pot[a:b], jri[a:b], e[a:b] if e.ndim == 2 else e, lmax,
                               energies_l0=None if e0 is None else (e0[a:b] if e0.ndim == 2 else e0),
                               device=dev, out=out[a:b], **kwargs)
        except Exception as exc:  # re-raised in the caller's thread
            errors.append(exc)

    threads = [threading.Thread(target=work, args=(k, d)) for k, d in enumerate(devices)]
    for t in threads:
        t.start()
    for t in threads:
        t.join()
    if errors:
        raise errors[0]
    return out
