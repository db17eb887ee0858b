#!/usr/bin/env python
"""Secondary measurements (not the bench.py contract line): the other BASELINE.json configurations through the
device-tensor API, CUDA-event timed after warm-up, with the CPU oracle timed on a bounded sample beside them.

  python scripts/bench_paths.py > profiles/rN_secondary.jsonl
"""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402
import torch  # noqa: E402
from phaseshifts_b200 import api, workloads  # noqa: E402
from oracle import oracle as orc  # noqa: E402


def timed(fn, reps=5, warm=3):
    for _ in range(warm):
        out = fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        out = fn()
    e1.record()
    torch.cuda.synchronize()
    return out, e0.elapsed_time(e1) / reps


def threads():
    return max(1, len(os.sched_getaffinity(0)))


def main():
    nt = threads()
    res = []
    # C1: Re fixture (launch-bound)
    w = workloads.c1_re_fixture()
    p = torch.from_numpy(w["pot"]).cuda()
    g = w["grid"]
    out, ms = timed(lambda: api.phsh_rel_batch(p, w["jri"], g["e_l"], w["lmax"], energies_l0=g["e_l0"], unwrap=True))
    t0 = time.perf_counter()
    ref = orc.rel_batch(w["pot"], w["jri"], g["e_l0"], g["e_l"], lsm=w["lmax"], nthreads=nt)["jf"]
    cpu = time.perf_counter() - t0
    ref = np.stack([orc.conphas(x) for x in ref])
    res.append(dict(config="C1 Re(0001) fixture: 2 blocks x 57 energies x 13 l, phsh_rel + conphas", units=out.numel(),
                    gpu_ms=ms, gpu_units_per_s=out.numel() / ms * 1e3, cpu_units_per_s=ref.size / cpu, cpu_threads=nt,
                    max_abs_err=float(np.abs(out.cpu().numpy() - ref).max())))
    # C2: cav
    w = workloads.c2_oxide()
    V, RX = torch.from_numpy(w["V"]).cuda(), torch.from_numpy(w["RX"]).cuda()
    out, ms = timed(lambda: api.phsh_cav_batch(V, RX, w["ntab"], w["rmt"], w["e_ryd"], NL=16, unwrap=True))
    t0 = time.perf_counter()
    ref, _ = orc.cav_batch(w["V"], w["RX"], w["ntab"], w["rmt"], w["e_ryd"], NL=16, nthreads=nt)
    cpu = time.perf_counter() - t0
    ref = np.stack([orc.conphas(x) for x in ref])
    res.append(dict(config="C2 oxide slab: 4 species x 1999 energies x 16 l, phsh_cav + conphas", units=out.numel(),
                    gpu_ms=ms, gpu_units_per_s=out.numel() / ms * 1e3, cpu_units_per_s=ref.size / cpu, cpu_threads=nt,
                    max_abs_err=float(np.abs(out.cpu().numpy() - ref).max())))
    # C3: Au fine grid
    w = workloads.c3_gold()
    p = torch.from_numpy(w["pot"]).cuda()
    out, ms = timed(lambda: api.phsh_rel_batch(p, w["jri"], w["e_ryd"], w["lmax"], dx=w["dx"], unwrap=True))
    sel = slice(0, 400)
    t0 = time.perf_counter()
    r = orc.rel_batch(w["pot"], w["jri"], w["e_ryd"][sel], lsm=w["lmax"], dx=w["dx"], nthreads=nt)
    cpu = time.perf_counter() - t0
    ref = orc.conphas(r["jf"][0])
    res.append(dict(config="C3 Au Z=79: 1 potential x 2000 energies x 19 l, dx=1/64 (J=683), phsh_rel + conphas",
                    units=out.numel(), gpu_ms=ms, gpu_units_per_s=out.numel() / ms * 1e3,
                    cpu_units_per_s=ref.size / cpu, cpu_threads=nt, cpu_sample="first 400 energies",
                    max_abs_err=float(np.abs(out[0, sel].cpu().numpy() - ref).max())))
    # C5: wil
    w = workloads.c5_wil()
    rs, zs = torch.from_numpy(w["rs"]).cuda(), torch.from_numpy(w["zs"]).cuda()
    out, ms = timed(lambda: api.phsh_wil_batch(rs, zs, w["nrows"], w["Z"], w["RT"], w["e"], NL=13, NR=101, unwrap=True),
                    reps=3, warm=2)
    n = 16 * nt
    t0 = time.perf_counter()
    r = orc.wil_batch(w["rs"][:n], w["zs"][:n], w["nrows"][:n], w["Z"][:n], w["RT"][:n], w["e"], NL=13, NR=101,
                      nthreads=nt)["del"]
    cpu = time.perf_counter() - t0
    ref = np.stack([orc.conphas(x) for x in r])
    res.append(dict(config="C5 scaling stress: 65536 potentials x 500 energies x 13 l, phsh_wil + conphas",
                    units=out.numel(), gpu_ms=ms, gpu_units_per_s=out.numel() / ms * 1e3,
                    cpu_units_per_s=ref.size / cpu, cpu_threads=nt, cpu_sample="first %d potentials" % n,
                    max_abs_err=float(np.abs(out[:n].cpu().numpy() - ref).max())))
    # C6: the producer of the sweep's potentials -- muffin-tin potential step for a batch of perturbed cells
    from oracle import cavpot_io as cio
    w = workloads.c6_cavpot()
    b = w["rela"]
    rho0, nx0 = api.cavpot_rela_density(b["rmin"], b["rmax"], b["rs"], b["iprint"])
    rho, nxs = rho0[None, :], np.array([nx0], dtype=np.int32)
    pk = api.cavpot_pack(w["structures"])
    out = api.cavpot_batch(pk, rho, nxs)
    t0 = time.perf_counter()
    reps = 5
    for _ in range(reps):
        out = api.cavpot_batch(pk, rho, nxs)
    ms = (time.perf_counter() - t0) / reps * 1e3   # host API on packed arrays: H2D, one launch, D2H (wall clock)
    n = 16 * nt
    rho2, nx2 = np.stack([rho0, rho0]), np.array([nx0, nx0], dtype=np.int32)

    def one(s):   # the oracle handles one structure per call; ctypes releases the GIL, so a thread pool uses every core
        st = w["structures"][s]
        c = dict(spa=st["spa"], rc=st["rc"], nr=2, nrr=np.array(st["nrr"], dtype=np.int32), z=np.array(st["z"]),
                 zc=np.array(st["zc"]), rmt=np.array(st["rmt"]), rk=st["rk"], nform=2, alpha=st["alpha"], nh=st["nh"])
        r = cio.cavpot(c, rho2, nx2, real32=False)
        k = r["npts"][0]
        return float(np.abs(out["pot"][2 * s, :k] - r["v"][0, :k]).max())

    from concurrent.futures import ThreadPoolExecutor
    one(0)
    t0 = time.perf_counter()
    one(1)
    cpu1 = time.perf_counter() - t0
    t0 = time.perf_counter()
    with ThreadPoolExecutor(max_workers=nt) as ex:
        err = max(ex.map(one, range(n)))
    cpu = time.perf_counter() - t0
    S = len(w["structures"])
    res.append(dict(config="C6 producer: 4096 perturbed Re(0001) cells (2 atom types, NH=10) through cavpot, host API",
                    units=S, gpu_ms=ms, gpu_units_per_s=S / ms * 1e3, cpu_units_per_s=n / cpu, cpu_threads=nt,
                    cpu_sample="first %d structures on %d threads" % (n, nt), cpu_single_thread_units_per_s=1.0 / cpu1,
                    max_abs_err=err, unit="structures"))
    # C7: the atomic step in front of the producer -- hartfock on the reference's own Re input, then a batch of atoms
    from oracle import hartfock_io as hio
    re_in = os.path.join(ROOT, "tests", "golden", "Re0001", "atorb_Re")
    spec = [v for c, v in hio.parse_atorb(re_in) if c == "a"][0]
    atom = dict(z=75.0, nr=1000, rel=1.0, orbitals=spec["orbitals"])
    api.hartfock_batch([atom])
    t0 = time.perf_counter()
    got = api.hartfock_batch([atom])[0]
    ms1 = (time.perf_counter() - t0) * 1e3
    t0 = time.perf_counter()
    want = hio.scf(75.0, 1000, 1.0, 0.0, 0, spec)
    cpu1 = time.perf_counter() - t0
    nb = 148
    t0 = time.perf_counter()
    api.hartfock_batch([atom] * nb)
    msb = (time.perf_counter() - t0) * 1e3
    res.append(dict(config="C7 atomic step: hartfock for Re (22 orbitals x 1000 points, 17 SCF iterations), host API",
                    units=1, gpu_ms=ms1, gpu_units_per_s=1e3 / ms1, cpu_units_per_s=1.0 / cpu1, cpu_threads=1,
                    max_abs_err=float(np.abs(got["rho"] - want["rho"]).max()), bit_identical=bool(np.array_equal(got["rho"], want["rho"])),
                    batch=dict(atoms=nb, gpu_ms=msb, atoms_per_s=nb / msb * 1e3), unit="atoms"))
    for r in res:
        r["speedup_vs_cpu_threads"] = r["gpu_units_per_s"] / r["cpu_units_per_s"]
        print(json.dumps(r), flush=True)


if __name__ == "__main__":
    main()
