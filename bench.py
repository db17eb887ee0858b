#!/usr/bin/env python
"""bench.py -- phase shifts delta_l(E) per second on the LEED-IV sweep (BASELINE.json config C4).

One "step" = one pass of the hot path over one batch: P potentials x 1000 energies x (lmax+1) l through
DLGKAP (both kappa) + SBFIT + the LVOR kappa-average + conphas unwrapping, i.e. the work of
``libphsh.phsh_rel`` + ``Conphas.calculate`` for P atom blocks.  Weak scaling: every rank owns its own
slice of P potentials of one global sweep (no data-path collective; results stay / are gathered per rank).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference]
  python -m torch.distributed.run --nproc-per-node N ... bench.py --gpus N ...

Prints ONE JSON line (see the keys at the bottom).  ``--impl reference`` times the CPU restatement of the
reference algorithm (oracle/, all host threads) on a bounded sample of the same workload.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

METRIC = "phase_shifts_per_sec"
UNIT = "delta_l(E)/s"


def host_threads():
    try:
        return max(1, len(os.sched_getaffinity(0)))
    except AttributeError:  # pragma: no cover
        return os.cpu_count() or 1


def make_workload(P, ne, lmax, first):
    from phaseshifts_b200 import workloads
    return workloads.c4_sweep(P=P, ne=ne, lmax=lmax, seed=4, first=first)


def cpu_reference_rate(wl, n_pot, threads):
    """delta/s of the CPU restatement (oracle) on the first n_pot potentials of the workload."""
    from oracle import oracle as orc
    orc.build()
    g = wl["grid"]
    t0 = time.perf_counter()
    r = orc.rel_batch(wl["pot"][:n_pot], wl["jri"][:n_pot], g["e_l0"], g["e_l"], lsm=wl["lmax"], nthreads=threads)
    jf = r["jf"]
    for p in range(n_pot):  # Conphas.calculate on every block (cheap)
        orc.conphas(jf[p])
    dt = time.perf_counter() - t0
    return jf.size / dt, dt, r


class ClockSampler:
    """nvidia-smi clock / throttle-reason sampler running during the timed region."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index = index
        self.f = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        self.p = None

    def start(self):
        try:
            self.p = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                       "--format=csv,noheader,nounits", "-lms", "100"], stdout=self.f,
                                      stderr=subprocess.DEVNULL)
        except OSError:
            self.p = None

    def stop(self):
        if self.p is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.p.terminate()
        try:
            self.p.wait(timeout=5)
        except subprocess.TimeoutExpired:  # pragma: no cover
            self.p.kill()
        self.f.flush()
        self.f.seek(0)
        sm, mx, pw, reasons = [], [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.f.read().splitlines():
            c = [x.strip() for x in ln.split(",")]
            if len(c) < 9:
                continue
            try:
                sm.append(float(c[1])); mx.append(float(c[2])); pw.append(float(c[3]))
            except ValueError:
                continue
            for n, v in zip(names, c[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        os.unlink(self.f.name)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        return {"sm_mhz": statistics.median(sm), "sm_max_mhz": max(mx), "power_w_max": max(pw), "samples": len(sm),
                "reasons": sorted(reasons)}


def run_reference(args, rank, world, out_fd):
    """--impl reference: the reference algorithm on the host cores (oracle port; the Fortran cannot be built here)."""
    if rank != 0:
        return
    threads = host_threads()
    n_pot = max(1, min(args.potentials, args.ref_potentials or 24 * threads))
    wl = make_workload(n_pot, args.energies, args.lmax, 0)
    rates, times = [], []
    for i in range(args.warmup + args.steps):
        rate, dt, _ = cpu_reference_rate(wl, n_pot, threads)
        if i >= args.warmup:
            rates.append(rate); times.append(dt)
    value = float(np.sum([n_pot * args.energies * (args.lmax + 1)] * len(times)) / np.sum(times))
    sample = "%d of %d potentials x %d energies x %d l per step" % (n_pot, args.potentials, args.energies, args.lmax + 1)
    out = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * float(np.mean(times)), "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": workload_config(args, world),
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    emit(out, out_fd)


def workload_config(args, world):
    return {"workload": "C4 LEED-IV sweep: %d perturbed muffin-tin potentials (R_mt, V0 jitter) per GPU x %d energies "
                        "(20..%d eV, 1 eV) x lmax=%d, phsh_rel (both kappa) + conphas" %
                        (args.potentials, args.energies, 19 + args.energies, args.lmax),
            "potentials_per_gpu": args.potentials, "energies": args.energies, "lmax": args.lmax,
            "radial_points": "316..325 (Waber grid dx=1/32)", "parallelism": "potentials sharded over %d GPU(s), no collective" % world,
            "l2": "each step writes %.0f MB of results (> 126 MB L2) and a 256 MB buffer is rewritten between steps" %
                  (args.potentials * args.energies * (args.lmax + 1) * 8 * 3 / 1e6),
            "arithmetic": "strict IEEE FP64, bit-identical to the CPU restatement (no FMA contraction)" if not args.fast
            else "FAST (FMA-contracted) -- not the parity-bearing mode"}


def emit(obj, fd):
    """Write the ONE JSON line to the real stdout (fd saved before libraries such as NCCL could print to it)."""
    os.write(fd, (json.dumps(obj) + "\n").encode())


def main():
    # everything libraries print to stdout (e.g. the "NCCL version" banner) goes to stderr; only the JSON line
    # reaches the real stdout
    sys.stdout.flush()
    real_stdout = os.dup(1)
    os.dup2(2, 1)
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--potentials", type=int, default=4096, help="potentials per GPU (C4: 4096)")
    ap.add_argument("--energies", type=int, default=1000)
    ap.add_argument("--lmax", type=int, default=15)
    ap.add_argument("--fast", action="store_true", help="FMA-contracted integrator (not bit-faithful)")
    ap.add_argument("--ref-potentials", type=int, default=0, help="size of the CPU sample (default 24 x host threads, ~15-20 s)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank, world, real_stdout)
        return

    import torch
    import torch.distributed as dist
    from phaseshifts_b200 import api, _lib
    import ctypes as C

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device -- the B200 path has no CPU fallback")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    def max_over_ranks(x):
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    P, NE, L1 = args.potentials, args.energies, args.lmax + 1
    wl = make_workload(P, NE, args.lmax, first=rank * P)
    g = wl["grid"]
    units_per_step = P * NE * L1

    # ---- host (pinned) and device copies of the inputs -------------------------------------------------
    pot_h = torch.from_numpy(wl["pot"]).pin_memory()
    jri_h = torch.from_numpy(wl["jri"]).pin_memory()
    e0_h = torch.from_numpy(g["e_l0"]).pin_memory()
    e1_h = torch.from_numpy(g["e_l"]).pin_memory()
    out_h = torch.empty((P, NE, L1), dtype=torch.float64).pin_memory()
    pot_d, jri_d = pot_h.to(dev), jri_h.to(dev)
    e0_d, e1_d = e0_h.to(dev), e1_h.to(dev)
    delta_d = torch.empty((P, NE, L1), dtype=torch.float64, device=dev)
    raw_d = torch.empty((P, L1, NE, 2), dtype=torch.float64, device=dev)
    flush = torch.empty(256 * 1024 * 1024 // 8, dtype=torch.float64, device=dev)
    lib = _lib.load()
    opts = _lib.RelOpts(api.XS_DEFAULT, api.DX_DEFAULT, args.lmax, _lib.UNWRAP | (_lib.FAST if args.fast else 0))
    stream = torch.cuda.current_stream(dev)

    def ptr(t):
        return C.c_void_p(t.data_ptr())

    def step():  # the public device-tensor call a user makes (kernel A + kernel B on torch's stream)
        api.phsh_rel_batch(pot_d, jri_d, e1_d, args.lmax, energies_l0=e0_d, unwrap=True, fast=args.fast, out=delta_d)

    def integrate_only(cnt=None):
        _lib.check(lib.phsh_b200_rel_integrate_device(ptr(pot_d), pot_d.shape[1], ptr(jri_d), P, ptr(e0_d), ptr(e1_d), 0,
                                                      NE, C.byref(opts), ptr(raw_d), ptr(cnt) if cnt is not None else None,
                                                      C.c_void_p(stream.cuda_stream)))

    # ---- device-resident throughput ("value") ---------------------------------------------------------
    for _ in range(args.warmup):
        step()
        flush.fill_(1.0)
    barrier()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    ev0.record(stream)
    for _ in range(args.steps):
        step()
        flush.fill_(1.0)
    ev1.record(stream)
    barrier()
    ms_total = max_over_ranks(ev0.elapsed_time(ev1))
    clocks = sampler.stop() if rank == 0 else None
    ms_per_step = ms_total / args.steps
    value = world * units_per_step / (ms_per_step * 1e-3)

    # ---- roofline of the dominant kernel (rel_integrate_kernel) -------------------------------------
    cnt = torch.zeros(4, dtype=torch.int64, device=dev)
    integrate_only(cnt)
    torch.cuda.synchronize(dev)
    c = cnt.cpu().numpy()
    flops_per_launch = 330.0 * c[0] + 23.0 * c[1] + 28.0 * c[2]
    k0, k1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    kms = []
    for _ in range(max(3, args.steps)):
        flush.fill_(1.0)
        k0.record(stream)
        integrate_only()
        k1.record(stream)
        torch.cuda.synchronize(dev)
        kms.append(k0.elapsed_time(k1))
    kernel_ms = float(np.mean(kms))
    peak_tf, peak_mhz = api.fp64_peak(local)
    achieved_tf = flops_per_launch / (kernel_ms * 1e-3) / 1e12
    # DRAM traffic per launch and FP64-pipe utilisation come from the committed ncu --set full capture of this very
    # launch (profiles/); they are only quoted when the configuration is the captured one
    traffic, ncu_pipe = None, None
    try:
        with open(os.path.join(ROOT, "profiles", "r1_rel_integrate_ncu_c4.json")) as f:
            cap = json.load(f)
        if (P, NE, args.lmax, args.fast) == (4096, 1000, 15, False):
            traffic = cap["traffic_bytes_per_launch"]
            ncu_pipe = cap["fp64_pipe_pct_of_peak_sustained_active"]
    except (OSError, KeyError, ValueError):
        pass

    # ---- end to end through the host-buffer C ABI (H2D + kernels + D2H inside the timed region) ------
    pot_np, jri_np, e0_np, e1_np, out_np = pot_h.numpy(), jri_h.numpy(), e0_h.numpy(), e1_h.numpy(), out_h.numpy()

    def e2e_step():
        api.phsh_rel_batch(pot_np, jri_np, e1_np, args.lmax, energies_l0=e0_np, unwrap=True, fast=args.fast, device=local,
                           out=out_np)

    for _ in range(2):
        e2e_step()
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        e2e_step()
    torch.cuda.synchronize(dev)
    e2e_s = max_over_ranks(time.perf_counter() - t0)
    e2e_value = world * units_per_step * args.steps / e2e_s
    h2d = pot_np.nbytes + jri_np.nbytes + e0_np.nbytes + e1_np.nbytes
    d2h = out_np.nbytes + 32
    # device result == host-path result (same kernels)
    same = bool(np.array_equal(delta_d.cpu().numpy(), out_np))

    # ---- CPU baseline on a bounded sample (rank 0, N=1 only) ----------------------------------------
    cpu = None
    parity = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        threads = host_threads()
        n_pot = max(1, min(P, args.ref_potentials or 24 * threads))
        rate, dt, r = cpu_reference_rate(wl, n_pot, threads)
        rate1, dt1, _ = cpu_reference_rate(wl, min(4, n_pot), 1)
        cpu = {"value": rate, "unit": UNIT, "cores": threads, "kind": "port",
               "sample": "%d of %d potentials x %d energies x %d l, %.1f s" % (n_pot, P, NE, L1, dt),
               "single_thread_value": rate1,
               "single_thread_sample": "%d potentials, %.1f s" % (min(4, n_pot), dt1)}
        from oracle import oracle as orc
        ref = np.stack([orc.conphas(x) for x in r["jf"]])
        parity = float(np.abs(out_np[:n_pot] - ref).max())

    # the same kernel seen against the HBM roofline (measured copy bandwidth of this pool's B200s): far from it
    hbm_peak, hbm_src = 6650.0, "fallback of B200_PROFILING.md"
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            hbm_peak, hbm_src = float(json.load(f)["hbm_gbs"]), "MEASURED_PEAKS.json"
    except (OSError, KeyError, ValueError):
        pass
    alg_bytes = pot_d.numel() * 8.0 + raw_d.numel() * 8.0
    hbm_view = {"achieved": alg_bytes / (kernel_ms * 1e-3) / 1e9, "peak": hbm_peak, "unit": "GB/s",
                "frac": alg_bytes / (kernel_ms * 1e-3) / 1e9 / hbm_peak, "peak_source": hbm_src,
                "algorithmic_bytes_per_launch": alg_bytes}

    if world > 1:
        dist.destroy_process_group()
    if rank != 0:
        return
    out = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms_per_step, "units_per_step": world * units_per_step, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None,
        "dtype": "f64", "data": "synthetic", "config": workload_config(args, world),
        "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h),
                "api": "phaseshifts_b200.api.phsh_rel_batch(host arrays) -> phsh_b200_rel_batch_host",
                "matches_device_path": same},
        "gpu_launches": 2 * args.steps,
        "clocks": clocks,
        "roofline": {"bound": "fp64", "achieved": achieved_tf, "peak": peak_tf, "unit": "TFLOP/s",
                     "frac": achieved_tf / peak_tf if peak_tf else None, "traffic": traffic,
                     "traffic_unit": "bytes per launch (dram__bytes_read.sum + dram__bytes_write.sum; algorithmic: "
                                     "%.3g in + %.3g per-kappa scratch out)" % (pot_d.numel() * 8.0, raw_d.numel() * 8.0),
                     "ncu_fp64_pipe_pct_of_peak_sustained_active": ncu_pipe,
                     "hbm_view": hbm_view,
                     "kernel": "rel_integrate_kernel", "kernel_ms": kernel_ms,
                     "flops_per_launch": flops_per_launch,
                     "corrector_evals_per_step": float(c[2]) / max(float(c[1]), 1.0),
                     "peak_source": "measured live: 8 independent DFMA chains/thread, all SMs (implies %.0f MHz); "
                                    "MEASURED_PEAKS.json has no FP64 figure" % peak_mhz,
                     "note": "flops = as-written FP64 add/mul count of the reference (330/call + 23/Milne step + "
                             "28/corrector evaluation); the strict kernel issues one FP64 instruction per add/mul, so "
                             "its ceiling is 0.5 of the DFMA peak"},
        "cpu_baseline": cpu,
        "parity_max_abs_err_rad": parity,
    }
    emit(out, real_stdout)


if __name__ == "__main__":
    main()
