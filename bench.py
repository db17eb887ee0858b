#!/usr/bin/env python
"""bench.py -- phase shifts delta_l(E) per second on the LEED-IV sweep (BASELINE.json config C4).

One "step" = one pass of the hot path over one batch: P potentials x 1000 energies x (lmax+1) l through
DLGKAP (both kappa) + SBFIT + the LVOR kappa-average + conphas unwrapping, i.e. the work of
``libphsh.phsh_rel`` + ``Conphas.calculate`` for P atom blocks.  STRONG scaling (the north-star configuration):
the sweep is a fixed 4096 potentials, rank r of N owns ``distributed.shard_range(4096, r, N)`` of it (512 per GPU at
N=8); no data-path collective, results stay / are gathered per rank.  The weak-scaling figure (4096 potentials on
every GPU) is measured as well at N>1 and reported under ``"weak"``.

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
    wl = make_workload(n_pot, args.energies, args.lmax, 0)  # the first n_pot potentials of the same sweep
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
        "scaling": "weak" if args.weak else "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": workload_config(args, world),
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    emit(out, out_fd)


def workload_config(args, world):
    total = args.potentials * world if args.weak else args.potentials
    per = args.potentials if args.weak else -(-args.potentials // world)
    return {"workload": "C4 LEED-IV sweep: %d perturbed muffin-tin potentials (R_mt, V0 jitter) x %d energies "
                        "(20..%d eV, 1 eV) x lmax=%d, phsh_rel (both kappa) + conphas" %
                        (total, args.energies, 19 + args.energies, args.lmax),
            "potentials_total": total, "potentials_per_gpu": per, "energies": args.energies, "lmax": args.lmax,
            "radial_points": "316..325 (Waber grid dx=1/32)",
            "parallelism": "the %d potentials sharded over %d GPU(s) (%s), no collective" %
                           (total, world, "weak: fixed per GPU" if args.weak else "strong: fixed total"),
            "l2": "each step writes %.0f MB of results per GPU and a 256 MB buffer (> 126 MB L2) is rewritten between "
                  "steps" % (per * args.energies * (args.lmax + 1) * 8 * 3 / 1e6),
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
    ap.add_argument("--potentials", type=int, default=4096,
                    help="potentials of the whole sweep (C4: 4096), sharded over the GPUs; with --weak: per GPU")
    ap.add_argument("--weak", action="store_true", help="weak scaling: --potentials on EVERY GPU")
    ap.add_argument("--energies", type=int, default=1000)
    ap.add_argument("--lmax", type=int, default=15)
    ap.add_argument("--fast", action="store_true", help="FMA-contracted integrator (not bit-faithful)")
    ap.add_argument("--ref-potentials", type=int, default=0, help="size of the CPU sample (default 24 x host threads, ~15-20 s)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-weak", action="store_true", help="skip the extra weak-scaling measurement at N>1")
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

    from phaseshifts_b200.distributed import shard_range
    NE, L1 = args.energies, args.lmax + 1
    lib = _lib.load()
    opts = _lib.RelOpts(api.XS_DEFAULT, api.DX_DEFAULT, args.lmax, _lib.UNWRAP | (_lib.FAST if args.fast else 0))
    stream = torch.cuda.current_stream(dev)
    flush = torch.empty(256 * 1024 * 1024 // 8, dtype=torch.float64, device=dev)

    def ptr(t):
        return C.c_void_p(t.data_ptr())

    class Shard:
        """This rank's potentials [first, first+P) of the sweep, as pinned host, pageable host and device copies."""

        def __init__(self, first, P):
            self.P, self.first = P, first
            self.wl = make_workload(P, NE, args.lmax, first=first)
            g = self.wl["grid"]
            self.units = P * NE * L1
            self.pot_h = torch.from_numpy(self.wl["pot"]).pin_memory()
            self.jri_h = torch.from_numpy(self.wl["jri"]).pin_memory()
            self.e0_h = torch.from_numpy(g["e_l0"]).pin_memory()
            self.e1_h = torch.from_numpy(g["e_l"]).pin_memory()
            self.out_h = torch.empty((P, NE, L1), dtype=torch.float64).pin_memory()
            self.pot_d, self.jri_d = self.pot_h.to(dev), self.jri_h.to(dev)
            self.e0_d, self.e1_d = self.e0_h.to(dev), self.e1_h.to(dev)
            self.delta_d = torch.empty((P, NE, L1), dtype=torch.float64, device=dev)
            self.raw_d = torch.empty((P, L1, NE, 2), dtype=torch.float64, device=dev)

        def step(self):  # the public device-tensor call a user makes (kernel A + kernel B on torch's stream)
            api.phsh_rel_batch(self.pot_d, self.jri_d, self.e1_d, args.lmax, energies_l0=self.e0_d, unwrap=True,
                               fast=args.fast, out=self.delta_d)

        def integrate_only(self, cnt=None):
            _lib.check(lib.phsh_b200_rel_integrate_device(
                ptr(self.pot_d), self.pot_d.shape[1], ptr(self.jri_d), self.P, ptr(self.e0_d), ptr(self.e1_d), 0, NE,
                C.byref(opts), ptr(self.raw_d), ptr(cnt) if cnt is not None else None, C.c_void_p(stream.cuda_stream)))

        def device_ms_per_step(self, steps, warmup, sampler=None):
            """W untimed + K timed steps, barrier + synchronize on both sides, CUDA events, max over ranks."""
            for _ in range(warmup):
                self.step()
                flush.fill_(1.0)
            barrier()
            if sampler is not None:
                sampler.start()
            ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            barrier()
            ev0.record(stream)
            for _ in range(steps):
                self.step()
                flush.fill_(1.0)
            ev1.record(stream)
            barrier()
            return max_over_ranks(ev0.elapsed_time(ev1)) / steps

        def e2e_seconds_per_step(self, steps, pinned=True):
            """Host buffers through api.phsh_rel_batch -> phsh_b200_rel_batch_host: H2D + kernels + D2H timed."""
            if pinned:
                a = [self.pot_h.numpy(), self.jri_h.numpy(), self.e0_h.numpy(), self.e1_h.numpy(), self.out_h.numpy()]
            else:  # what a numpy user passes: ordinary pageable arrays
                a = [np.array(self.wl["pot"]), np.array(self.wl["jri"]), np.array(self.wl["grid"]["e_l0"]),
                     np.array(self.wl["grid"]["e_l"]), np.zeros((self.P, NE, L1))]
            pot_np, jri_np, e0_np, e1_np, out_np = a

            def one():
                api.phsh_rel_batch(pot_np, jri_np, e1_np, args.lmax, energies_l0=e0_np, unwrap=True, fast=args.fast,
                                   device=local, out=out_np)

            for _ in range(2):
                one()
            barrier()
            t0 = time.perf_counter()
            for _ in range(steps):
                one()
            torch.cuda.synchronize(dev)
            sec = max_over_ranks(time.perf_counter() - t0) / steps
            h2d = pot_np.nbytes + jri_np.nbytes + e0_np.nbytes + e1_np.nbytes
            return sec, int(h2d), int(out_np.nbytes + 32), out_np

    if args.weak:
        a, b = rank * args.potentials, (rank + 1) * args.potentials
        total_pots = args.potentials * world
    else:
        a, b = shard_range(args.potentials, rank, world)
        total_pots = args.potentials
    sh = Shard(a, b - a)
    P = sh.P
    total_units = total_pots * NE * L1

    # ---- device-resident throughput ("value") ---------------------------------------------------------
    sampler = ClockSampler(local) if rank == 0 else None
    ms_per_step = sh.device_ms_per_step(args.steps, args.warmup, sampler)
    clocks = sampler.stop() if rank == 0 else None
    value = total_units / (ms_per_step * 1e-3)

    # ---- roofline of the dominant kernel (rel_integrate_kernel), on this rank's shard ---------------
    cnt = torch.zeros(4, dtype=torch.int64, device=dev)
    sh.integrate_only(cnt)
    torch.cuda.synchronize(dev)
    c = cnt.cpu().numpy()
    flops_per_launch = 330.0 * c[0] + 23.0 * c[1] + 28.0 * c[2]
    k0, k1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    kms = []
    for _ in range(max(3, args.steps)):
        flush.fill_(1.0)
        k0.record(stream)
        sh.integrate_only()
        k1.record(stream)
        torch.cuda.synchronize(dev)
        kms.append(k0.elapsed_time(k1))
    kernel_ms = float(np.mean(kms))
    peak_tf, peak_mhz = api.fp64_peak(local)
    achieved_tf = flops_per_launch / (kernel_ms * 1e-3) / 1e12
    # DRAM traffic per launch and FP64-pipe utilisation come from the committed ncu --set full capture of this very
    # launch (profiles/); they are only quoted when the configuration is the captured one
    traffic, ncu_pipe = None, None
    for name in ("r2_rel_integrate_ncu_c4.json", "r1_rel_integrate_ncu_c4.json"):  # the latest committed capture
        try:
            with open(os.path.join(ROOT, "profiles", name)) as f:
                cap = json.load(f)
            if (P, NE, args.lmax, args.fast) == (4096, 1000, 15, False):
                traffic = cap["traffic_bytes_per_launch"]
                ncu_pipe = cap["fp64_pipe_pct_of_peak_sustained_active"]
            break
        except (OSError, KeyError, ValueError):
            continue

    # ---- end to end through the host-buffer C ABI (H2D + kernels + D2H inside the timed region) ------
    e2e_s, h2d, d2h, out_np = sh.e2e_seconds_per_step(args.steps, pinned=True)
    e2e_value = total_units / e2e_s
    same = bool(np.array_equal(sh.delta_d.cpu().numpy(), out_np))  # device result == host-path result (same kernels)
    page_s, _, _, out_pg = sh.e2e_seconds_per_step(max(2, min(args.steps, 5)), pinned=False)
    same_pg = bool(np.array_equal(out_pg, out_np))

    # ---- CPU baseline on a bounded sample (rank 0, N=1 only) ----------------------------------------
    cpu = None
    parity = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        threads = host_threads()
        n_pot = max(1, min(P, args.ref_potentials or 24 * threads))
        rate, dt, r = cpu_reference_rate(sh.wl, n_pot, threads)
        rate1, dt1, _ = cpu_reference_rate(sh.wl, min(4, n_pot), 1)
        cpu = {"value": rate, "unit": UNIT, "cores": threads, "kind": "port",
               "sample": "%d of %d potentials x %d energies x %d l, %.1f s" % (n_pot, P, NE, L1, dt),
               "single_thread_value": rate1,
               "single_thread_sample": "%d potentials, %.1f s" % (min(4, n_pot), dt1)}
        from oracle import oracle as orc
        ref = np.stack([orc.conphas(x) for x in r["jf"]])
        parity = float(np.abs(out_np[:n_pot] - ref).max())

    # ---- the other scaling mode, for the record (N>1 only): 4096 potentials on EVERY GPU ----------------
    other = None
    if world > 1 and not args.weak and not args.no_weak:
        wsh = Shard(rank * args.potentials, args.potentials)
        wsteps = max(2, min(args.steps, 5))
        wms = wsh.device_ms_per_step(wsteps, 2)
        ws, _, _, _ = wsh.e2e_seconds_per_step(wsteps, pinned=True)
        wunits = world * wsh.units
        other = {"scaling": "weak", "potentials_per_gpu": args.potentials, "steps": wsteps, "ms_per_step": wms,
                 "value": wunits / (wms * 1e-3), "e2e_value": wunits / ws, "unit": UNIT}

    # the same kernel seen against the HBM roofline (measured copy bandwidth of this pool's B200s): far from it
    hbm_peak, hbm_src = 6650.0, "fallback of B200_PROFILING.md"
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            hbm_peak, hbm_src = float(json.load(f)["hbm_gbs"]), "MEASURED_PEAKS.json"
    except (OSError, KeyError, ValueError):
        pass
    alg_bytes = sh.pot_d.numel() * 8.0 + sh.raw_d.numel() * 8.0
    hbm_view = {"achieved": alg_bytes / (kernel_ms * 1e-3) / 1e9, "peak": hbm_peak, "unit": "GB/s",
                "frac": alg_bytes / (kernel_ms * 1e-3) / 1e9 / hbm_peak, "peak_source": hbm_src,
                "algorithmic_bytes_per_launch": alg_bytes}

    if world > 1:
        dist.destroy_process_group()
    if rank != 0:
        return
    nominal_tf = 148 * 64 * 2 * (clocks.get("sm_mhz") or 1965.0) * 1e6 / 1e12 if clocks else None
    out = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms_per_step, "units_per_step": total_units, "higher_is_better": True,
        "scaling": "weak" if args.weak else "strong", "vs_baseline": None,
        "dtype": "f64", "data": "synthetic", "config": workload_config(args, world),
        "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                "api": "phaseshifts_b200.api.phsh_rel_batch(host arrays) -> phsh_b200_rel_batch_host",
                "host_buffers": "pinned", "matches_device_path": same,
                "pageable": {"value": total_units / page_s, "unit": UNIT, "ratio_to_pinned": e2e_s / page_s,
                             "matches_pinned": same_pg,
                             "note": "ordinary numpy arrays: results leave through the library's pinned staging ring"}},
        "gpu_launches": 2 * args.steps,
        "clocks": clocks,
        "roofline": {"bound": "fp64", "achieved": achieved_tf, "peak": peak_tf, "unit": "TFLOP/s",
                     "frac": achieved_tf / peak_tf if peak_tf else None,
                     "frac_of_nominal": achieved_tf / nominal_tf if nominal_tf else None, "nominal_peak": nominal_tf,
                     "traffic": traffic,
                     "traffic_unit": "bytes per launch (dram__bytes_read.sum + dram__bytes_write.sum; algorithmic: "
                                     "%.3g in + %.3g per-kappa scratch out)" % (sh.pot_d.numel() * 8.0, P * L1 * NE * 16.0),
                     "ncu_fp64_pipe_pct_of_peak_sustained_active": ncu_pipe,
                     "hbm_view": hbm_view,
                     "kernel": "rel_integrate_kernel", "kernel_ms": kernel_ms, "potentials_in_launch": P,
                     "flops_per_launch": flops_per_launch,
                     "corrector_evals_per_step": float(c[2]) / max(float(c[1]), 1.0),
                     "peak_source": "measured live: 8 independent DFMA chains/thread, all SMs (implies %.0f MHz); "
                                    "MEASURED_PEAKS.json has no FP64 figure; nominal = 148 SM x 64 DFMA/clk x 2 x the "
                                    "median SM clock of the timed region" % peak_mhz,
                     "note": "flops = as-written FP64 add/mul count of the reference (330/call + 23/Milne step + "
                             "28/corrector evaluation); the strict kernel issues one FP64 instruction per add/mul, so "
                             "its ceiling is 0.5 of the DFMA peak"},
        "cpu_baseline": cpu,
        "parity_max_abs_err_rad": parity,
    }
    if other is not None:
        out["weak"] = other
    # the live FP64 peak next to the clocks it was measured at (copied to profiles/ per round)
    try:
        os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
        with open(os.path.join(ROOT, "gpurun_out", "fp64_peak_n%d.json" % world), "w") as f:
            json.dump({"fp64_dfma_peak_tflops": peak_tf, "implied_sm_mhz": peak_mhz, "nominal_peak_tflops": nominal_tf,
                       "clocks": clocks, "n_gpus": world}, f)
    except OSError:
        pass
    emit(out, real_stdout)


if __name__ == "__main__":
    main()
