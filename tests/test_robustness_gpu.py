"""Robustness of the GPU path: re-entrancy from several host threads (the Fortran is not thread-safe,
README.md:241-249 of the reference), randomised parity over grids / sizes / energy ranges (hypothesis), and the
behaviour on inputs the Fortran would choke on (out-of-range JRI on the device, negative energies)."""
import threading

import numpy as np
import pytest
from hypothesis import HealthCheck, given, settings, strategies as st

pytestmark = pytest.mark.gpu
TOL = 1e-8


def test_host_api_is_reentrant_across_threads(oracle):
    from phaseshifts_b200 import api, workloads
    jobs = []
    for k in range(6):
        w = workloads.c4_sweep(P=5 + k, ne=64 + 7 * k, lmax=3 + k, seed=100 + k)
        jobs.append(w)
    out = [None] * len(jobs)

    def run(i):
        w = jobs[i]
        g = w["grid"]
        out[i] = api.phsh_rel_batch(w["pot"], w["jri"], g["e_l"], w["lmax"], energies_l0=g["e_l0"], unwrap=True)

    th = [threading.Thread(target=run, args=(i,)) for i in range(len(jobs))]
    for t in th:
        t.start()
    for t in th:
        t.join()
    for i, w in enumerate(jobs):
        g = w["grid"]
        ref = oracle.rel_batch(w["pot"], w["jri"], g["e_l0"], g["e_l"], lsm=w["lmax"])["jf"]
        ref = np.stack([oracle.conphas(x) for x in ref])
        assert out[i] is not None and np.abs(out[i] - ref).max() <= TOL


@settings(max_examples=25, deadline=None, suppress_health_check=[HealthCheck.function_scoped_fixture])
@given(seed=st.integers(0, 2 ** 31 - 1), P=st.integers(1, 4), ne=st.integers(1, 70), lmax=st.integers(0, 18),
       dx_den=st.sampled_from([16, 32, 48, 64]), e0=st.floats(0.05, 30.0), de=st.floats(0.01, 3.0),
       unwrap=st.booleans(), asbuilt=st.booleans())
def test_randomised_rel_parity(oracle, seed, P, ne, lmax, dx_den, e0, de, unwrap, asbuilt):
    """Random Z, radius, V0, radial step, energy grid (Rydberg), lmax: every phase within 1e-8 rad of the oracle and
    identical operation counts (i.e. identical branch decisions in every Milne step)."""
    from phaseshifts_b200 import api, workloads
    rng = np.random.default_rng(seed)
    dx = 1.0 / dx_den
    xs = -9.03065133
    J = 2 * ((int(1 + round((np.log(3.4) - xs) / dx)) + 2) // 2)
    r = np.exp(xs + np.arange(J) * dx)
    pot = np.stack([workloads.moliere_rv(int(rng.integers(1, 93)), r) + rng.normal(0, 0.4) * r for _ in range(P)])
    jri = np.array([int(1 + round((np.log(rng.uniform(1.2, 3.3)) - xs) / dx)) for _ in range(P)], dtype=np.int32)
    e = e0 + de * np.arange(ne)
    got, cnt = api.phsh_rel_batch(pot, jri, e, lmax, dx=dx, unwrap=unwrap, asbuilt=asbuilt, counters=True)
    ref = oracle.rel_batch(pot, jri, e, lsm=lmax, dx=dx, asbuilt=asbuilt)
    jf = np.stack([oracle.conphas(x) for x in ref["jf"]]) if unwrap else ref["jf"]
    assert np.abs(got - jf).max() <= TOL
    assert (cnt["calls"], cnt["milne_steps"], cnt["corr_evals"]) == (ref["calls"], ref["milne_steps"], ref["corr_evals"])


@settings(max_examples=10, deadline=None, suppress_health_check=[HealthCheck.function_scoped_fixture])
@given(seed=st.integers(0, 2 ** 31 - 1), P=st.integers(1, 3), ne=st.integers(1, 40), nl=st.integers(1, 15),
       nr=st.sampled_from([21, 57, 101, 151, 201]))
def test_randomised_wil_parity(oracle, seed, P, ne, nl, nr):
    from phaseshifts_b200 import api, workloads
    w = workloads.c5_wil(P=P, ne=ne, nl=nl, nr=nr, seed=seed)
    got = api.phsh_wil_batch(w["rs"], w["zs"], w["nrows"], w["Z"], w["RT"], w["e"], NL=nl, NR=nr)
    ref = oracle.wil_batch(w["rs"], w["zs"], w["nrows"], w["Z"], w["RT"], w["e"], NL=nl, NR=nr)["del"]
    assert np.abs(got - ref).max() <= TOL


@settings(max_examples=10, deadline=None, suppress_health_check=[HealthCheck.function_scoped_fixture])
@given(seed=st.integers(0, 2 ** 31 - 1), ne=st.integers(1, 60), nl=st.integers(1, 20), asbuilt=st.booleans())
def test_randomised_cav_parity(oracle, seed, ne, nl, asbuilt):
    from phaseshifts_b200 import api, workloads
    rng = np.random.default_rng(seed)
    n = 250
    r = workloads.loucks_r(n)
    Z = rng.integers(1, 93, size=3)
    V = np.stack([workloads.moliere_rv(int(z), r) / r for z in Z])
    rmt = rng.uniform(1.0, 3.2, size=3)
    ntab = np.array([min(int(20.0 * (np.log(x) + 8.8) + 2.0) + int(rng.integers(0, 8)), n) for x in rmt], dtype=np.int32)
    e = np.sort(rng.uniform(0.05, 80.0, size=ne))
    got = api.phsh_cav_batch(V, np.tile(r, (3, 1)), ntab, rmt, e, NL=nl, asbuilt=asbuilt)
    ref, rc = oracle.cav_batch(V, np.tile(r, (3, 1)), ntab, rmt, e, NL=nl, asbuilt=asbuilt)
    assert rc == 0 and np.abs(got - ref).max() <= TOL


def test_bad_inputs_on_device_give_nan_not_faults(oracle):
    import torch
    from phaseshifts_b200 import api, workloads, PhshB200Error
    w = workloads.c4_sweep(P=4, ne=40, lmax=3)
    g = w["grid"]
    pot = torch.from_numpy(w["pot"]).cuda()
    jri = torch.tensor([321, 6, 999, 326], dtype=torch.int32, device="cuda")  # 6 and 999 are out of range
    d = api.phsh_rel_batch(pot, jri, g["e_l"], 3, energies_l0=g["e_l0"], unwrap=True).cpu().numpy()
    assert np.isfinite(d[0]).all() and np.isfinite(d[3]).all()
    assert np.isnan(d[1]).all() and np.isnan(d[2]).all()
    ref = oracle.rel_batch(w["pot"][[0, 3]], [321, 326], g["e_l0"], g["e_l"], lsm=3)["jf"]
    ref = np.stack([oracle.conphas(x) for x in ref])
    assert np.abs(d[[0, 3]] - ref).max() <= TOL
    # the device-side error flag: the counters count every lane that rejected its jri (2 potentials x 40 energies,
    # once per l-group CTA), and none on a clean batch
    _, cnt = api.phsh_rel_batch(pot, jri, g["e_l"], 3, energies_l0=g["e_l0"], counters=True)
    assert cnt["rejected_lanes"] >= 2 * 40 and cnt["rejected_lanes"] % (2 * 40) == 0
    _, cnt = api.phsh_rel_batch(pot[[0, 3]], jri[[0, 3]], g["e_l"], 3, energies_l0=g["e_l0"], counters=True)
    assert cnt["rejected_lanes"] == 0
    with pytest.raises(PhshB200Error):  # the host entry point can check, and does
        api.phsh_rel_batch(w["pot"], [321, 6, 999, 326], g["e_l"], 3)
    # E <= 0: SBFIT's sqrt(E) is NaN / 0, every comparison on NaN is false and the reference falls through to its
    # pi/2 default (phsh2.f:1064-1066); the kernel must follow the same path, without a fault
    e = np.array([-0.5, 0.0, 0.5])
    got = api.phsh_rel_batch(w["pot"][:1], w["jri"][:1], e, 2)
    ref = oracle.rel_batch(w["pot"][:1], w["jri"][:1], e, lsm=2)["jf"]
    assert np.array_equal(np.isnan(got), np.isnan(ref))
    ok = ~np.isnan(ref)
    assert np.abs(got[ok] - ref[ok]).max() <= TOL
    assert abs(abs(got[0, 0, 0]) - float(np.float32(3.141592654)) / 2) < 1e-12


def test_single_process_multi_device_sharding(oracle):
    """phsh_rel_batch_multi: one host thread per device writing its slice of one result array (here the same
    device several times, which exercises the concurrency on a one-GPU box as well)."""
    from phaseshifts_b200 import workloads, _lib
    from phaseshifts_b200.distributed import phsh_rel_batch_multi
    w = workloads.c4_sweep(P=11, ne=90, lmax=6)
    g = w["grid"]
    n = _lib.device_count()
    devs = list(range(n)) if n > 1 else [0, 0, 0]
    got = phsh_rel_batch_multi(w["pot"], w["jri"], g["e_l"], 6, energies_l0=g["e_l0"], unwrap=True, devices=devs)
    ref = oracle.rel_batch(w["pot"], w["jri"], g["e_l0"], g["e_l"], lsm=6)["jf"]
    ref = np.stack([oracle.conphas(x) for x in ref])
    assert np.abs(got - ref).max() <= TOL


def test_very_fine_radial_grid_uses_opt_in_shared_memory(oracle):
    """dx = 1/256 -> ~2700 radial points per potential: the staging buffers exceed the 48 KB default and the
    oracle leaves its stack arrays for the heap; results must still agree bit for bit in the integration."""
    from phaseshifts_b200 import api, workloads
    dx, xs = 1.0 / 256.0, -9.03065133
    J = 2 * ((int(1 + round((np.log(3.0) - xs) / dx)) + 2) // 2)
    r = np.exp(xs + np.arange(J) * dx)
    pot = np.stack([workloads.moliere_rv(z, r) for z in (29, 79)])
    jri = np.array([int(1 + round((np.log(rm) - xs) / dx)) for rm in (2.4, 2.9)], dtype=np.int32)
    assert jri.max() > 2500
    e = np.linspace(0.5, 40.0, 70)
    got, cnt = api.phsh_rel_batch(pot, jri, e, 6, dx=dx, counters=True)
    ref = oracle.rel_batch(pot, jri, e, lsm=6, dx=dx)
    assert np.abs(got - ref["jf"]).max() <= TOL
    assert (cnt["milne_steps"], cnt["corr_evals"]) == (ref["milne_steps"], ref["corr_evals"])
