"""BASELINE.json configurations at their full sizes on the GPU, checked through size-independent properties
(determinism, shard invariance, unwrap structure, operation-count identity) plus oracle parity on samples the
CPU finishes in seconds.  Tolerance 1e-8 rad (north_star); the strict rel integration is bit-identical."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu
TOL = 1e-8


def test_c1_re_fixture(oracle):
    from phaseshifts_b200 import api, workloads
    w = workloads.c1_re_fixture()
    g = w["grid"]
    d = api.phsh_rel_batch(w["pot"], w["jri"], g["e_l"], w["lmax"], energies_l0=g["e_l0"], unwrap=True)
    ref = oracle.rel_batch(w["pot"], w["jri"], g["e_l0"], g["e_l"], lsm=w["lmax"])["jf"]
    ref = np.stack([oracle.conphas(x) for x in ref])
    assert d.shape == (2, 57, 13) and np.abs(d - ref).max() <= TOL
    assert np.array_equal(d[0], d[1])  # the two fixture blocks are the same potential


def test_c4_full_sweep_properties(oracle):
    """4096 potentials x 1000 energies x lmax=15 (the bench workload) on device tensors."""
    import torch
    from phaseshifts_b200 import api, workloads
    w = workloads.c4_sweep()
    g = w["grid"]
    pot = torch.from_numpy(w["pot"]).cuda()
    raw, kap, cnt = api.phsh_rel_batch(pot, w["jri"], g["e_l"], 15, energies_l0=g["e_l0"], per_kappa=True, counters=True)
    unw = api.phsh_rel_batch(pot, w["jri"], g["e_l"], 15, energies_l0=g["e_l0"], unwrap=True)
    again = api.phsh_rel_batch(pot, w["jri"], g["e_l"], 15, energies_l0=g["e_l0"], unwrap=True)
    torch.cuda.synchronize()
    assert raw.shape == (4096, 1000, 16) and torch.isfinite(raw).all() and torch.isfinite(unw).all()
    assert torch.equal(unw, again)  # deterministic, bit for bit
    # wrapped phases live in [-pi/2, pi/2]; unwrapping only adds integer multiples of pi ...
    assert float(raw.abs().max()) <= np.pi / 2 + 1e-12
    k = (unw - raw) / np.pi
    assert float((k - k.round()).abs().max()) < 1e-9
    # ... and conphas adds exactly -pi per upward jump >= pi/2 of the RAW neighbours (conphas.py:445-454)
    jumps = torch.zeros_like(raw)
    dif = raw[:, 1:] - raw[:, :-1]
    jumps[:, 1:] = (dif <= -np.pi / 2).double() - (dif >= np.pi / 2).double()
    assert torch.equal(k.round(), jumps.cumsum(1))
    # shard invariance: any slice of potentials computed alone equals the slice of the full run
    sl = slice(1000, 1037)
    part = api.phsh_rel_batch(pot[sl].contiguous(), w["jri"][sl], g["e_l"], 15, energies_l0=g["e_l0"], unwrap=True)
    assert torch.equal(part, unw[sl])
    # oracle parity + operation-count identity on a sample
    idx = [0, 1, 777, 4095]
    ref = oracle.rel_batch(w["pot"][idx], w["jri"][idx], g["e_l0"], g["e_l"], lsm=15, per_kappa=True, nthreads=8)
    assert np.abs(raw[idx].cpu().numpy() - ref["jf"]).max() <= TOL
    assert np.abs(kap[idx].cpu().numpy().transpose(0, 2, 1, 3) - ref["jfk"]).max() <= TOL
    assert cnt["calls"] == 4096 * 1000 * 31
    assert cnt["milne_steps"] == int(1000 * 31 * (w["jri"].astype(np.int64) - 6).sum())
    _, c4 = api.phsh_rel_batch(w["pot"][idx], w["jri"][idx], g["e_l"], 15, energies_l0=g["e_l0"], counters=True)
    assert (c4["calls"], c4["milne_steps"], c4["corr_evals"]) == (ref["calls"], ref["milne_steps"], ref["corr_evals"])


def test_c3_gold_fine_grid_both_kappa(oracle):
    """Z=79, dx=1/64 (J~683), 2000 energies, lmax=18: per-kappa and averaged phases."""
    from phaseshifts_b200 import api, workloads
    w = workloads.c3_gold()
    d, kap = api.phsh_rel_batch(w["pot"], w["jri"], w["e_ryd"], w["lmax"], dx=w["dx"], per_kappa=True)
    assert d.shape == (1, 2000, 19) and np.isfinite(d).all()
    sel = np.arange(0, 2000, 37)
    ref = oracle.rel_batch(w["pot"], w["jri"], w["e_ryd"][sel], lsm=w["lmax"], dx=w["dx"], per_kappa=True, nthreads=8)
    # per-kappa phases do not depend on the energy neighbours, the LVOR average does: compare the former directly
    assert np.abs(kap[0].transpose(1, 0, 2)[sel] - ref["jfk"][0]).max() <= TOL
    full = oracle.rel_batch(w["pot"], w["jri"], w["e_ryd"][:300], lsm=w["lmax"], dx=w["dx"], nthreads=8)["jf"]
    assert np.abs(d[0, :300] - full[0]).max() <= TOL
    # spin-orbit splitting is visible at Z=79: kappa=l and kappa=-l-1 differ for l=1
    assert np.abs(kap[0, 1, :, 0] - kap[0, 1, :, 1]).max() > 1e-2


def test_c2_oxide_full(oracle):
    """4 species x 1999 energies x lmax=15 through phsh_cav, full oracle comparison."""
    from phaseshifts_b200 import api, workloads
    w = workloads.c2_oxide()
    got = api.phsh_cav_batch(w["V"], w["RX"], w["ntab"], w["rmt"], w["e_ryd"], NL=16, unwrap=True)
    ref, rc = oracle.cav_batch(w["V"], w["RX"], w["ntab"], w["rmt"], w["e_ryd"], NL=16, nthreads=8)
    assert rc == 0 and got.shape == (4, 1999, 16)
    ref = np.stack([oracle.conphas(x) for x in ref])
    assert np.abs(got - ref).max() <= TOL


def test_c5_wil_full_scale(oracle):
    """65536 synthetic potentials x 500 energies x lmax=12 through phsh_wil + conphas."""
    import torch
    from phaseshifts_b200 import api, workloads
    w = workloads.c5_wil()
    rs, zs = torch.from_numpy(w["rs"]).cuda(), torch.from_numpy(w["zs"]).cuda()
    out = api.phsh_wil_batch(rs, zs, w["nrows"], w["Z"], w["RT"], w["e"], NL=13, NR=101, unwrap=True)
    torch.cuda.synchronize()
    assert out.shape == (65536, 500, 13) and bool(torch.isfinite(out).all())
    idx = [0, 1, 2, 31337, 65535]
    ref = oracle.wil_batch(w["rs"][idx], w["zs"][idx], w["nrows"][idx], w["Z"][idx], w["RT"][idx], w["e"], NL=13,
                           NR=101, nthreads=5)["del"]
    ref = np.stack([oracle.conphas(x) for x in ref])
    assert np.abs(out[idx].cpu().numpy() - ref).max() <= TOL
    sl = slice(40000, 40016)
    part = api.phsh_wil_batch(rs[sl].contiguous(), zs[sl].contiguous(), w["nrows"][sl], w["Z"][sl], w["RT"][sl], w["e"],
                              NL=13, NR=101, unwrap=True)
    assert torch.equal(part, out[sl])


def test_conphas_kernel_vs_golden_reference_conphas():
    """The stand-alone unwrap kernel on the phasout sections the reference's own Conphas was run on."""
    import os
    from phaseshifts_b200 import api
    con = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "conphas")
    for name, lmax in (("Re_A", 10), ("jumps", 12), ("jumps", 5)):
        lines = open(os.path.join(con, name + ".ph")).read().splitlines()
        n, lmf = int(lines[1].split()[2]), int(lines[1].split()[3])
        data = np.array(" ".join(ln.replace("-", " -") for ln in lines[2:]).split(), dtype=float).reshape(n, lmf + 2)
        got = api.conphas(data[:, 1:], lmax=lmax)
        blocks = open(os.path.join(con, "dataph_%s_lmax%d.d" % (name, lmax))).read().split('"L = ')[1:]
        for l, b in enumerate(blocks):
            arr = np.array([[float(x) for x in ln.split()] for ln in b.strip().split("\n")[1:]])
            assert np.abs(arr[:, 1] - got[:, l]).max() < 1e-7
        assert np.array_equal(got[:, lmax + 1:], data[:, lmax + 2:])  # columns above lmax are passed through


def test_edge_cases():
    from phaseshifts_b200 import api, workloads
    w = workloads.c4_sweep(P=2, ne=1)
    g = w["grid"]
    one = api.phsh_rel_batch(w["pot"], w["jri"], g["e_l"], 0, energies_l0=g["e_l0"], unwrap=True)  # one energy, l=0 only
    assert one.shape == (2, 1, 1) and np.isfinite(one).all()
    empty = api.phsh_rel_batch(np.zeros((0, 326)), np.zeros(0, dtype=np.int32), g["e_l"], 3)
    assert empty.shape == (0, 1, 4)
    assert api.conphas(np.zeros((1, 5))).shape == (1, 5)
    # ragged: every potential its own JRI, including the smallest legal one
    w = workloads.c4_sweep(P=3, ne=33)
    jri = np.array([7, 100, 326], dtype=np.int32)
    d = api.phsh_rel_batch(w["pot"], jri, w["grid"]["e_l"], 2)
    assert d.shape == (3, 33, 3) and np.isfinite(d).all()


def test_c6_cavpot_full_batch_properties():
    """The producer workload at full size (4096 perturbed Re(0001) cells, one launch) through properties that need no
    oracle run: determinism, independence of the position in the batch and of the batch size (128-thread CTAs vs the
    512-thread CTA a lone structure gets), invariance under relabelling the two atom types, the analytic effect of a slab
    shift, and r*V -> -2Z at the origin; plus oracle parity on a sample.  (Moving an atom by a lattice vector is NOT an
    invariance of this algorithm: MTZ tests the spheres of the home cell and its 7 upper neighbours only, phsh1.f:840-858,
    so atoms are expected inside the home cell -- that case is checked against the oracle instead.)"""
    from oracle import cavpot_io as cio
    from phaseshifts_b200 import api, workloads
    w = workloads.c6_cavpot()
    b = w["rela"]
    rho0, nx0 = api.cavpot_rela_density(b["rmin"], b["rmax"], b["rs"], b["iprint"])
    rho, nx = rho0[None, :], np.array([nx0], dtype=np.int32)
    S = len(w["structures"])
    pk = api.cavpot_pack(w["structures"])
    out = api.cavpot_batch(pk, rho, nx)
    again = api.cavpot_batch(pk, rho, nx)
    assert out["pot"].shape == (2 * S, 421) and np.array_equal(out["pot"], again["pot"]) and np.array_equal(out["mtz"], again["mtz"])
    assert np.isfinite(out["pot"]).all() and (out["npts"] >= 300).all() and (out["npts"] <= 340).all()
    assert np.all(out["mtz"] < -1.4) and np.all(out["mtz"] > -1.9)
    assert np.abs(out["pot"][:, 0] / -150.0 - 1.0).max() < 2e-3                      # r*V(r -> 0) = -2Z, Z = 75
    # position in the batch / batch size / CTA width do not matter
    idx = [0, 1, 777, 2048, 4095]
    sub = api.cavpot_batch([w["structures"][i] for i in idx], rho, nx)
    for k, i in enumerate(idx):
        assert np.array_equal(sub["pot"][2 * k:2 * k + 2], out["pot"][2 * i:2 * i + 2]) and sub["mtz"][k] == out["mtz"][i]
    # relabelling the two atom types permutes the rows.  The neighbour scan and the MTZ sums run in another order, so not
    # bit for bit: 1e-9 on values up to 150
    st = dict(w["structures"][5])
    sw = dict(st, rk=np.asarray(st["rk"])[::-1].copy())
    # an atom moved by a lattice vector: the same crystal
    mv = dict(st, rk=np.asarray(st["rk"]) + np.array([[0.0, 0.0, 0.0], np.asarray(st["rc"])[0]]))
    r = api.cavpot_batch([st, sw, mv], rho, nx)
    n = int(r["npts"][0])
    assert np.abs(r["pot"][2, :n] - r["pot"][1, :n]).max() < 1e-9 and np.abs(r["pot"][3, :n] - r["pot"][0, :n]).max() < 1e-9
    assert abs(r["mtz"][1] - r["mtz"][0]) < 1e-11
    cm = dict(spa=mv["spa"], rc=mv["rc"], nr=2, nrr=np.array(mv["nrr"], dtype=np.int32), z=np.array(mv["z"]), zc=np.array(mv["zc"]),
              rmt=np.array(mv["rmt"]), rk=mv["rk"], nform=2, alpha=mv["alpha"], nh=mv["nh"])
    refm = cio.cavpot(cm, np.stack([rho0, rho0]), np.array([nx0, nx0], dtype=np.int32), real32=False)
    assert abs(refm["mtz"] - r["mtz"][0]) > 1e-3                                       # the reference algorithm's own behaviour
    assert np.abs(r["pot"][4:6, :n] - refm["v"][:, :n]).max() <= 1e-11 * 150.0 and r["vhar"][2] == refm["vhar"]
    # slab run with the own muffin-tin zero + d: V shifted by -d, i.e. r*V by -d*r on the Waber grid
    sl = api.cavpot_batch([dict(st, slab=1, esht=float(r["mtz"][0]) + 0.125)], rho, nx)
    rgrid = 60.0 * np.exp(0.03125 * (np.arange(1, n + 1) - 421))
    assert np.abs((sl["pot"][0, :n] - r["pot"][0, :n]) + 0.125 * rgrid).max() < 1e-6
    # oracle parity on a sample
    for i in (3, 1500, 4000):
        s_ = w["structures"][i]
        c = dict(spa=s_["spa"], rc=s_["rc"], nr=2, nrr=np.array(s_["nrr"], dtype=np.int32), z=np.array(s_["z"]),
                 zc=np.array(s_["zc"]), rmt=np.array(s_["rmt"]), rk=s_["rk"], nform=2, alpha=s_["alpha"], nh=s_["nh"])
        ref = cio.cavpot(c, np.stack([rho0, rho0]), np.array([nx0, nx0], dtype=np.int32), real32=False)
        k = int(ref["npts"][0])
        assert list(out["npts"][2 * i:2 * i + 2]) == list(ref["npts"])
        assert np.abs(out["pot"][2 * i:2 * i + 2, :k] - ref["v"][:, :k]).max() <= 1e-11 * 150.0
        assert out["vhar"][i] == ref["vhar"] and out["vex"][i] == ref["vex"]
