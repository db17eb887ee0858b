"""The muffin-tin potential step (CAVPOT) on the GPU, through the C ABI, against the CPU restatement in promoted-double
arithmetic (oracle/cavpot_oracle.hpp) and the reference's golden pair cluster_Re_bulk.i + atomic_Re.i -> mufftin_Re_bulk.d.

Tolerance: the kernel performs the oracle's operations in the oracle's order without FMA contraction; the only
differences are CUDA's vs glibc's log / pow (<= 2 ulp), so potentials must agree to 1e-11 relative (observed 2e-16) and every
integer quantity (mesh indices, neighbour shells, interstitial sample counts) exactly."""
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))
RE = os.path.join(HERE, "golden", "Re0001")


@pytest.fixture(scope="module")
def re_case():
    from oracle import cavpot_io as cio
    cl = cio.read_cluster(os.path.join(RE, "cluster_Re_bulk.i"), real32=False)
    blocks = cio.read_atomic(os.path.join(RE, "atomic_Re.i"), cl["nr"], real32=False)
    rho = np.zeros((cl["nr"], 250))
    nx = np.zeros(cl["nr"], dtype=np.int32)
    for i, b in enumerate(blocks):
        rho[i], nx[i] = cio.rela_density(b, real32=False)
    return cl, blocks, rho, nx


def _struct(c, **kw):
    d = dict(spa=c["spa"], rc=c["rc"], nrr=c["nrr"], z=c["z"], zc=c["zc"], rmt=c["rmt"], rk=c["rk"],
             dens=kw.pop("dens", np.arange(c["nr"])), alpha=c["alpha"], nh=c["nh"], nform=c["nform"])
    d.update(kw)
    return d


def _check(out, t0, ref, nr):
    """one structure of a GPU batch (first type row t0) against the oracle's result"""
    n = ref["npts"]
    assert list(out["npts"][t0:t0 + nr]) == list(n)
    for i in range(nr):
        a, b = out["pot"][t0 + i, :n[i]], ref["v"][i, :n[i]]
        assert np.isfinite(a).all()
        assert (np.abs(a - b) <= 1e-11 * np.maximum(1.0, np.abs(b))).all()


def test_rela_density_interpolation_on_the_gpu(re_case):
    from oracle import cavpot_io as cio
    from phaseshifts_b200 import api
    cl, blocks, rho, nx = re_case
    for i, b in enumerate(blocks):
        g_rho, g_nx = api.cavpot_rela_density(b["rmin"], b["rmax"], b["rs"], b["iprint"])
        assert g_nx == nx[i] == 250
        assert np.array_equal(g_rho, rho[i])      # CHGRID is add/mul/div only: bit-identical
    # a density whose mesh ends inside Loucks' mesh: NX is reset by CHGRID; the listing branch truncates at 1e-9
    b = dict(blocks[0])
    b["rmax"] = 20.0
    for iprint in (0, 1):
        b["iprint"] = iprint
        o_rho, o_nx = cio.rela_density(b, real32=False)
        g_rho, g_nx = api.cavpot_rela_density(b["rmin"], b["rmax"], b["rs"], iprint)
        assert g_nx == o_nx and np.array_equal(g_rho[:o_nx], o_rho[:o_nx])
    assert o_nx < 250


def test_table_driven_division_and_mesh_index_are_exact():
    """cp_div (quotient from a tabulated correctly rounded reciprocal, two residual corrections) == IEEE division for all
    748 x 2 mesh-difference divisors; cp_index_tab == its threshold table, whose entries were bisected on the host with
    the very formula the oracle evaluates (INT(20*(ALOG(x)+8.8)+add))."""
    from phaseshifts_b200 import api
    assert api.selftest_cavpot(1 << 24, seed=3) == 0
    assert api.selftest_cavpot(1 << 20, seed=11) == 0


@pytest.mark.parametrize("nform", [2, 0, 1])
def test_re_bulk_matches_oracle(re_case, nform):
    from oracle import cavpot_io as cio
    from phaseshifts_b200 import api
    cl, _, rho, nx = re_case
    c = dict(cl)
    c["nform"] = nform
    ref = cio.cavpot(c, rho, nx, real32=False)
    out = api.cavpot_batch([_struct(c)], rho, nx, diagnostics=True)
    _check(out, 0, ref, 2)
    assert out["vhar"][0] == ref["vhar"] and out["vex"][0] == ref["vex"]
    assert list(out["stats"][0][:3]) == [ref["nint"], ref["npoint"], ref["jrws"]] == [283, 1000, 199]
    assert list(out["ncon"]) == list(ref["ncon"]) == [30, 30]
    assert np.array_equal(out["shell_ad"], ref["ad"]) and np.array_equal(out["shell_na"], ref["na"])
    assert np.array_equal(out["shell_ia"], ref["ia"])
    j = ref["jrmt"][0]
    assert np.abs(out["vh"][:, :j] - ref["vh"][:, :j]).max() <= 1e-11 * np.abs(ref["vh"][:, :j]).max()
    assert np.abs(out["vs"][:, :j] - ref["vs"][:, :j]).max() <= 1e-11 * np.abs(ref["vs"][:, :j]).max()


def test_slab_mtzm_and_madelung_branches(re_case):
    from oracle import cavpot_io as cio
    from phaseshifts_b200 import api
    cl, _, rho, nx = re_case
    slab = cio.read_cluster(os.path.join(RE, "cluster_Re_slab.i"), real32=False)
    ref = cio.cavpot(slab, rho, nx, slab=1, esht=-1.6425, real32=False)
    out = api.cavpot_batch([_struct(slab, slab=1, esht=-1.6425)], rho, nx, diagnostics=True)
    _check(out, 0, ref, 2)
    assert out["stats"][0][0] == ref["nint"] == 933
    # MTZM: NH = 0 with one atom type (spherical average between R_mt and R_ws)
    c1 = dict(cl)
    c1.update(nh=0, nr=1, nrr=np.array([2], dtype=np.int32), z=cl["z"][:1], zc=cl["zc"][:1], rmt=cl["rmt"][:1])
    ref = cio.cavpot(c1, rho[:1], nx[:1], real32=False)
    out = api.cavpot_batch([_struct(c1, dens=[0])], rho[:1], nx[:1])
    _check(out, 0, ref, 1)
    assert abs(out["mtz"][0] - ref["mtz"]) <= 1e-12
    # MAD: a non-zero valence switches the Ewald sums on (exp/sin/cos on the device: looser)
    c2 = dict(cl)
    c2["zc"] = np.array([1.0, -1.0])
    ref = cio.cavpot(c2, rho, nx, real32=False)
    out = api.cavpot_batch([_struct(c2)], rho, nx, diagnostics=True)
    j = ref["jrmt"][0]
    assert np.abs(ref["vmad"][:, :j]).max() > 0.05
    assert np.abs(out["vmad"][:, :j] - ref["vmad"][:, :j]).max() <= 1e-10
    n = ref["npts"][0]
    assert np.abs(out["pot"][:, :n] - ref["v"][:, :n]).max() <= 1e-9
    # NH = 0 with two atom types: no muffin-tin zero is computed at all (VHAR = VEX = 0)
    c3 = dict(cl)
    c3["nh"] = 0
    ref = cio.cavpot(c3, rho, nx, real32=False)
    out = api.cavpot_batch([_struct(c3)], rho, nx)
    assert out["mtz"][0] == 0.0 == ref["mtz"]
    _check(out, 0, ref, 2)


def test_ragged_batch_of_perturbed_structures(re_case):
    """64 structures of 1-3 atom types / 1-4 atoms with jittered radii, positions, exchange parameters, sampling grids
    (NH=1 lands inside a sphere first and is doubled) and three different charge densities, in one launch."""
    from oracle import cavpot_io as cio
    from phaseshifts_b200 import api
    cl, _, rho0, nx0 = re_case
    rng = np.random.default_rng(77)
    rho = np.stack([rho0[0], 0.6 * rho0[0], 0.3 * rho0[0]])
    zel = np.array([75.0, 45.0, 22.0])
    nx = np.array([250, 240, 231], dtype=np.int32)
    structs, cls = [], []
    for s in range(64):
        nr = int(rng.integers(1, 4))
        nrr = rng.integers(1, 3, size=nr).astype(np.int32) if nr < 3 else np.ones(nr, dtype=np.int32)
        n = int(nrr.sum())
        base = np.array([[0.0, 0.0, 0.0], [-0.5, 0.2887, -0.8072], [0.5, 0.2887, 0.8], [0.0, 0.5774, 0.1]])[:n]
        dens = rng.integers(0, 3, size=nr)
        c = dict(spa=cl["spa"] * (1 + 0.05 * rng.uniform(-1, 1)), rc=cl["rc"] * np.array([[1, 1, 1], [1, 1, 1], [1, 1, 1.5 if n > 2 else 1]]),
                 nr=nr, nrr=nrr, z=zel[dens], zc=np.zeros(nr), rmt=2.2 + 0.3 * rng.uniform(-1, 1, size=nr),
                 rk=base + 0.02 * rng.normal(size=base.shape), nform=int(rng.integers(0, 3)), alpha=0.7 + 0.05 * rng.uniform(),
                 nh=int(rng.choice([1, 2, 5, 8])))
        cls.append((c, dens))
        structs.append(_struct(c, dens=dens))
    out = api.cavpot_batch(structs, rho, nx, diagnostics=True)
    assert out["type_off"][-1] == sum(c["nr"] for c, _ in cls)
    for s, (c, dens) in enumerate(cls):
        ref = cio.cavpot(c, rho[dens], nx[dens], real32=False)
        t0 = out["type_off"][s]
        _check(out, t0, ref, c["nr"])
        assert list(out["stats"][s][:2]) == [ref["nint"], ref["npoint"]]
        assert abs(out["mtz"][s] - ref["mtz"]) <= 1e-12 * max(1.0, abs(ref["mtz"]))
        assert np.array_equal(out["shell_na"][t0:t0 + c["nr"]], ref["na"])


def test_neighbour_shells_of_random_and_nearly_degenerate_cells(re_case):
    """NBR in isolation: 240 random cells (skewed axes, 1-4 atoms of 1-3 types) whose atoms sit on symmetric positions plus
    offsets of 1e-5 .. 3e-4 lattice units, i.e. shell distances that differ by about the 1e-4 tolerance of the reference's
    ordered scan.  The GPU's warp-parallel scan (with its unordered pre-pass that discards far candidates) must end
    with exactly the oracle's lists: same distances bit for bit, same counts, same types."""
    from oracle import cavpot_io as cio
    from phaseshifts_b200 import api
    cl, _, rho0, nx0 = re_case
    rng = np.random.default_rng(2024)
    sym = np.array([[0, 0, 0], [0.5, 0.5, 0.5], [0.5, 0.5, 0.0], [0.5, 0.0, 0.5], [0.0, 0.5, 0.5], [1 / 3, 2 / 3, 0.5],
                    [0.25, 0.25, 0.25], [0.75, 0.25, 0.5]])
    cells = [np.eye(3), np.array([[0.5, 0.866, 0.0], [0.5, -0.866, 0.0], [0.0, 0.0, 1.6]]),
             np.array([[0.0, 0.5, 0.5], [0.5, 0.0, 0.5], [0.5, 0.5, 0.0]]), np.array([[1.0, 0.0, 0.0], [0.3, 0.9, 0.0], [0.1, 0.2, 1.3]])]
    structs, cls = [], []
    for s in range(240):
        nr = int(rng.integers(1, 4))
        nrr = rng.integers(1, 3, size=nr).astype(np.int32)
        n = int(nrr.sum())
        pos = sym[rng.choice(len(sym), size=n, replace=False)].copy()
        pos += rng.choice([0.0, 1e-5, 5e-5, 1e-4, 3e-4], size=pos.shape) * rng.choice([-1.0, 1.0], size=pos.shape)
        rc = cells[s % 4] * (1.0 + rng.choice([0.0, 1e-5, 1e-4]) * rng.normal(size=(3, 3)) * (s % 3 == 0))
        c = dict(spa=float(rng.uniform(4.0, 9.0)), rc=rc, nr=nr, nrr=nrr, z=np.full(nr, 75.0), zc=np.zeros(nr),
                 rmt=np.full(nr, 1.2), rk=pos, nform=2, alpha=0.7, nh=2)
        cls.append(c)
        structs.append(_struct(c, dens=np.zeros(nr, dtype=np.int32)))
    out = api.cavpot_batch(structs, rho0[:1], nx0[:1], diagnostics=True)
    for s, c in enumerate(cls):
        ref = cio.cavpot(c, np.repeat(rho0[:1], c["nr"], axis=0), np.repeat(nx0[:1], c["nr"]), real32=False)
        t0 = out["type_off"][s]
        sl = slice(t0, t0 + c["nr"])
        assert np.array_equal(out["ncon"][sl], ref["ncon"]), s
        assert np.array_equal(out["shell_na"][sl], ref["na"]) and np.array_equal(out["shell_ia"][sl], ref["ia"]), s
        assert np.array_equal(out["shell_ad"][sl], ref["ad"]), s


def test_herman_skillman_and_clementi_densities_and_mixed_atomic_file(tmp_path):
    """The two other charge-density forms of atomic.i: HSIN (table -> CHGRID on the GPU, bit-identical to the oracle) and
    CLEMIN (Slater orbitals evaluated on the GPU: device exp vs libm, 1e-13 relative); then a cluster whose three atom types
    take their densities from a RELA, a CLEM and a HERM block of one atomic file, through the file interface."""
    from cavpot_inputs import hydrogen_clem, hydrogen_herm, neon_like_clem
    from oracle import cavpot_io as cio
    from phaseshifts_b200 import api
    rela = open(os.path.join(RE, "atomic_Re.i")).read().split("\n")[:1004]
    (tmp_path / "a.i").write_text("\n".join(rela + neon_like_clem() + hydrogen_herm()) + "\n")
    blocks = cio.read_atomic(str(tmp_path / "a.i"), 3, real32=False)
    assert [b["kind"] for b in blocks] == ["RELA", "CLEM", "HERM"]
    o_c, n_c = cio.density(blocks[1], real32=False)
    g_c, gn_c = api.cavpot_clemin_density(blocks[1]["bases"], blocks[1]["states"])
    assert gn_c == n_c and np.abs(g_c - o_c).max() <= 1e-12 * o_c.max() and (g_c[n_c:] == 0).all()
    b = blocks[2]
    o_h, n_h = cio.density(b, real32=False)
    g_h, gn_h = api.cavpot_hsin_density(b["z"], b["nm"], b["lc"], b["n"], b["frac"], b["wf"], b["iprint"])
    assert gn_h == n_h and np.array_equal(g_h, o_h)
    (tmp_path / "c.i").write_text("\n".join([
        "mixed densities", "  7.0000", "  1.0000  0.0000  0.0000", "  0.0000  1.0000  0.0000", "  0.0000  0.0000  1.0000", "   3",
        "HEAVY", "   1 75.0000  0.0000  2.2000", "  0.0000  0.0000  0.0000",
        "TEN ELECTRONS", "   1 10.0000  0.0000  1.4000", "  0.5000  0.5000  0.0000",
        "ONE ELECTRON", "   1  1.0000  0.0000  0.9000", "  0.5000  0.0000  0.5000",
        "   0", "  0.7000", "   6"]) + "\n")
    g = [str(tmp_path / n) for n in ("g.d", "g.i", "g.txt")]
    o = [str(tmp_path / n) for n in ("o.d", "o.i", "o.txt")]
    mtz = api.cavpot_file("", 0, str(tmp_path / "a.i"), str(tmp_path / "c.i"), *g)
    omtz, res, _ = cio.cavpot_file("", 0, str(tmp_path / "a.i"), str(tmp_path / "c.i"), *o, real32=False)
    assert abs(mtz - omtz) < 1e-6
    la, lb = open(g[0]).read().split("\n"), open(o[0]).read().split("\n")
    assert len(la) == len(lb) and la[0] == "   3"
    bad = [i for i, (x, y) in enumerate(zip(la, lb)) if x != y]
    for i in bad:    # 5 printed digits: at most a last-digit flip
        xv, yv = np.array(la[i].split(), dtype=float), np.array(lb[i].split(), dtype=float)
        assert np.abs(xv - yv).max() <= 1.0001e-4 * np.abs(yv).max()
    assert len(bad) <= 3
    # 'POTE' is refused with the reason
    (tmp_path / "p.i").write_text("POTE\n   3\n")
    with pytest.raises(api._lib.PhshB200Error) as e:
        api.cavpot_file("", 0, str(tmp_path / "p.i"), str(tmp_path / "c.i"), *g)
    assert "POTE" in str(e.value)


def test_asbuilt_neighbour_search(re_case, tmp_path, monkeypatch):
    """PHSH_B200_ASBUILT: the neighbour search of libphsh.f as built (`exit` out of the loop over atom types).  Same lists
    and potentials as the oracle's as-built restatement; the second atom type loses shells, and the golden mufftin.d (made
    by the original program) is no longer reproduced."""
    from oracle import cavpot_io as cio
    from phaseshifts_b200 import api, libphsh
    cl, _, rho, nx = re_case
    ref = cio.cavpot(cl, rho, nx, real32=False, asbuilt=True)
    good = cio.cavpot(cl, rho, nx, real32=False)
    out = api.cavpot_batch([_struct(cl)], rho, nx, diagnostics=True, asbuilt=True)
    assert np.array_equal(out["shell_ad"], ref["ad"]) and np.array_equal(out["shell_na"], ref["na"])
    assert np.array_equal(out["shell_ia"], ref["ia"]) and list(out["ncon"]) == list(ref["ncon"])
    assert ref["na"][1].sum() < good["na"][1].sum()
    _check(out, 0, ref, 2)
    c6 = cio.read_cluster(os.path.join(HERE, "golden", "examples", "cluster_C6Ni6.i"), real32=False)
    blocks = cio.read_atomic(os.path.join(HERE, "golden", "examples", "atomic_C6Ni6.i"), 8, real32=False)
    rho8 = np.zeros((8, 250))
    nx8 = np.zeros(8, dtype=np.int32)
    for i, b in enumerate(blocks):
        rho8[i], nx8[i] = cio.density(b, real32=False)
    ref8 = cio.cavpot(c6, rho8, nx8, real32=False, asbuilt=True)
    out8 = api.cavpot_batch([_struct(c6)] * 150, rho8, nx8, diagnostics=True, asbuilt=True)   # 128-thread CTAs
    assert np.array_equal(out8["shell_na"][-8:], ref8["na"]) and np.array_equal(out8["shell_ad"][:8], ref8["ad"])
    _check(out8, 8 * 149, ref8, 8)
    monkeypatch.setenv("PHSH_B200_ASBUILT", "1")
    g = [str(tmp_path / n) for n in ("m.d", "b.i", "i.txt")]
    libphsh.cavpot("", 0, os.path.join(RE, "atomic_Re.i"), os.path.join(RE, "cluster_Re_bulk.i"), *g)
    gold, mine = _values(os.path.join(RE, "mufftin_Re_bulk.d")), _values(g[0])
    assert np.abs(gold[0][3] - mine[0][3]).max() > 1e-2 or np.abs(gold[1][3] - mine[1][3]).max() > 1e-2


def test_fine_sampling_grid_spans_many_passes(re_case):
    """NH = 24 -> 13824 sample points, ~3900 interstitial: the ordered compaction list is filled and drained dozens of times
    (128-thread CTAs in a 150-structure batch, 512-thread CTA alone); VHAR/VEX must stay bit-identical to the serial sum."""
    from oracle import cavpot_io as cio
    from phaseshifts_b200 import api
    cl, _, rho, nx = re_case
    c = dict(cl)
    c["nh"] = 24
    ref = cio.cavpot(c, rho, nx, real32=False)
    assert ref["npoint"] == 13824 and ref["nint"] > 3000
    one = api.cavpot_batch([_struct(c)], rho, nx, diagnostics=True)
    assert one["vhar"][0] == ref["vhar"] and one["vex"][0] == ref["vex"]
    assert list(one["stats"][0][:2]) == [ref["nint"], ref["npoint"]]
    c2 = dict(cl)
    c2["nh"] = 2
    many = api.cavpot_batch([_struct(c2)] * 100 + [_struct(c)] + [_struct(c2)] * 49, rho, nx)
    assert many["vhar"][100] == ref["vhar"] and many["vex"][100] == ref["vex"]
    _check(many, 200, ref, 2)


def test_every_route_through_the_muffin_tin_zero_gives_the_same_bits(re_case, monkeypatch):
    """MTZ's interstitial points go through one flat list over the batch (chunked when the list would not fit the scratch
    budget) or, for sample grids that are doubled / too fine for the list / of unknown size, through the in-CTA form:
    VHAR, VEX and the potentials must not depend on the route.  NH = 41 is beyond the list's limit (40) and is checked
    against the oracle directly."""
    from oracle import cavpot_io as cio
    from phaseshifts_b200 import api
    cl, _, rho, nx = re_case
    rng = np.random.default_rng(5)
    structs = []
    for s in range(300):
        c = dict(cl)
        c["spa"] = cl["spa"] * (1 + 0.03 * rng.uniform(-1, 1))
        c["rmt"] = cl["rmt"] * (1 + 0.05 * rng.uniform(-1, 1, size=2))
        c["nh"] = int(rng.choice([1, 3, 6, 10]))
        structs.append(_struct(c))
    ref = api.cavpot_batch(structs, rho, nx, diagnostics=True)
    assert (ref["stats"][:, 0] > 0).all()
    monkeypatch.setenv("PHSH_B200_CAVPOT_LIST_MB", "1")     # 1 MB / (1000 entries * 20 B) -> chunks of 52 structures
    chunked = api.cavpot_batch(structs, rho, nx, diagnostics=True)
    monkeypatch.delenv("PHSH_B200_CAVPOT_LIST_MB")
    monkeypatch.setenv("PHSH_B200_CAVPOT_INCTA", "1")
    incta = api.cavpot_batch(structs, rho, nx, diagnostics=True)
    monkeypatch.delenv("PHSH_B200_CAVPOT_INCTA")
    for other in (chunked, incta):
        for k in ("vhar", "vex", "pot", "npts", "stats"):
            assert np.array_equal(ref[k], other[k]), k
    for s in (0, 17, 299):
        c = dict(cl, spa=structs[s]["spa"], rmt=structs[s]["rmt"], nh=structs[s]["nh"])
        o = cio.cavpot(c, rho, nx, real32=False)
        # VHAR is add/mul/div/sqrt only (bit-identical); VEX goes through pow (CUDA's vs glibc's: a few ulp per point)
        assert ref["vhar"][s] == o["vhar"], (s, ref["vhar"][s], o["vhar"])
        assert abs(ref["vex"][s] - o["vex"]) <= 1e-13 * abs(o["vex"]), (s, ref["vex"][s], o["vex"])
        assert list(ref["stats"][s][:2]) == [o["nint"], o["npoint"]]
    c = dict(cl)
    c["nh"] = 41
    o = cio.cavpot(c, rho, nx, real32=False)
    fine = api.cavpot_batch([_struct(c)] * 2, rho, nx, diagnostics=True)
    assert fine["vhar"][1] == o["vhar"] and fine["vex"][1] == o["vex"] and list(fine["stats"][1][:2]) == [o["nint"], o["npoint"]]


def test_many_atom_types_in_both_launch_configurations(re_case):
    """20 atom types / 40 atoms in a large cell (154 KB of shared memory per CTA): alone (one 512-thread CTA) and as part of
    a 150-structure batch (128-thread CTAs) -- same bits either way, and equal to the oracle."""
    from oracle import cavpot_io as cio
    from phaseshifts_b200 import api
    cl, _, rho0, nx0 = re_case
    rng = np.random.default_rng(9)
    nr = 20
    pos = rng.uniform(0, 1, size=(40, 3))
    c = dict(spa=14.0, rc=np.array([[1.0, 0.0, 0.0], [0.1, 1.0, 0.0], [0.0, 0.2, 1.1]]), nr=nr, nrr=np.full(nr, 2, dtype=np.int32),
             z=np.full(nr, 75.0), zc=np.zeros(nr), rmt=rng.uniform(0.9, 1.3, size=nr), rk=pos, nform=2, alpha=0.7, nh=4)
    rho, nx = rho0[:1], nx0[:1]
    ref = cio.cavpot(c, np.repeat(rho, nr, axis=0), np.repeat(nx, nr), real32=False)
    st = _struct(c, dens=np.zeros(nr, dtype=np.int32))
    one = api.cavpot_batch([st], rho, nx, diagnostics=True)
    _check(one, 0, ref, nr)
    assert np.array_equal(one["shell_na"], ref["na"]) and list(one["stats"][0][:2]) == [ref["nint"], ref["npoint"]]
    small = _struct(cl)
    small["dens"] = np.zeros(2, dtype=np.int32)
    many = api.cavpot_batch([small] * 75 + [st] + [small] * 74, rho, nx)
    t0 = many["type_off"][75]
    assert np.array_equal(many["pot"][t0:t0 + nr], one["pot"]) and many["mtz"][75] == one["mtz"][0]
    assert np.array_equal(many["pot"][0:2], many["pot"][-2:])


def test_structures_to_phase_shifts_on_the_device(re_case):
    """cavpot_batch(on_device) hands `pot`/`npts` straight to phsh_rel_batch (same grid, r*V, even stride): geometry
    variants -> muffin-tin potentials -> delta_l(E) without a host round trip; against the oracle chain to 1e-8 rad."""
    import torch
    from oracle import cavpot_io as cio, oracle as orc
    from phaseshifts_b200 import api
    cl, _, rho, nx = re_case
    rng = np.random.default_rng(5)
    cls = []
    for s in range(6):
        c = dict(cl)
        c["rmt"] = np.round(cl["rmt"] * (1 + 0.08 * rng.uniform(-1, 1)), 4)
        c["rk"] = cl["rk"] + np.round(0.01 * rng.normal(size=cl["rk"].shape), 4)
        cls.append(c)
    d2, p2 = api.structures_to_phase_shifts([_struct(c) for c in cls], rho, nx, api.rel_energy_grid(20.0, 5.0, 300.0)["e_l"], 12,
                                            energies_l0=api.rel_energy_grid(20.0, 5.0, 300.0)["e_l0"])
    dev = api.cavpot_batch([_struct(c) for c in cls], rho, nx, on_device=True)
    assert torch.equal(p2["pot"], dev["pot"])
    host = api.cavpot_batch([_struct(c) for c in cls], rho, nx)
    assert dev["pot"].is_cuda and dev["pot"].shape == (12, 422)
    assert np.array_equal(dev["pot"].cpu().numpy()[:, :421], host["pot"]) and np.array_equal(dev["npts"].cpu().numpy(), host["npts"])
    assert np.array_equal(dev["mtz"].cpu().numpy(), host["mtz"])
    g = api.rel_energy_grid(20.0, 5.0, 300.0)
    delta = api.phsh_rel_batch(dev["pot"], dev["npts"], g["e_l"], 12, energies_l0=g["e_l0"], unwrap=True)
    torch.cuda.synchronize()
    assert torch.equal(d2, delta)
    delta = delta.cpu().numpy()
    assert delta.shape == (12, 57, 13) and np.isfinite(delta).all()
    for s, c in enumerate(cls):
        ref = cio.cavpot(c, rho, nx, real32=False)
        pot = np.zeros((2, 422))
        pot[:, :421] = ref["v"][:, :421]
        jf = orc.rel_batch(pot, ref["npts"], g["e_l0"], g["e_l"], lsm=12)["jf"]
        want = np.stack([orc.conphas(x) for x in jf])
        assert np.abs(delta[2 * s:2 * s + 2] - want).max() <= 1e-8


def test_real_eight_type_input_nform1(tmp_path):
    """phaseshifts/examples/input_files/cluster_C6Ni6.i: 8 atom types, 20 atoms, NFORM=1 (Williams), 6 different charge
    densities -- batch entry point and file interface against the oracle."""
    from oracle import cavpot_io as cio
    from phaseshifts_b200 import api
    ex = os.path.join(HERE, "golden", "examples")
    cl = cio.read_cluster(os.path.join(ex, "cluster_C6Ni6.i"), real32=False)
    assert cl["nr"] == 8 and len(cl["rk"]) == 20 and cl["nform"] == 1
    blocks = cio.read_atomic(os.path.join(ex, "atomic_C6Ni6.i"), cl["nr"], real32=False)
    rho = np.zeros((8, 250))
    nx = np.zeros(8, dtype=np.int32)
    for i, b in enumerate(blocks):
        rho[i], nx[i] = cio.rela_density(b, real32=False)
        g_rho, g_nx = api.cavpot_rela_density(b["rmin"], b["rmax"], b["rs"], b["iprint"])
        assert g_nx == nx[i] and np.array_equal(g_rho, rho[i])
    ref = cio.cavpot(cl, rho, nx, real32=False)
    out = api.cavpot_batch([_struct(cl)], rho, nx, diagnostics=True)
    _check(out, 0, ref, 8)
    assert list(out["stats"][0][:2]) == [ref["nint"], ref["npoint"]] and abs(out["mtz"][0] - ref["mtz"]) < 1e-12
    assert np.array_equal(out["shell_na"], ref["na"]) and np.array_equal(out["shell_ia"], ref["ia"])
    g = [str(tmp_path / n) for n in ("g.d", "g.i", "g.txt")]
    o = [str(tmp_path / n) for n in ("o.d", "o.i", "o.txt")]
    api.cavpot_file("", 0, os.path.join(ex, "atomic_C6Ni6.i"), os.path.join(ex, "cluster_C6Ni6.i"), *g)
    cio.cavpot_file("", 0, os.path.join(ex, "atomic_C6Ni6.i"), os.path.join(ex, "cluster_C6Ni6.i"), *o, real32=False)
    la, lb = open(g[0]).read().split("\n"), open(o[0]).read().split("\n")
    assert len(la) == len(lb) and la[0] == " &NL2 NRR= 8 &end" and la[1].startswith(" &NL16 Z=")
    assert sum(x != y for x, y in zip(la, lb)) <= 2
    # and the Williams phase-shift program reads what we wrote
    api.phsh_wil_file(g[0], str(tmp_path / "phasout"), str(tmp_path / "dataph"), str(tmp_path / "zph"))
    assert len(open(tmp_path / "phasout").read().splitlines()) > 8


def _values(path):
    lines = open(path).read().split("\n")
    blocks, i = [], 0
    while i + 2 < len(lines) and lines[i].strip():
        n = int(lines[i + 2][14:18])
        rows = -(-n // 5)
        vals = []
        for ln in lines[i + 3:i + 3 + rows]:
            vals += [float(ln[k:k + 14]) for k in range(0, len(ln.rstrip()), 14)]
        blocks.append((lines[i].rstrip(), lines[i + 1], lines[i + 2], np.array(vals)))
        i += 3 + rows
    return blocks


def test_file_interface_against_golden_and_oracle_files(tmp_path):
    from oracle import cavpot_io as cio
    from phaseshifts_b200 import libphsh
    g = [str(tmp_path / n) for n in ("g_mufftin.d", "g_bmtz.i", "g_info.txt")]
    o = [str(tmp_path / n) for n in ("o_mufftin.d", "o_bmtz.i", "o_info.txt")]
    mtz = libphsh.cavpot("", 0, os.path.join(RE, "atomic_Re.i"), os.path.join(RE, "cluster_Re_bulk.i"), *g)
    omtz, _, _ = cio.cavpot_file("", 0, os.path.join(RE, "atomic_Re.i"), os.path.join(RE, "cluster_Re_bulk.i"), *o, real32=False)
    assert isinstance(mtz, float) and abs(mtz - omtz) < 1e-6 and abs(float(open(g[1]).read()) - mtz) < 1e-6
    la, lb = open(g[0]).read().split("\n"), open(o[0]).read().split("\n")
    assert len(la) == len(lb) == 137
    differing = [i for i, (x, y) in enumerate(zip(la, lb)) if x != y]
    assert len(differing) <= 2       # a last-digit flip where a value sits on a rounding boundary, at most
    gold = _values(os.path.join(RE, "mufftin_Re_bulk.d"))
    mine = _values(g[0])
    for (gn, gh, gz, gv), (mn, mh, mz, mv) in zip(gold, mine):
        assert gn == mn and gz == mz and gh[:47] == mh[:47]
        assert np.abs(gv - mv).max() < 6e-4   # the fixture's own REAL*4 noise (see tests/test_cavpot_oracle.py)
    assert "MUFFIN-TIN ZERO" in open(g[2]).read()
    # slab run: potential shifted by (esht - MTZ); the MTZ file stays empty like the Fortran leaves it
    s = [str(tmp_path / n) for n in ("s_mufftin.d", "s_mtz.i", "s_info.txt")]
    r = libphsh.cavpot("%.6f" % (mtz + 0.5), 1, os.path.join(RE, "atomic_Re.i"), os.path.join(RE, "cluster_Re_bulk.i"), *s)
    assert abs(r - (mtz + 0.5)) < 1e-6 and os.path.getsize(s[1]) == 0
    sv = _values(s[0])
    rgrid = 60.0 * np.exp(0.03125 * (np.arange(1, 322) - 421))
    assert np.allclose(sv[0][3] - mine[0][3], -0.5 * rgrid, atol=1.1e-4)   # 7 printed digits of values up to 150


def test_chain_cluster_to_phase_shifts_matches_golden_leedph(tmp_path, monkeypatch):
    """cluster.i + atomic.i -> cavpot -> phsh_rel -> conphas entirely on this package's code: the .phs table agrees with
    the reference's golden leedph_Re.d to two units of its 4th decimal (its potentials were made in REAL*4)."""
    from phaseshifts_b200 import api, libphsh
    from phaseshifts_b200.conphas import Conphas
    muff = str(tmp_path / "mufftin.d")
    libphsh.cavpot("", 0, os.path.join(RE, "atomic_Re.i"), os.path.join(RE, "cluster_Re_bulk.i"), muff, str(tmp_path / "bmtz.i"),
                   str(tmp_path / "info.txt"))
    libphsh.phsh_rel(muff, str(tmp_path / "phasout"), str(tmp_path / "dataph"), str(tmp_path / "inpdat"))
    monkeypatch.chdir(tmp_path)
    files = Conphas.split_phasout(str(tmp_path / "phasout"), output_filenames=[str(tmp_path / "A.ph"), str(tmp_path / "B.ph")])
    con = Conphas(input_files=[files[0]], output_file=str(tmp_path / "A.phs"), formatting="cleed", lmax=10)
    con.calculate()
    lines = open(tmp_path / "A.phs").read().splitlines()
    gl = open(os.path.join(RE, "leedph_Re.d")).read().splitlines()
    got = np.array([[float(x) for x in lines[2 + 2 * i].replace("-", " -").split()] for i in range(57)])
    gold = np.array([[float(x) for x in gl[2 * i + 1].replace("-", " -").split()] for i in range(57)])
    assert np.abs(got - gold).max() <= 2.5e-4


def test_error_behaviour(tmp_path):
    from phaseshifts_b200 import api, PhshB200Error
    ok_c, ok_a = os.path.join(RE, "cluster_Re_bulk.i"), os.path.join(RE, "atomic_Re.i")
    outs = [str(tmp_path / n) for n in ("m.d", "b.i", "i.txt")]
    with pytest.raises(PhshB200Error) as e:
        api.cavpot_file("", 0, ok_a, str(tmp_path / "nope.i"), *outs)
    assert e.value.code == -4
    trunc = tmp_path / "trunc.i"
    trunc.write_text("".join(open(ok_c).readlines()[:8]))
    with pytest.raises(PhshB200Error) as e:
        api.cavpot_file("", 0, ok_a, str(trunc), *outs)
    assert e.value.code == -1
    herm = tmp_path / "herm.i"
    herm.write_text("XXXX\n" + "".join(open(ok_a).readlines()[1:]))
    with pytest.raises(PhshB200Error) as e:
        api.cavpot_file("", 0, str(herm), ok_c, *outs)
    assert e.value.code == -1 and "RELA" in str(e.value)
    with pytest.raises(PhshB200Error):
        api.cavpot_file("not-a-number", 1, ok_a, ok_c, *outs)
    # a muffin-tin radius off Loucks' mesh, or overlapping spheres that leave no interstitial point, are refused / reported
    from oracle import cavpot_io as cio
    cl = cio.read_cluster(ok_c, real32=False)
    rho = np.ones((2, 250))
    nx = np.array([250, 250], dtype=np.int32)
    bad = _struct(cl)
    bad["rmt"] = np.array([2.6, 50.0])
    with pytest.raises(PhshB200Error):
        api.cavpot_batch([bad], rho, nx)
    assert api.cavpot_batch([], rho, nx)["pot"].shape == (0, 421)
    # spheres that fill the whole cell: MTZ halves its sampling step until NH exceeds 256 and gives up -- the batch entry
    # point hands back a NaN muffin-tin zero for that structure only, the file interface reports it
    full = _struct(cl)
    full["rmt"] = np.array([6.0, 6.0])
    ok = _struct(cl)
    r = api.cavpot_batch([ok, full, ok], rho, nx, diagnostics=True)
    assert np.isnan(r["mtz"][1]) and np.isfinite(r["mtz"][[0, 2]]).all() and r["mtz"][0] == r["mtz"][2]
    assert r["stats"][1][0] == 0 and r["stats"][1][3] > 256
    big = tmp_path / "big.i"
    big.write_text(open(ok_c).read().replace("2.6000", "6.0000"))
    with pytest.raises(PhshB200Error) as e:
        api.cavpot_file("", 0, ok_a, str(big), *outs)
    assert "muffin-tin zero" in str(e.value)
