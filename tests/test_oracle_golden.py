"""Pins the CPU oracle on the reference's own golden vectors (tests/Re0001 of the reference, copied to
tests/golden/Re0001) and on outputs of the reference's own Python Conphas (tests/golden/conphas, generated
by tests/golden/make_golden.py).  CPU only."""
import os

import numpy as np
import pytest

from oracle import fortran_io as fio
from oracle import oracle as orc

HERE = os.path.dirname(os.path.abspath(__file__))
RE = os.path.join(HERE, "golden", "Re0001")
CON = os.path.join(HERE, "golden", "conphas")


def _golden_dataph():
    blocks = open(os.path.join(RE, "dataph_Re.d")).read().split('"L=')[1:]
    return [np.array([[float(x) for x in ln.split()] for ln in b.strip().split("\n")[1:] if ln.strip()]) for b in blocks]


def _golden_leedph():
    lines = open(os.path.join(RE, "leedph_Re.d")).read().splitlines()
    e = np.array([float(lines[2 * i]) for i in range(57)])
    ph = np.array([[float(x) for x in lines[2 * i + 1].replace("-", " -").split()] for i in range(57)])
    return e, ph


@pytest.fixture(scope="module")
def re_phasout(tmp_path_factory):
    d = tmp_path_factory.mktemp("re")
    res = fio.phsh_rel_file(os.path.join(RE, "mufftin_Re_slab.d"), str(d / "phasout"), str(d / "dataph"),
                            str(d / "inpdat"), real32=True)
    secs = ["RELATIVISTIC" + s for s in open(d / "phasout").read().split("RELATIVISTIC")[1:]]
    paths = []
    for i, s in enumerate(secs):
        p = d / ("sec%d.ph" % i)
        p.write_text(s)
        paths.append(str(p))
    return res, paths


def test_mufftin_parse(re_blocks):
    assert len(re_blocks) == 2
    b = re_blocks[0]
    assert (b["name"], b["es"], b["de"], b["ue"], b["lsm"], b["nz"], b["jri"]) == \
        ("RHENIUM A       ", 20.0, 5.0, 300.0, 12, 75, 321)
    assert abs(b["rmt"] - 2.6) < 1e-12 and len(b["zp"]) == 321
    assert b["zp"][0] == -149.8956 and b["zp"][-1] == -0.1370789
    assert np.array_equal(re_blocks[0]["zp"], re_blocks[1]["zp"])


def test_energy_grid_matches_fixture_header():
    g = orc.rel_energy_grid(20.0, 5.0, 300.0)
    assert g["n_header"] == 57 and len(g["e_ev"]) == 57
    assert abs(g["e_ev"][0] - 20.0) < 1e-12 and abs(g["e_ev"][-1] - 300.0) < 1e-10
    # phsh2007 semantics: l>=1 loops round-trip E through eV like the l=0 loop => identical sequences
    assert np.array_equal(g["e_l0"], g["e_l"])
    ga = orc.rel_energy_grid(20.0, 5.0, 300.0, asbuilt=True)
    assert np.abs(ga["e_l"] - g["e_l"]).max() < 1e-14


def test_rel_oracle_reproduces_dataph_Re(re_phasout):
    """FP32-as-written oracle -> phasout text -> conphas == the reference's dataph_Re.d to one unit of the last
    printed place (F8.5 => 1e-5; the 9th phase of a row is an F9.4 field => 1e-4)."""
    _, paths = re_phasout
    en, con = fio.conphas_file(paths[0], 10)
    gold = _golden_dataph()
    assert len(gold) == 11
    nbad = 0
    for l in range(11):
        assert np.abs(en / 27.21 - gold[l][:, 0]).max() < 2e-6
        d = np.abs(con[:, l] - gold[l][:, 1])
        assert d.max() < (1.05e-5 if l != 8 else 1.0e-4)
        nbad += int((d > 7e-6).sum())
    # FP32 rounding noise of a different compiler/libm flips the last digit in a few per cent of the entries
    assert nbad <= 0.06 * 57 * 11


def test_rel_oracle_reproduces_leedph_Re(re_phasout):
    _, paths = re_phasout
    e_gold, ph_gold = _golden_leedph()
    for p in paths:  # both blocks are the same potential
        en, con = fio.conphas_file(p, 10)
        assert np.abs(en / 27.21 - e_gold).max() <= 5.001e-5
        assert np.abs(con - ph_gold).max() <= 1.0001e-4


def test_promoted_oracle_is_close_to_as_written(re_blocks):
    b = re_blocks[0]
    g = orc.rel_energy_grid(b["es"], b["de"], b["ue"])
    f32 = orc.rel_batch(b["zp"][None], [b["jri"]], g["e_l0"], g["e_l"], lsm=12, real32=True)["jf"]
    f64 = orc.rel_batch(b["zp"][None], [b["jri"]], g["e_l0"], g["e_l"], lsm=12, real32=False)["jf"]
    assert 1e-8 < np.abs(f32 - f64).max() < 3e-6  # FP32 matching noise: why 1e-8 is only meaningful vs the FP64 oracle


def test_asbuilt_sbfit_is_not_what_the_fixture_encodes(re_blocks):
    """The as-built libphsh.f SBFIT (one recurrence step whatever l) is exact for l=1 and wrong by O(1) rad elsewhere."""
    b = re_blocks[0]
    g = orc.rel_energy_grid(b["es"], b["de"], b["ue"])
    ok = orc.rel_batch(b["zp"][None], [b["jri"]], g["e_l0"], g["e_l"], lsm=6, per_kappa=True)["jfk"][0]
    bad = orc.rel_batch(b["zp"][None], [b["jri"]], g["e_l0"], g["e_l"], lsm=6, per_kappa=True, asbuilt=True)["jfk"][0]
    assert np.abs(ok[:, 1] - bad[:, 1]).max() == 0.0
    for l in (0, 2, 3, 4):
        assert np.abs(ok[:, l] - bad[:, l]).max() > 0.5


def test_dlgkap_counts_and_radius(re_blocks):
    pot = np.abs(re_blocks[0]["zp"])
    v, rmax, steps, evals = orc.dlgkap(pot, 100.0 / 13.6, -1)
    assert steps == 321 - 6 and steps <= evals <= 6 * steps
    assert abs(rmax - np.exp(-9.03065133 + 320 * 0.03125)) < 1e-12
    v2, _, _, ev2 = orc.dlgkap(pot, 100.0 / 13.6, -13)
    assert ev2 > evals  # more corrector passes at high |kappa|
    _, _, _, _, U, W = orc.dlgkap(pot, 100.0 / 13.6, -1, want_uw=True)
    assert U[0] == 1e-25 and np.isfinite(U).all() and np.isfinite(W).all()


@pytest.mark.parametrize("name,lmax", [("Re_A", 10), ("jumps", 12), ("jumps", 5)])
def test_conphas_oracle_vs_reference_conphas(name, lmax):
    """oracle unwrap == the reference's own Conphas.calculate (golden files written by phaseshifts.conphas)."""
    en, con = fio.conphas_file(os.path.join(CON, name + ".ph"), lmax)
    # dataph_<name>.d: "%9.7f\t%9.7f" per (energy/27.21, conpha)
    blocks = open(os.path.join(CON, "dataph_%s_lmax%d.d" % (name, lmax))).read().split('"L = ')[1:]
    assert len(blocks) == lmax + 1
    for l, b in enumerate(blocks):
        arr = np.array([[float(x) for x in ln.split()] for ln in b.strip().split("\n")[1:]])
        assert np.abs(arr[:, 0] - en / 27.21).max() < 1e-7
        assert np.abs(arr[:, 1] - con[:, l]).max() < 1e-7
    # cleed layout: header, then per energy "%7.4f" and (lmax+1) x "%7.4f"
    lines = open(os.path.join(CON, "%s_lmax%d_cleed.out" % (name, lmax))).read().splitlines()
    assert lines[0] == "%d %d neng lmax" % (len(en), lmax)
    for ie in range(len(en)):
        assert lines[1 + 2 * ie] == "%7.4f" % (en[ie] / 27.21)
        assert lines[2 + 2 * ie] == "".join("%7.4f" % v for v in con[ie])
    if name == "jumps":
        assert np.abs(con).max() > np.pi  # the synthetic set really exercises the unwrapping


def test_conphas_edge_cases():
    assert orc.conphas(np.zeros((1, 3))).shape == (1, 3)
    x = np.array([[1.5], [-1.5], [1.5], [-1.5]])  # alternating +-3 rad jumps: ifak goes +1, 0, +1
    out = orc.conphas(x)[:, 0]
    assert np.allclose(out, [1.5, -1.5 + np.pi, 1.5, -1.5 + np.pi])
    y = np.array([[0.0], [np.pi / 2]])  # dif == pi/2 exactly counts as a jump (>=)
    assert np.allclose(orc.conphas(y)[:, 0], [0.0, np.pi / 2 - np.pi])


def test_fortran_format_helpers():
    assert fio.ffmt(20.0, 9, 4) == "  20.0000"
    assert fio.ffmt(-0.958, 8, 5) == "-0.95800"
    assert fio.ffmt(0.05501, 8, 5) == " 0.05501"
    assert fio.ffmt(-0.958, 7, 4) == "-0.9580"
    assert fio.ffmt(-12.345678, 8, 5) == "********"
    assert fio.ffmt(-0.5, 6, 4) == "-.5000"
    assert fio.dfmt(20.0, 12, 4) == "  0.2000D+02"
    assert fio.dfmt(-149.8956, 14, 6) == " -0.149896D+03"
    assert fio.fread("  0.2000D+02", 4) == 20.0 and fio.fread("     12345", 3) == 12.345 and fio.fread("      ", 2) == 0.0
    assert fio.list_directed_real(1.0, 8) == "   1.0000000000000000     "
    assert fio.list_directed_real(1.0, 4) == "   1.00000000    "
    assert fio.list_directed_real(0.00386, 4).strip().endswith("E-03")
