"""The four-file interface of libphsh served by the GPU library, against the oracle's file layer and the
reference's golden Re(0001) outputs.  Numeric gate on the 5-decimal text: one unit of the last printed
place (the FP64 kernel vs the FP32-as-written oracle differ by <= 1.4e-6 before rounding)."""
import filecmp
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))
RE = os.path.join(HERE, "golden", "Re0001")


def _numbers(path, skip=2):
    vals = []
    with open(path) as f:
        for ln in f.readlines():
            s = ln.replace("-", " -").split()
            try:
                vals.append([float(x) for x in s])
            except ValueError:
                vals.append(None)
    return vals


def _golden_dataph():
    blocks = open(os.path.join(RE, "dataph_Re.d")).read().split('"L=')[1:]
    return [np.array([[float(x) for x in ln.split()] for ln in b.strip().split("\n")[1:] if ln.strip()]) for b in blocks]


def test_rel_file_real32_vs_oracle_file(tmp_path):
    """REAL32 epilogue (matching arithmetic rounded to REAL*4 like the Fortran declares it) vs the FP32-as-written
    oracle run through its own file layer: identical layout, inpdat byte-identical, phases equal to one unit of
    the last printed place (libm's sinf/cosf/atanf are not correctly rounded, so single float ulps differ)."""
    from oracle import fortran_io as fio
    from phaseshifts_b200 import api
    muff = os.path.join(RE, "mufftin_Re_slab.d")
    g = [str(tmp_path / n) for n in ("g_phasout", "g_dataph", "g_inpdat")]
    o = [str(tmp_path / n) for n in ("o_phasout", "o_dataph", "o_inpdat")]
    api.phsh_rel_file(muff, *g, real32=True)
    fio.phsh_rel_file(muff, *o, real32=True)
    assert filecmp.cmp(g[2], o[2], shallow=False)
    la, lb = open(g[0]).read().splitlines(), open(o[0]).read().splitlines()
    assert len(la) == len(lb) == 2 * (2 + 57 * 2)
    ndiff = 0
    for x, y in zip(la, lb):
        assert len(x) == len(y)
        if x != y:
            ndiff += 1
            xv = np.array([float(t) for t in x.replace("-", " -").split()])
            yv = np.array([float(t) for t in y.replace("-", " -").split()])
            assert np.abs(xv - yv).max() <= 1.0001e-5
    assert ndiff <= 0.05 * len(la)
    da, db = open(g[1]).read().splitlines(), open(o[1]).read().splitlines()
    assert len(da) == len(db)
    for x, y in zip(da, db):
        assert len(x) == len(y)
        if x != y:
            assert np.abs(np.array(x.split(), dtype=float) - np.array(y.split(), dtype=float)).max() < 2e-6


def test_rel_file_default_vs_golden_outputs(tmp_path):
    """FP64 file path -> reference-format phasout -> conphas -> the reference's own dataph_Re.d / leedph_Re.d."""
    from oracle import fortran_io as fio
    from phaseshifts_b200 import api
    muff = os.path.join(RE, "mufftin_Re_slab.d")
    ph, dp, ip = [str(tmp_path / n) for n in ("phasout", "dataph", "inpdat")]
    api.phsh_rel_file(muff, ph, dp, ip)
    txt = open(ph).read()
    secs = ["RELATIVISTIC" + s for s in txt.split("RELATIVISTIC")[1:]]
    assert len(secs) == 2 and secs[0].splitlines()[0] == "RELATIVISTIC PHASE SHIFTS FOR RHENIUM A       "
    assert secs[0].splitlines()[1] == "   20.0000   5.0000   57   12"
    sec = str(tmp_path / "A.ph")
    open(sec, "w").write(secs[0])
    en, con = fio.conphas_file(sec, 10)
    gold = _golden_dataph()
    for l in range(11):
        assert np.abs(en / 27.21 - gold[l][:, 0]).max() < 2e-6
        assert np.abs(con[:, l] - gold[l][:, 1]).max() < (1.05e-5 if l != 8 else 1.0e-4)
    leed = [ln for ln in open(os.path.join(RE, "leedph_Re.d")).read().splitlines()]
    for ie in range(57):
        row = [float(x) for x in leed[2 * ie + 1].replace("-", " -").split()]
        assert np.abs(np.array(row) - con[ie]).max() <= 1.0001e-4


def test_rel_file_caps_and_errors(tmp_path):
    from phaseshifts_b200 import api, PhshB200Error
    muff = os.path.join(RE, "mufftin_Re_slab.d")
    lines = open(muff).read().splitlines(keepends=True)
    # what Wrapper._apply_energy_range writes (phsh.py:276): 1 eV steps to 600 eV => N=581 in the header, 250 rows
    lines[1] = ("%12.4f%12.4f%12.4f    %3i    %12.4f\n" % (20.0, 1.0, 600.0, 12, 0.0))
    m2 = str(tmp_path / "m2.d")
    open(m2, "w").write("".join(lines))
    ph = str(tmp_path / "ph")
    api.phsh_rel_file(m2, ph, str(tmp_path / "dp"), str(tmp_path / "ip"))
    secs = open(ph).read().split("RELATIVISTIC")[1:]
    a = secs[0].splitlines()
    assert a[1] == "   20.0000   1.0000  581   12"
    assert len(a) == 2 + 250 * 2
    assert len(secs[1].splitlines()) == 2 + 57 * 2  # second block keeps its own header (phsh.py:272-276)
    with pytest.raises(PhshB200Error):
        api.phsh_rel_file(str(tmp_path / "missing.d"), ph, ph + "1", ph + "2")


def _cav_atoms():
    from phaseshifts_b200 import workloads
    w = workloads.c2_oxide(ne=4, nl=12)
    names = ["OXYGEN", "ALUMINIUM", "TITANIUM", "STRONTIUM"]
    return [dict(name=names[i], z=[8, 13, 22, 38][i], rmt=w["rmt"][i], mtz=-0.8, rx=w["RX"][i, :w["ntab"][i]],
                 v=w["V"][i, :w["ntab"][i]]) for i in range(4)]


def test_cav_file_matches_oracle_file(tmp_path):
    from oracle import fortran_io as fio
    from phaseshifts_b200 import api
    muff = str(tmp_path / "mufftin_cav.d")
    fio.write_mufftin_cav(muff, _cav_atoms())
    for asb, hdr in ((False, False), (False, True), (True, False)):
        g = [str(tmp_path / ("g%d%d_" % (asb, hdr) + n)) for n in ("phasout", "dataph", "zph")]
        o = [str(tmp_path / ("o%d%d_" % (asb, hdr) + n)) for n in ("phasout", "dataph", "zph")]
        api.phsh_cav_file(muff, *g, asbuilt=asb, conphas_header=hdr)
        fio.phsh_cav_file(muff, *o, asbuilt=asb, conphas_header=hdr)
        for a, b in zip(g[:2], o[:2]):
            ga, gb = _numbers(a), _numbers(b)
            assert len(ga) == len(gb)
            for x, y in zip(ga, gb):
                assert (x is None) == (y is None)
                if x is not None:
                    assert np.abs(np.array(x) - np.array(y)).max() <= 1.0001e-4 if x else True
        assert open(g[0]).read().splitlines()[:1] == open(o[0]).read().splitlines()[:1]


def test_cav_file_conphas_header_is_parseable(tmp_path):
    from oracle import fortran_io as fio
    from phaseshifts_b200 import api
    muff = str(tmp_path / "mufftin_cav.d")
    fio.write_mufftin_cav(muff, _cav_atoms())
    ph = str(tmp_path / "phasout")
    api.phsh_cav_file(muff, ph, str(tmp_path / "dataph"), str(tmp_path / "zph"), conphas_header=True)
    secs = ["NON" + s for s in open(ph).read().split("NON")[1:]]
    assert len(secs) == 4
    assert secs[0].splitlines()[0] == "NON-RELATIVISTIC PHASE SHIFTS FOR OXYG    "
    one = str(tmp_path / "one.ph")
    open(one, "w").write(secs[1])
    en, con = fio.conphas_file(one, 10)
    assert en.shape == (45,) and con.shape == (45, 11) and abs(en[0] - 27.21) < 1e-3


def test_wil_file_matches_oracle_file(tmp_path):
    from oracle import fortran_io as fio
    from phaseshifts_b200 import api, workloads
    w = workloads.c5_wil(P=3, ne=4, nl=9, seed=21)
    atoms = []
    for i in range(3):
        n = w["nrows"][i] - 1
        atoms.append(dict(z=w["Z"][i], rt=round(w["RT"][i], 4), r=w["rs"][i, :n], rv=w["zs"][i, :n]))
    muff = str(tmp_path / "mufftin_wil.d")
    fio.write_mufftin_wil(muff, atoms)
    g = [str(tmp_path / ("g_" + n)) for n in ("phasout", "dataph", "zph")]
    o = [str(tmp_path / ("o_" + n)) for n in ("phasout", "dataph", "zph")]
    api.phsh_wil_file(muff, *g)
    fio.phsh_wil_file(muff, *o)
    ga, gb = open(g[0]).read().splitlines(), open(o[0]).read().splitlines()
    assert ga[0] == gb[0] == " BE CAREFUL ABOUT THE ORDER OF THE ELEMENTS"
    assert len(ga) == len(gb) == 1 + 30 * 4
    for x, y in zip(ga[1:], gb[1:]):
        xv = np.array([float(t) for t in x.replace("-", " -").split()])
        yv = np.array([float(t) for t in y.replace("-", " -").split()])
        assert np.abs(xv - yv).max() <= 1.0001e-4
