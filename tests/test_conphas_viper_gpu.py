"""The reference's own Conphas/ViPErLEED tests (reference tests/test_conphas_viper.py), re-stated against the mirror
class ``phaseshifts_b200.conphas.Conphas`` (C++ parser/writer, unwrap on the GPU): same hand-written PHASOUT files,
same expectations."""
import os

import pytest

pytestmark = pytest.mark.gpu


def _write_phasout(tmp, filename="sample.ph", lmax=1, emin=0.0, emax=5.0, nenergies=2):
    """Minimal PHASOUT file with configurable lmax and number of energies (test_conphas_viper.py:16-45)."""
    path = os.path.join(tmp, filename)
    energies = [emin + i * (emax - emin) / max(nenergies - 1, 1) for i in range(nenergies)]
    with open(path, "w") as f:
        f.write("HEADER\n")
        f.write(f"{emin:.1f} {emax:.1f} {nenergies:d} {lmax:d}\n")
        values, k = [], 1
        for e in energies:
            values.append(f"{e:.1f}")
            for _ in range(lmax + 1):
                values.append(f"{0.1 * k:.1f}")
                k += 1
        f.write(" ".join(values) + "\n")
    return path


def _lines(path):
    with open(path) as f:
        return [ln.strip() for ln in f.readlines() if ln.strip()]


def test_writes_viper_phaseshifts_format(tmp_path):
    from phaseshifts_b200.conphas import Conphas
    src = _write_phasout(str(tmp_path))
    out = str(tmp_path / "PHASESHIFTS")
    Conphas(input_files=[src], output_file=out, formatting="viper", lmax=1).calculate()
    lines = _lines(out)
    tok = lines[0].split()
    assert tok[0] == "1" and tok[1:5] == ["-10.17", "-0.08", "-74.19", "19.18"] and tok[5] == "phaseshifts"
    assert len(tok) >= 7
    assert abs(float(lines[1]) - 0.0) < 1e-4
    assert [float(v) for v in lines[2].split()] == [0.1, 0.2]
    assert abs(float(lines[3]) - 5.0 / 27.21) < 1e-4
    assert [float(v) for v in lines[4].split()] == [0.3, 0.4]


def test_writes_viper_phaseshifts_with_custom_v0_params(tmp_path):
    from phaseshifts_b200.conphas import Conphas
    src = _write_phasout(str(tmp_path))
    out = str(tmp_path / "PHASESHIFTS_CUSTOM")
    Conphas(input_files=[src], output_file=out, formatting="viper", lmax=1, v0_params=[1.1, 2.2, 3.3, 4.4]).calculate()
    tok = open(out).readline().split()
    assert tok[1:5] == ["1.10", "2.20", "3.30", "4.40"] and tok[5] == "phaseshifts"
    with pytest.raises(ValueError):
        Conphas(input_files=[src], output_file=out, formatting="viper", lmax=1, v0_params=[1.0, 2.0]).calculate()
    with pytest.raises(ValueError):
        Conphas(input_files=[src], output_file=out, formatting="viper", lmax=1, v0_params=["a", 1, 2, 3]).calculate()


def test_writes_viper_phaseshifts_multiline_for_large_lmax(tmp_path):
    from phaseshifts_b200.conphas import Conphas
    src = _write_phasout(str(tmp_path), filename="sample_lmax11.ph", lmax=11)
    out = str(tmp_path / "PHASESHIFTS_LMAX11")
    Conphas(input_files=[src], output_file=out, formatting="viper", lmax=11).calculate()
    lines = [ln.rstrip() for ln in open(out) if ln.strip()]
    assert len(lines) == 1 + 2 * (1 + 2)  # header + per energy: energy line, 10 phases, 2 phases
    assert len(lines[2].split()) == 10 and len(lines[3].split()) == 2


def test_writes_viper_phaseshifts_multiblock_for_multiple_sites(tmp_path):
    from phaseshifts_b200.conphas import Conphas
    a = _write_phasout(str(tmp_path), filename="sample_site1.ph", lmax=3)
    b = _write_phasout(str(tmp_path), filename="sample_site2.ph", lmax=3)
    out = str(tmp_path / "PHASESHIFTS_MULTISITE")
    Conphas(input_files=[a, b], output_file=out, formatting="viper", lmax=3).calculate()
    lines = _lines(out)
    assert lines[0].split()[0] == "2"
    assert len(lines) == 1 + 2 * (1 + 2)
