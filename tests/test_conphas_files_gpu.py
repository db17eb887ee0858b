"""conphas file layer (Conphas.split_phasout / load_data / calculate + cleed / curve / viperleed / plain writers)
through the C ABI, byte for byte against files written by the reference's own ``phaseshifts.conphas.Conphas``
(tests/golden/conphas, see tests/golden/make_golden.py).  The pi-jump removal runs in ``conphas_kernel``."""
import filecmp
import os
import shutil

import pytest

pytestmark = pytest.mark.gpu
CON = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "conphas")


def _lines(p):
    return open(p).read().splitlines(keepends=True)


@pytest.mark.parametrize("name,lmax", [("Re_A", 10), ("jumps", 12), ("jumps", 5)])
@pytest.mark.parametrize("fmt", ["cleed", "curve", None, "viperleed"])
def test_calculate_matches_reference_conphas_files(tmp_path, name, lmax, fmt):
    from phaseshifts_b200.conphas import Conphas
    src = str(tmp_path / (name + ".ph"))
    shutil.copyfile(os.path.join(CON, name + ".ph"), src)
    golden = os.path.join(CON, "%s_lmax%d_%s.out" % (name, lmax, fmt))
    # the reference derives the "# phase shift curve:" title from the output file name
    dst = str(tmp_path / os.path.basename(golden))
    c = Conphas(input_files=[src], output_file=dst, formatting=fmt, lmax=lmax, user="tester", timestamp="T")
    c.calculate()
    got, exp = _lines(dst), _lines(golden)
    if fmt == "cleed":
        assert got[0] == exp[0].rstrip("\n") + " (calculated by tester on T)\n"
        got, exp = got[1:], exp[1:]
    elif fmt == "viperleed":
        assert got[0] == exp[0].rstrip("\n") + " T\n"
        got, exp = got[1:], exp[1:]
    assert got == exp
    assert filecmp.cmp(str(tmp_path / ("dataph_%s.d" % name)), os.path.join(CON, "dataph_%s_lmax%d.d" % (name, lmax)),
                       shallow=False)


def test_split_phasout_and_two_block_output(tmp_path):
    from phaseshifts_b200.conphas import Conphas
    both = str(tmp_path / "phasout")
    shutil.copyfile(os.path.join(CON, "two_blocks_phasout"), both)
    names = Conphas.split_phasout(both, [str(tmp_path / "split_A.ph"), str(tmp_path / "split_B.ph")])
    assert [os.path.basename(n) for n in names] == ["split_A.ph", "split_B.ph"]
    for n in names:
        assert filecmp.cmp(n, os.path.join(CON, os.path.basename(n)), shallow=False)
    dst = str(tmp_path / "two.out")
    Conphas(input_files=names, output_file=dst, formatting="cleed", lmax=10, user="u", timestamp="t").calculate()
    got, exp = _lines(dst), _lines(os.path.join(CON, "two_blocks_lmax10_cleed.out"))
    assert got[0].startswith("114 10 neng lmax (calculated by u on t)") and got[1:] == exp[1:]
    # the reference leaves every input but the last one wrapped in the combined file (conphas.py:434); opt-in fix:
    fixed = str(tmp_path / "two_fixed.out")
    Conphas(input_files=names, output_file=fixed, formatting="cleed", lmax=10, user="u", timestamp="t",
            unwrap_all_inputs=True).calculate()
    fx = _lines(fixed)
    row = lambda ln: [float(x) for x in ln.replace("-", " -").split()]
    assert any(row(a)[:11] != row(b)[:11] for a, b in zip(got[2::2], fx[2::2]))    # block A differs ...
    assert all(row(a)[11:] == row(b)[11:] for a, b in zip(got[2::2], fx[2::2]))    # ... block B does not
    assert all(row(b)[:11] == row(b)[11:] for b in fx[2::2])                      # same potential, same phases
    # guessed names: last blank-separated word of the title + ".ph" (conphas.py:232-234); PHSH_REL's blank-padded
    # 4A4 titles make that "" -> ".ph" for every section, exactly as in the reference
    cwd = os.getcwd()
    os.chdir(tmp_path)
    try:
        guessed = Conphas.split_phasout(both)
    finally:
        os.chdir(cwd)
    assert guessed == [".ph", ".ph"]
    by_str = Conphas.split_phasout(both, str(tmp_path / "atom.ph"))
    assert [os.path.basename(n) for n in by_str] == ["atom_0.ph", "atom_1.ph"]


def test_error_behaviour(tmp_path):
    from phaseshifts_b200.conphas import Conphas
    src = str(tmp_path / "jumps.ph")
    shutil.copyfile(os.path.join(CON, "jumps.ph"), src)
    with pytest.raises(IndexError):  # lmax beyond the phases in the file (IndexError in the reference too)
        Conphas(input_files=[src], output_file=str(tmp_path / "o"), formatting="cleed", lmax=15).calculate()
    bad = str(tmp_path / "bad.ph")
    lines = _lines(src)
    open(bad, "w").write("".join(lines[:-3]))  # truncated section
    with pytest.raises(IndexError):
        Conphas(input_files=[bad], output_file=str(tmp_path / "o"), formatting="cleed", lmax=5).calculate()
    with pytest.raises(IOError):
        Conphas.split_phasout(str(tmp_path / "missing"))
    c = Conphas(input_files=[src, str(tmp_path / "nope.ph")], output_file=str(tmp_path / "o"), lmax=40)
    assert c.input_files == [src] and not hasattr(c, "lmax")  # same constructor filtering as conphas.py:95-98
    assert c.load_data(src)[2:4] == (40, 12)
