"""The product must not route through the oracle or any CPU fallback: nothing under phaseshifts_b200/ (Python or
native) may import, link or execute anything under oracle/, and the hot-path translation units contain no host
implementation of the integrators.  CPU only."""
import os
import re
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "phaseshifts_b200")


def _sources():
    for d, _, files in os.walk(PKG):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".h", ".inc")):
                yield os.path.join(d, f)


def test_package_never_references_the_oracle():
    pat = re.compile(r"(^\s*(from|import)\s+oracle\b)|(oracle/)|(libphsh_oracle)|(phsh_oracle)", re.M)
    offenders = []
    for path in _sources():
        txt = open(path, encoding="utf-8").read()
        for m in pat.finditer(txt):
            line = txt[: m.start()].count("\n") + 1
            snippet = txt.splitlines()[line - 1]
            if "see oracle/" in snippet or "oracle/phsh_oracle.hpp" in snippet and snippet.lstrip().startswith(("//", "#", "*")):
                continue  # a comment pointing the reader at the checker is fine
            offenders.append("%s:%d: %s" % (os.path.relpath(path, ROOT), line, snippet.strip()))
    assert not offenders, "\n".join(offenders)


def test_shared_library_does_not_link_the_oracle():
    from phaseshifts_b200 import build
    so = build.build()
    out = subprocess.run(["ldd", so], capture_output=True, text=True).stdout
    assert "oracle" not in out
    syms = subprocess.run(["nm", "-D", "--defined-only", so], capture_output=True, text=True).stdout
    assert "orc_" not in syms


def test_integrators_exist_only_as_device_code():
    """dlgkap / Numerov / Hamming are __device__/__global__ functions; the host translation units hold none."""
    rel = open(os.path.join(PKG, "csrc", "phsh_rel.cuh")).read()
    assert "__device__ __forceinline__ double dlgkap_dev" in rel and "__global__" in rel
    for host_tu in ("phsh_files.cpp", "phsh_conphas_files.cpp"):
        txt = open(os.path.join(PKG, "csrc", host_tu)).read()
        assert "milne" not in txt.lower() and "dlgkap" not in txt.lower() and "numerov" not in txt.lower()
        assert "phsh_b200_" in txt  # they only call the batched GPU entry points
