"""Latency of the whole file chain of one LEED calculation (the reference's C1-sized use): cluster.i + atomic.i ->
cavpot (bulk, then slab) -> phsh_rel -> split_phasout + conphas -> .phs, every step through this package's file
interface on the GPU; beside it the CPU oracle's file layers doing the same steps on one host thread."""
import json, os, sys, tempfile, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np
from phaseshifts_b200 import api, libphsh
from phaseshifts_b200.conphas import Conphas
from oracle import cavpot_io as cio, fortran_io as fio

RE = os.path.join(os.path.dirname(__file__), "..", "..", "tests", "golden", "Re0001")
d = tempfile.mkdtemp()
os.chdir(d)
P = lambda n: os.path.join(d, n)
atomic, bulk, slab = (os.path.join(RE, n) for n in ("atomic_Re.i", "cluster_Re_bulk.i", "cluster_Re_slab.i"))


def gpu_chain():
    t = {}
    t0 = time.perf_counter()
    mtz = libphsh.cavpot("", 0, atomic, bulk, P("b_mufftin.d"), P("bmtz.i"), P("b_info.txt"))
    t["cavpot_bulk"] = time.perf_counter() - t0
    t0 = time.perf_counter()
    libphsh.cavpot("%.6f" % mtz, 1, atomic, slab, P("s_mufftin.d"), P("mtz.i"), P("s_info.txt"))
    t["cavpot_slab"] = time.perf_counter() - t0
    t0 = time.perf_counter()
    libphsh.phsh_rel(P("s_mufftin.d"), P("phasout"), P("dataph"), P("inpdat"))
    t["phsh_rel"] = time.perf_counter() - t0
    t0 = time.perf_counter()
    files = Conphas.split_phasout(P("phasout"), output_filenames=[P("A.ph"), P("B.ph")])
    for f in files:
        c = Conphas(input_files=[f], output_file=f + "s", formatting="cleed", lmax=10)
        c.calculate()
    t["conphas"] = time.perf_counter() - t0
    return t


def cpu_chain():
    t = {}
    t0 = time.perf_counter()
    mtz, _, _ = cio.cavpot_file("", 0, atomic, bulk, P("ob_mufftin.d"), P("obmtz.i"), None, real32=False)
    t["cavpot_bulk"] = time.perf_counter() - t0
    t0 = time.perf_counter()
    cio.cavpot_file("%.6f" % mtz, 1, atomic, slab, P("os_mufftin.d"), P("omtz.i"), None, real32=False)
    t["cavpot_slab"] = time.perf_counter() - t0
    t0 = time.perf_counter()
    fio.phsh_rel_file(P("os_mufftin.d"), P("ophasout"), P("odataph"), P("oinpdat"), real32=False)
    t["phsh_rel"] = time.perf_counter() - t0
    t0 = time.perf_counter()
    txt = open(P("ophasout")).read()
    for k, s in enumerate(["RELATIVISTIC" + x for x in txt.split("RELATIVISTIC")[1:]]):
        open(P("o%d.ph" % k), "w").write(s)
        fio.conphas_file(P("o%d.ph" % k), 10)
    t["conphas"] = time.perf_counter() - t0
    return t


def reference_python_conphas():
    """The reference's OWN conphas step (phaseshifts.conphas.Conphas, pure Python, staged under baseline/_ref): the same
    work as the GPU chain's conphas step -- split phasout, parse, unwrap, write dataph_<name>.d and the cleed .phs --
    which the oracle's step above does not do (it parses and unwraps, and writes nothing)."""
    ref = os.path.join(os.path.dirname(__file__), "..", "..", "baseline", "_ref")
    if not os.path.isfile(os.path.join(ref, "phaseshifts", "conphas.py")):
        return None
    if ref not in sys.path:
        sys.path.insert(0, ref)
    from phaseshifts_b200 import backend
    backend.install(force_standin=True)
    from phaseshifts.conphas import Conphas as RefConphas
    import io, contextlib
    ts = []
    for _ in range(7):
        t0 = time.perf_counter()
        with contextlib.redirect_stdout(io.StringIO()):
            files = RefConphas.split_phasout(P("phasout"), output_filenames=[P("RA.ph"), P("RB.ph")])
            for f in files:
                RefConphas(input_files=[f], output_file=f + "s", formatting="cleed", lmax=10).calculate()
        ts.append(time.perf_counter() - t0)
    return float(np.median(ts) * 1e3)


for _ in range(3):
    gpu_chain()
g = [gpu_chain() for _ in range(15)]
c = [cpu_chain() for _ in range(3)]
med = lambda runs: {k: float(np.median([r[k] for r in runs]) * 1e3) for k in runs[0]}
gm, cm = med(g), med(c)
a = [float(x) for l in open(P("A.phs")).read().splitlines()[2::2] for x in l.replace("-", " -").split()]
print(json.dumps(dict(config="C1 chain: Re(0001) cluster -> cavpot x2 -> phsh_rel (2 x 57 x 13) -> conphas -> .phs, files in / files out",
                      gpu_ms=gm, gpu_total_ms=sum(gm.values()), cpu_oracle_ms=cm, cpu_oracle_total_ms=sum(cm.values()),
                      cpu_threads=1, phs_values=len(a),
                      reference_python_conphas_ms=reference_python_conphas(),
                      note="cpu_oracle_ms.conphas only parses and unwraps (no dataph_<name>.d, no .phs written); "
                           "reference_python_conphas_ms is the reference's own Conphas doing what gpu_ms.conphas does")))
