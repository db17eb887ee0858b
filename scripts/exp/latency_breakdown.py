import os, sys, time, tempfile
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np
from phaseshifts_b200 import api, workloads, libphsh
RE = os.path.join(os.path.dirname(__file__), "..", "..", "tests", "golden", "Re0001")
d = tempfile.mkdtemp(); P = lambda n: os.path.join(d, n)
def med(fn, n=30, warm=3):
    for _ in range(warm): fn()
    ts = []
    for _ in range(n):
        t0 = time.perf_counter(); fn(); ts.append(time.perf_counter() - t0)
    return 1e3 * float(np.median(ts))
w = workloads.c1_re_fixture(); g = w["grid"]
print("rel_batch_host C1        %.3f ms" % med(lambda: api.phsh_rel_batch(w["pot"], w["jri"], g["e_l"], 12, energies_l0=g["e_l0"])))
import torch
pt = torch.from_numpy(w["pot"]).cuda()
def dev():
    api.phsh_rel_batch(pt, w["jri"], g["e_l"], 12, energies_l0=g["e_l0"]); torch.cuda.synchronize()
print("rel_batch_device C1+sync %.3f ms" % med(dev))
print("phsh_rel_file C1         %.3f ms" % med(lambda: api.phsh_rel_file(os.path.join(RE, "mufftin_Re_slab.d"), P("po"), P("dp"), P("in"))))
c6 = workloads.c6_cavpot(S=1); b = c6["rela"]
print("rela_density             %.3f ms" % med(lambda: api.cavpot_rela_density(b["rmin"], b["rmax"], b["rs"], 0)))
rho, nx = api.cavpot_rela_density(b["rmin"], b["rmax"], b["rs"], 0)
pk = api.cavpot_pack(c6["structures"])
print("cavpot_batch_host S=1    %.3f ms" % med(lambda: api.cavpot_batch(pk, rho[None, :], [nx])))
print("cavpot_file              %.3f ms" % med(lambda: api.cavpot_file("", 0, os.path.join(RE, "atomic_Re.i"), os.path.join(RE, "cluster_Re_bulk.i"), P("m.d"), P("o.i"), P("i.txt"))))
from phaseshifts_b200.conphas import Conphas
os.chdir(d)
print("split_phasout            %.3f ms" % med(lambda: Conphas.split_phasout(P("po"), output_filenames=[P("A.ph"), P("B.ph")])))
def con():
    c = Conphas(input_files=[P("A.ph")], output_file=P("A.phs"), formatting="cleed", lmax=10); c.calculate()
print("conphas_files one input  %.3f ms" % med(con))
