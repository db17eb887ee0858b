"""GPU cavpot vs oracle<double> on the Re(0001) fixture and a few perturbed structures (scratch)."""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from oracle import cavpot_io as cio
from phaseshifts_b200 import api

G = os.path.join(os.path.dirname(__file__), "..", "..", "tests", "golden", "Re0001")
cl = cio.read_cluster(os.path.join(G, "cluster_Re_bulk.i"), real32=False)
blocks = cio.read_atomic(os.path.join(G, "atomic_Re.i"), cl["nr"], real32=False)
rho = np.zeros((cl["nr"], 250)); nx = np.zeros(cl["nr"], dtype=np.int32)
for i, b in enumerate(blocks):
    rho[i], nx[i] = cio.rela_density(b, real32=False)
    g_rho, g_nx = api.cavpot_rela_density(b["rmin"], b["rmax"], b["rs"], b["iprint"])
    print("rela", i, "nx", nx[i], g_nx, "max abs diff", np.abs(g_rho - rho[i]).max(), "max", np.abs(rho[i]).max())

def as_struct(c, **kw):
    d = dict(spa=c["spa"], rc=c["rc"], nrr=c["nrr"], z=c["z"], zc=c["zc"], rmt=c["rmt"], rk=c["rk"], dens=np.arange(c["nr"]),
             alpha=c["alpha"], nh=c["nh"], nform=c["nform"])
    d.update(kw)
    return d

def compare(c, slab=0, esht=0.0, tag=""):
    ref = cio.cavpot(c, rho, nx, slab=slab, esht=esht, real32=False)
    t = time.time()
    out = api.cavpot_batch([as_struct(c, slab=slab, esht=esht)], rho, nx, diagnostics=True)
    dt = time.time() - t
    n = ref["npts"]
    err = max(np.abs(out["pot"][i, :n[i]] - ref["v"][i, :n[i]]).max() for i in range(c["nr"]))
    print(tag, "mtz", out["mtz"][0], ref["mtz"], "dvhar", out["vhar"][0] - ref["vhar"], "dvex", out["vex"][0] - ref["vex"],
          "npts", out["npts"], n, "ncon", out["ncon"], ref["ncon"], "stats", out["stats"][0], ref["nint"], ref["npoint"],
          "pot err", err, "vh err", np.abs(out["vh"] - ref["vh"])[:, :190].max(), "vs err", np.abs(out["vs"] - ref["vs"])[:, :190].max(),
          "shell", np.abs(out["shell_ad"] - ref["ad"]).max(), (out["shell_na"] != ref["na"]).sum(), (out["shell_ia"] != ref["ia"]).sum(),
          "t %.1f ms" % (dt * 1e3))
    return err

compare(cl, tag="bulk nform2")
for nform in (0, 1):
    c2 = dict(cl); c2["nform"] = nform
    compare(c2, tag="bulk nform%d" % nform)
c3 = cio.read_cluster(os.path.join(G, "cluster_Re_slab.i"), real32=False)
compare(c3, slab=1, esht=-1.64, tag="slab")
c4 = dict(cl); c4["nh"] = 0; c4["nr"] = 1
c4.update(nrr=np.array([2], dtype=np.int32), z=cl["z"][:1], zc=cl["zc"][:1], rmt=cl["rmt"][:1], names=cl["names"][:1])
ref = cio.cavpot(c4, rho[:1], nx[:1], real32=False)
out = api.cavpot_batch([as_struct(c4, dens=[0])], rho[:1], nx[:1], diagnostics=True)
print("mtzm", out["mtz"], ref["mtz"], np.abs(out["pot"][0, :ref["npts"][0]] - ref["v"][0, :ref["npts"][0]]).max())
c5 = dict(cl); c5["zc"] = np.array([1.0, -1.0])
compare(c5, tag="ionic (MAD)")
# batch timing
rng = np.random.default_rng(1)
S = 512
structs = []
for s in range(S):
    c = dict(cl)
    c["rmt"] = cl["rmt"] * (1 + 0.1 * rng.uniform(-1, 1))
    c["rk"] = cl["rk"] + 0.01 * rng.normal(size=cl["rk"].shape)
    structs.append(as_struct(c))
api.cavpot_batch(structs[:4], rho, nx)
t = time.time(); out = api.cavpot_batch(structs, rho, nx); dt = time.time() - t
print("batch", S, "structures: %.1f ms (%.3f ms per structure)" % (dt * 1e3, dt * 1e3 / S))
t = time.time()
for s in range(8):
    c = dict(cl); c["rmt"] = structs[s]["rmt"]; c["rk"] = structs[s]["rk"]
    r = cio.cavpot(c, rho, nx, real32=False)
    e = np.abs(out["pot"][2 * s, :r["npts"][0]] - r["v"][0, :r["npts"][0]]).max()
    assert (out["npts"][2 * s:2 * s + 2] == r["npts"]).all()
    print(" s", s, "err", e, "mtz", out["mtz"][s] - r["mtz"])
print("oracle %.2f ms per structure" % ((time.time() - t) * 1e3 / 8))
