// C entry points of the CAVPOT oracle -- TEST INFRASTRUCTURE, NOT PRODUCT CODE (see cavpot_oracle.hpp).
#include "cavpot_oracle.hpp"

using namespace cavpot_oracle;

namespace {

template <class Real>
int rela_impl(double rmin, double rmax, int nr, const double* rs, int iprint, double* rho_out, int* nx_out) {
    std::vector<Real> rx;
    loucks_grid<Real>(rx);
    std::vector<Real> rsv(nr + 1, Real(0));
    for (int i = 1; i <= nr; ++i) rsv[i] = static_cast<Real>(rs[i - 1]);
    Density<Real> d = rela_density<Real>(static_cast<Real>(rmin), static_cast<Real>(rmax), nr, rsv.data(), rx, iprint);
    for (int i = 1; i <= NGRID; ++i) rho_out[i - 1] = static_cast<double>(d.rho[i]);
    *nx_out = d.nx;
    return 0;
}

template <class Real>
int hsin_impl(double z, int nm, int nc, const int* lc, const int* n, const double* frac, const double* wf, int wf_stride,
              int iprint, double* rho_out, int* nx_out) {
    std::vector<Real> rx;
    loucks_grid<Real>(rx);
    std::vector<HsState<Real>> st(nc);
    for (int ic = 0; ic < nc; ++ic) {
        st[ic].lc = lc[ic];
        st[ic].n = n[ic];
        st[ic].frac = static_cast<Real>(frac[ic]);
        st[ic].wf.assign(n[ic] + 1, Real(0));
        for (int i = 1; i <= n[ic]; ++i) st[ic].wf[i] = static_cast<Real>(wf[static_cast<size_t>(ic) * wf_stride + i - 1]);
    }
    Density<Real> d = hsin_density<Real>(static_cast<Real>(z), nm, st, rx, iprint);
    for (int i = 1; i <= NGRID; ++i) rho_out[i - 1] = static_cast<double>(d.rho[i]);
    *nx_out = d.nx;
    return 0;
}

template <class Real>
int clemin_impl(int nbasis, const int* ns, const int* nt, const double* ex, int b_stride, int nc, const int* basis,
                const int* lc, const double* fac, const double* frac, double* rho_out, int* nx_out) {
    std::vector<Real> rx;
    loucks_grid<Real>(rx);
    std::vector<ClemBasis<Real>> bases(nbasis);
    for (int b = 0; b < nbasis; ++b)
        for (int k = 0; k < ns[b]; ++k) {
            bases[b].nt.push_back(nt[b * b_stride + k]);
            bases[b].ex.push_back(static_cast<Real>(ex[b * b_stride + k]));
        }
    std::vector<ClemState<Real>> st(nc);
    for (int ic = 0; ic < nc; ++ic) {
        st[ic].basis = basis[ic];
        st[ic].lc = lc[ic];
        st[ic].frac = static_cast<Real>(frac[ic]);
        for (int k = 0; k < ns[basis[ic]]; ++k) st[ic].fac.push_back(static_cast<Real>(fac[ic * b_stride + k]));
    }
    Density<Real> d = clemin_density<Real>(bases, st, rx);
    for (int i = 1; i <= NGRID; ++i) rho_out[i - 1] = static_cast<double>(d.rho[i]);
    *nx_out = d.nx;
    return 0;
}

template <class Real>
int cavpot_impl(int asbuilt, double spa, const double* rc_in, int nr, const int* nrr, const double* z, const double* zc,
                const double* rmt, const double* rk_in, int nform, double alpha, int nh, int slab, double esht,
                const double* rho, const int* nx, double* mtz_out, int* npts, double* out_r, double* out_v, int* jrmt,
                int* ncon, double* ad, int* na, int* ia, double* vh, double* vs, double* vmad, int* stats) {
    Structure<Real> s;
    const Real SPA = static_cast<Real>(spa);
    for (int j = 1; j <= 3; ++j)
        for (int i = 1; i <= 3; ++i) s.rc[i][j] = SPA * static_cast<Real>(rc_in[3 * (j - 1) + (i - 1)]);
    s.nr = nr;
    s.nrr.assign(nr + 1, 0);
    s.z.assign(nr + 1, Real(0));
    s.zc = s.z;
    s.rmt = s.z;
    s.zmv.assign(1, Real(0));
    int n = 0;
    for (int ir = 1; ir <= nr; ++ir) {
        s.nrr[ir] = nrr[ir - 1];
        s.z[ir] = static_cast<Real>(z[ir - 1]);
        s.zc[ir] = static_cast<Real>(zc[ir - 1]);
        s.rmt[ir] = static_cast<Real>(rmt[ir - 1]);
        for (int k = 0; k < nrr[ir - 1]; ++k) s.zmv.push_back(s.zc[ir]);
        n += nrr[ir - 1];
    }
    s.n = n;
    s.rk.resize(3 * n);
    for (int k = 0; k < 3 * n; ++k) s.rk[k] = SPA * static_cast<Real>(rk_in[k]);
    s.nform = nform;
    s.alpha = static_cast<Real>(alpha);
    s.nh = nh;
    s.slab = slab;
    s.esht = static_cast<Real>(esht);
    std::vector<Density<Real>> dens(nr + 1);
    for (int ir = 1; ir <= nr; ++ir) {
        dens[ir].rho.assign(NGRID + 2, Real(0));
        for (int i = 1; i <= NGRID; ++i) dens[ir].rho[i] = static_cast<Real>(rho[(ir - 1) * NGRID + i - 1]);
        dens[ir].nx = nx[ir - 1];
    }
    Result<Real> R = cavpot<Real>(s, dens, asbuilt != 0);
    mtz_out[0] = static_cast<double>(R.vhar);
    mtz_out[1] = static_cast<double>(R.vex);
    mtz_out[2] = static_cast<double>(R.vhar + R.vex);  // what CAVPOT returns / writes to unit 10
    mtz_out[3] = static_cast<double>(R.av);
    mtz_out[4] = static_cast<double>(R.rws);
    for (int ir = 1; ir <= nr; ++ir) {
        npts[ir - 1] = R.npts[ir];
        for (int k = 0; k < R.npts[ir]; ++k) {
            out_r[(ir - 1) * 551 + k] = static_cast<double>(R.out_r[ir][k]);
            out_v[(ir - 1) * 551 + k] = static_cast<double>(R.out_v[ir][k]);
        }
        if (jrmt) jrmt[ir - 1] = R.jrmt[ir];
        if (ncon) ncon[ir - 1] = R.ncon.empty() ? 0 : R.ncon[ir];
        for (int ic = 1; ic <= MC; ++ic) {
            if (ad) ad[(ir - 1) * MC + ic - 1] = R.shells.ad.empty() ? 0.0 : static_cast<double>(R.shells.ad[ir][ic]);
            if (na) na[(ir - 1) * MC + ic - 1] = R.shells.na.empty() ? 0 : R.shells.na[ir][ic];
            if (ia) ia[(ir - 1) * MC + ic - 1] = R.shells.ia.empty() ? 0 : R.shells.ia[ir][ic];
        }
        for (int i = 1; i <= NGRID; ++i) {
            if (vh) vh[(ir - 1) * NGRID + i - 1] = static_cast<double>(R.vh[ir][i]);
            if (vs) vs[(ir - 1) * NGRID + i - 1] = static_cast<double>(R.vs[ir][i]);
            if (vmad) vmad[(ir - 1) * NGRID + i - 1] = static_cast<double>(R.vmad[ir][i]);
        }
    }
    if (stats) {
        stats[0] = R.mtz_stats.nint;
        stats[1] = R.mtz_stats.npoint;
        stats[2] = R.jrws;
    }
    return 0;
}

}  // namespace

extern "C" {

// RELA (phsh1.f:1120-1170): density given on the atomic program's logarithmic mesh -> Loucks mesh.
int orc_rela_density(int real32, double rmin, double rmax, int nr, const double* rs, int iprint, double* rho_out, int* nx_out) {
    return real32 ? rela_impl<float>(rmin, rmax, nr, rs, iprint, rho_out, nx_out)
                  : rela_impl<double>(rmin, rmax, nr, rs, iprint, rho_out, nx_out);
}

// HSIN (phsh1.f:572-635): Herman-Skillman orbitals wf[ic][0..n[ic]) on the H-S mesh -> density on the Loucks mesh.
int orc_hsin_density(int real32, double z, int nm, int nc, const int* lc, const int* n, const double* frac, const double* wf,
                     int wf_stride, int iprint, double* rho_out, int* nx_out) {
    return real32 ? hsin_impl<float>(z, nm, nc, lc, n, frac, wf, wf_stride, iprint, rho_out, nx_out)
                  : hsin_impl<double>(z, nm, nc, lc, n, frac, wf, wf_stride, iprint, rho_out, nx_out);
}

// CLEMIN (phsh1.f:493-570): Clementi orbitals; basis b has ns[b] (nt, ex) pairs, state ic uses basis[ic] with fac[ic][..].
int orc_clemin_density(int real32, int nbasis, const int* ns, const int* nt, const double* ex, int b_stride, int nc,
                       const int* basis, const int* lc, const double* fac, const double* frac, double* rho_out, int* nx_out) {
    return real32 ? clemin_impl<float>(nbasis, ns, nt, ex, b_stride, nc, basis, lc, fac, frac, rho_out, nx_out)
                  : clemin_impl<double>(nbasis, ns, nt, ex, b_stride, nc, basis, lc, fac, frac, rho_out, nx_out);
}

// CAVPOT from the densities on (phsh1.f:188-426).  rc_in[3*(j-1)+i-1] = RC(I,J) and rk_in[3*(a-1)+i-1] = RK(I,a) in units of
// spa, exactly as read from cluster.i.  out_r/out_v are [nr][551].
int orc_cavpot(int real32, int asbuilt, double spa, const double* rc_in, int nr, const int* nrr, const double* z,
               const double* zc, const double* rmt, const double* rk_in, int nform, double alpha, int nh, int slab, double esht,
               const double* rho, const int* nx, double* mtz_out, int* npts, double* out_r, double* out_v, int* jrmt,
               int* ncon, double* ad, int* na, int* ia, double* vh, double* vs, double* vmad, int* stats) {
    return real32 ? cavpot_impl<float>(asbuilt, spa, rc_in, nr, nrr, z, zc, rmt, rk_in, nform, alpha, nh, slab, esht, rho, nx,
                                       mtz_out, npts, out_r, out_v, jrmt, ncon, ad, na, ia, vh, vs, vmad, stats)
                  : cavpot_impl<double>(asbuilt, spa, rc_in, nr, nrr, z, zc, rmt, rk_in, nform, alpha, nh, slab, esht, rho, nx,
                                        mtz_out, npts, out_r, out_v, jrmt, ncon, ad, na, ia, vh, vs, vmad, stats);
}
}
