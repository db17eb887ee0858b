// phsh_oracle_capi.cpp -- C entry points of the CPU ORACLE (test infrastructure only;
// see the header of phsh_oracle.hpp).  Loaded through ctypes by oracle/oracle.py.
#include "phsh_oracle.hpp"

#include <atomic>
#include <thread>

using namespace phsh_oracle;

namespace {

template <class F>
void parallel_items(long n, int nthreads, F&& fn) {
    if (nthreads <= 1 || n <= 1) {
        for (long i = 0; i < n; ++i) fn(i, 0);
        return;
    }
    std::atomic<long> next{0};
    std::vector<std::thread> th;
    const int nt = static_cast<int>(std::min<long>(nthreads, n));
    for (int t = 0; t < nt; ++t)
        th.emplace_back([&, t]() {
            for (;;) {
                const long i = next.fetch_add(1);
                if (i >= n) break;
                fn(i, t);
            }
        });
    for (auto& x : th) x.join();
}

template <class Real>
void rel_batch_t(const double* pot, long pot_stride, const int* jri, int P, const double* e_l0, const double* e_l,
                 long e_stride, int NE, int lsm, const RelGrid& g, bool asbuilt, double* jf, double* jfk,
                 long long* counters, int nthreads) {
    const int L1 = lsm + 1;
    // |ZP| once per potential (phsh2.f:817-820)
    std::vector<double> abspot(static_cast<size_t>(P) * pot_stride);
    for (int p = 0; p < P; ++p)
        for (int j = 0; j < jri[p]; ++j) {
            double v = pot[p * pot_stride + j];
            if (v < 0.0) v = -v;
            abspot[p * pot_stride + j] = v;
        }
    const int nt = std::max(1, nthreads);
    std::vector<Counters> cnts(nt);
    const long nitems = static_cast<long>(P) * L1;
    parallel_items(nitems, nthreads, [&](long item, int tid) {
        const int p = static_cast<int>(item / L1);
        const int L = lsm - static_cast<int>(item % L1);  // costly l first
        std::vector<Real> col(static_cast<size_t>(NE) * L1), colk;
        if (jfk) colk.resize(static_cast<size_t>(NE) * L1 * 2);
        rel_column<Real>(abspot.data() + p * pot_stride, jri[p], e_l0 + p * e_stride, e_l + p * e_stride, NE, lsm, L,
                         g, asbuilt, col.data(), jfk ? colk.data() : nullptr, &cnts[tid]);
        for (int j = 0; j < NE; ++j) {
            const size_t o = (static_cast<size_t>(p) * NE + j) * L1 + L;
            jf[o] = static_cast<double>(col[j * L1 + L]);
            if (jfk) {
                jfk[o * 2 + 0] = static_cast<double>(colk[(j * L1 + L) * 2 + 0]);
                jfk[o * 2 + 1] = static_cast<double>(colk[(j * L1 + L) * 2 + 1]);
            }
        }
    });
    if (counters) {
        Counters tot;
        for (auto& c : cnts) tot.add(c);
        counters[0] = tot.calls;
        counters[1] = tot.milne_steps;
        counters[2] = tot.corr_evals;
    }
}

template <class Real>
int cav_batch_t(const double* V, const double* RX, long stride, const int* ntab, const double* rmt, int P,
                const double* e_ryd, int NE, int NL, bool asbuilt, double* phs, int nthreads) {
    std::atomic<int> rc{0};
    parallel_items(static_cast<long>(P) * NE, nthreads, [&](long item, int) {
        const int p = static_cast<int>(item / NE), ie = static_cast<int>(item % NE);
        const int n = ntab[p];
        std::vector<Real> v(n), rx(n), out(NL);
        for (int i = 0; i < n; ++i) {
            v[i] = static_cast<Real>(V[p * stride + i]);
            rx[i] = static_cast<Real>(RX[p * stride + i]);
        }
        const int r = ps<Real>(v.data(), rx.data(), n, static_cast<Real>(rmt[p]), static_cast<Real>(e_ryd[ie]),
                               out.data(), NL, asbuilt);
        if (r != 0) rc = r;
        for (int l = 0; l < NL; ++l) phs[(static_cast<size_t>(p) * NE + ie) * NL + l] = r ? NAN : static_cast<double>(out[l]);
    });
    return rc.load();
}

template <class Real>
void wil_batch_t(const double* rs, const double* zs, long stride, const int* nrows, const double* Z, const double* RT,
                 int P, const double* e, int NE, int NL, int NR, double* del, double* raw, double* grid_out,
                 int nthreads) {
    parallel_items(P, nthreads, [&](long p, int) {
        const int n = nrows[p];
        std::vector<Real> r(n), z(n), ee(NE), d(static_cast<size_t>(NE) * NL), rw(static_cast<size_t>(NE) * NL);
        for (int i = 0; i < n; ++i) {
            r[i] = static_cast<Real>(rs[p * stride + i]);
            z[i] = static_cast<Real>(zs[p * stride + i]);
        }
        for (int i = 0; i < NE; ++i) ee[i] = static_cast<Real>(e[i]);
        WilAtom<Real> a;
        wil_s16<Real>(r.data(), z.data(), n, static_cast<Real>(Z[p]), static_cast<Real>(RT[p]), NR, NL, a);
        if (grid_out)
            for (int i = 1; i <= NR; ++i) {
                grid_out[(static_cast<size_t>(p) * NR + (i - 1)) * 2 + 0] = static_cast<double>(a.R[i]);
                grid_out[(static_cast<size_t>(p) * NR + (i - 1)) * 2 + 1] = static_cast<double>(a.V[i]);
            }
        wil_atom<Real>(a, ee.data(), NE, d.data(), rw.data());
        for (size_t k = 0; k < d.size(); ++k) {
            del[static_cast<size_t>(p) * NE * NL + k] = static_cast<double>(d[k]);
            if (raw) raw[static_cast<size_t>(p) * NE * NL + k] = static_cast<double>(rw[k]);
        }
    });
}

}  // namespace

extern "C" {

double orc_dlgkap(const double* pot, int jri, double E, int kappa, double xs, double dx, int ipt, double* rmax,
                  long long* steps, long long* evals, double* U_out, double* W_out) {
    RelGrid g;
    g.xs = xs;
    g.dx = dx;
    g.ipt = ipt;
    Counters c;
    const double r = dlgkap(pot, jri, E, kappa, g, rmax, &c, U_out, W_out);
    if (steps) *steps = c.milne_steps;
    if (evals) *evals = c.corr_evals;
    return r;
}

double orc_sbfit(double T, double E, int L, double R, int real32, int asbuilt) {
    return real32 ? static_cast<double>(sbfit<float>(T, E, L, R, asbuilt != 0)) : sbfit<double>(T, E, L, R, asbuilt != 0);
}

// Returns N (after the cap); arrays must hold nmax entries (N is clipped to nmax).
int orc_rel_energy_grid(double es, double de, double ue, int ncap, int asbuilt, double* e_l0, double* e_l, double* e_ev,
                        int nmax, int* n_header) {
    std::vector<double> a, b, c;
    int N = rel_energy_grid(es, de, ue, ncap, asbuilt != 0, a, b, c, n_header);
    if (N > nmax) N = nmax;
    for (int i = 0; i < N; ++i) {
        e_l0[i] = a[i];
        e_l[i] = b[i];
        e_ev[i] = c[i];
    }
    return N;
}

// pot[p*pot_stride + j] raw r*V values (sign as in mufftin.d; |.| is taken here like PHSH_REL does).
// e_l0/e_l: [NE] shared (e_stride=0) or per potential (e_stride=NE).  jf: [P][NE][lsm+1];
// jfk (optional): [P][NE][lsm+1][2] = (JFD for kappa=l, JFU for kappa=-l-1).  counters[3] = calls, steps, evals.
int orc_rel_batch(const double* pot, long pot_stride, const int* jri, int P, const double* e_l0, const double* e_l,
                  long e_stride, int NE, int lsm, double xs, double dx, int ipt, int real32, int asbuilt, double* jf,
                  double* jfk, long long* counters, int nthreads) {
    for (int p = 0; p < P; ++p)
        if (jri[p] < 7 || jri[p] > pot_stride) return -1;
    RelGrid g;
    g.xs = xs;
    g.dx = dx;
    g.ipt = ipt;
    if (real32)
        rel_batch_t<float>(pot, pot_stride, jri, P, e_l0, e_l, e_stride, NE, lsm, g, asbuilt != 0, jf, jfk, counters,
                           nthreads);
    else
        rel_batch_t<double>(pot, pot_stride, jri, P, e_l0, e_l, e_stride, NE, lsm, g, asbuilt != 0, jf, jfk, counters,
                            nthreads);
    return 0;
}

void orc_conphas(const double* phas, int NE, int ld, int lmax, double* out) { conphas_unwrap(phas, NE, ld, lmax, out); }

// V,RX: [P][stride] (Rydberg, bohr); e_ryd[NE]; phs [P][NE][NL]
int orc_cav_batch(const double* V, const double* RX, long stride, const int* ntab, const double* rmt, int P,
                  const double* e_ryd, int NE, int NL, int real32, int asbuilt, double* phs, int nthreads) {
    return real32 ? cav_batch_t<float>(V, RX, stride, ntab, rmt, P, e_ryd, NE, NL, asbuilt != 0, phs, nthreads)
                  : cav_batch_t<double>(V, RX, stride, ntab, rmt, P, e_ryd, NE, NL, asbuilt != 0, phs, nthreads);
}

int orc_cav_energy_grid(double emin, double emax, double estep, int real32, double* e_ha, int nmax) {
    int n = 0;
    if (real32) {
        std::vector<float> e;
        n = cav_energy_grid<float>(static_cast<float>(emin), static_cast<float>(emax), static_cast<float>(estep), e);
        for (int i = 0; i < std::min(n, nmax); ++i) e_ha[i] = e[i];
    } else {
        std::vector<double> e;
        n = cav_energy_grid<double>(emin, emax, estep, e);
        for (int i = 0; i < std::min(n, nmax); ++i) e_ha[i] = e[i];
    }
    return std::min(n, nmax);
}

// rs,zs: [P][stride] rows of (r, r*V) exactly as in the NFORM=1 file (incl. the r<0 terminator row if any);
// e[NE] in Rydberg; del/raw: [P][NE][NL]; grid_out (optional): [P][NR][2] = (R, V) of S16.
int orc_wil_batch(const double* rs, const double* zs, long stride, const int* nrows, const double* Z, const double* RT,
                  int P, const double* e, int NE, int NL, int NR, int real32, double* del, double* raw,
                  double* grid_out, int nthreads) {
    if (NL < 1 || NL > 15 || NR < 5) return -1;
    if (real32)
        wil_batch_t<float>(rs, zs, stride, nrows, Z, RT, P, e, NE, NL, NR, del, raw, grid_out, nthreads);
    else
        wil_batch_t<double>(rs, zs, stride, nrows, Z, RT, P, e, NE, NL, NR, del, raw, grid_out, nthreads);
    return 0;
}

int orc_wil_energy_grid(double e1, double e2, int ne, int real32, double* e) {
    if (real32) {
        std::vector<float> v;
        wil_energy_grid<float>(static_cast<float>(e1), static_cast<float>(e2), ne, v);
        for (int i = 0; i < ne; ++i) e[i] = v[i];
    } else {
        std::vector<double> v;
        wil_energy_grid<double>(e1, e2, ne, v);
        for (int i = 0; i < ne; ++i) e[i] = v[i];
    }
    return ne;
}

}  // extern "C"
