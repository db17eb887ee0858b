// phsh_files.cpp -- the four-file interface of libphsh on top of the batched CUDA entry points.
//
// Replaces the Fortran formatted I/O of (paths relative to the reference checkout)
//   PHSH_REL  phaseshifts/lib/libphsh.f:4572-4577, 4595-4611, 4635-4637, 4701-4718
//   PHSH_CAV  .phsh.orig/phsh2.f:117-172  (as-built: libphsh.f:3611-3691 writes the rows to DATAPH instead)
//   PHSH_WIL  .phsh.orig/phsh2.f:343-425  (libphsh.f:3929-4040)
// byte-for-byte where a downstream reader exists (`phasout`, parsed by phaseshifts/conphas.py:171-257).
// All numerics happen on the GPU through phsh_b200_*_batch_host; nothing is computed here.
#include <algorithm>
#include <charconv>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <map>
#include <new>
#include <stdexcept>
#include <sstream>
#include <string>
#include <vector>

#include "phsh_common.h"

namespace {

using phsh::set_error;

// ---------------------------------------------------------------- Fortran formatted input
std::string pad(const std::string& s, size_t n) {
    std::string t = s;
    while (!t.empty() && (t.back() == '\n' || t.back() == '\r')) t.pop_back();
    if (t.size() < n) t.append(n - t.size(), ' ');
    return t;
}

// Fw.d / Ew.d / Dw.d input editing of one field (BLANK=NULL): blanks ignored, D exponent, implied decimals
bool fread_field(const std::string& field, int d, double* out) {
    char s[96];
    size_t n = 0;
    bool has_point = false, in_exp = false;
    for (char c : field) {
        if (c == ' ' || c == '\t') continue;
        if (n + 2 >= sizeof s) return false;
        if (c == 'D' || c == 'd' || c == 'e') c = 'E';
        // exponent without a letter: 1.5+03
        if ((c == '+' || c == '-') && n > 0 && s[n - 1] != 'E' && !in_exp) {
            s[n++] = 'E';
            in_exp = true;
        }
        if (c == 'E') in_exp = true;
        if (c == '.' && !in_exp) has_point = true;
        s[n++] = c;
    }
    if (n == 0) {
        *out = 0.0;
        return true;
    }
    const char* b = s + (s[0] == '+' ? 1 : 0);  // from_chars takes no leading '+'
    double v = 0.0;
    const auto r = std::from_chars(b, s + n, v);
    if (r.ec == std::errc::result_out_of_range) {  // strtod's behaviour: +-HUGE_VAL or 0
        s[n] = '\0';
        char* end = nullptr;
        v = strtod(s, &end);
        if (end != s + n) return false;
    } else if (r.ec != std::errc() || r.ptr != s + n || b == s + n) {
        return false;
    }
    *out = has_point ? v : v * std::pow(10.0, -d);
    return true;
}

bool iread_field(const std::string& field, int* out) {
    std::string s;
    for (char c : field)
        if (c != ' ' && c != '\t') s.push_back(c);
    if (s.empty()) {
        *out = 0;
        return true;
    }
    char* end = nullptr;
    const long v = strtol(s.c_str(), &end, 10);
    if (end == s.c_str() || *end != '\0') return false;
    *out = static_cast<int>(v);
    return true;
}

// ---------------------------------------------------------------- Fortran formatted output
// shortest-path conversions: std::to_chars with an explicit precision prints exactly what printf's %.*f / %.*e
// print (correctly rounded decimal of the binary value) at a fraction of the cost
inline int fixed_chars(char* buf, size_t cap, double x, int d) {
    const auto r = std::to_chars(buf, buf + cap - 1, x, std::chars_format::fixed, d);
    if (r.ec != std::errc()) return snprintf(buf, cap, "%.*f", d, x);
    *r.ptr = '\0';
    return static_cast<int>(r.ptr - buf);
}
inline int sci_chars(char* buf, size_t cap, double x, int d) {
    const auto r = std::to_chars(buf, buf + cap - 1, x, std::chars_format::scientific, d);
    if (r.ec != std::errc()) return snprintf(buf, cap, "%.*e", d, x);
    *r.ptr = '\0';
    return static_cast<int>(r.ptr - buf);
}

// Fw.d as gfortran writes it: optional leading zero dropped when the field is too narrow, asterisks on overflow
std::string ffmt(double x, int w, int d) {
    char buf[512];
    const int n = fixed_chars(buf, sizeof buf, x, d);
    const char* p = buf;
    int len = n;
    char tmp[512];
    if (len > w) {
        if (buf[0] == '0' && buf[1] == '.') {
            p = buf + 1;
            --len;
        } else if (buf[0] == '-' && buf[1] == '0' && buf[2] == '.') {
            tmp[0] = '-';
            memcpy(tmp + 1, buf + 2, n - 1);
            p = tmp;
            --len;
        }
    }
    if (len > w) return std::string(w, '*');
    std::string s(w, ' ');
    memcpy(&s[w - len], p, len);
    return s;
}

// Dw.d / Ew.d: 0.ddddD+ee
std::string dfmt(double x, int w, int d, char letter) {
    std::string body;
    if (x == 0.0) {
        body = "0." + std::string(d, '0') + letter + "+00";
    } else {
        char buf[128];
        sci_chars(buf, sizeof buf, std::fabs(x), d - 1);
        const char* pe = strchr(buf, 'e');
        const int ex = atoi(pe + 1) + 1;
        body.reserve(w + 4);
        if (x < 0) body.push_back('-');
        body += "0.";
        body.push_back(buf[0]);
        if (d > 1) body.append(buf + 2, pe - (buf + 2));
        char eb[16];
        snprintf(eb, sizeof eb, "%c%c%02d", letter, ex >= 0 ? '+' : '-', std::abs(ex));
        body += eb;
    }
    if (static_cast<int>(body.size()) > w) {
        if (body.compare(0, 2, "0.") == 0)
            body.erase(0, 1);
        else if (body.compare(0, 3, "-0.") == 0)
            body.erase(1, 1);
    }
    if (static_cast<int>(body.size()) > w) return std::string(w, '*');
    return std::string(w - body.size(), ' ') + body;
}

// gfortran list-directed REAL(4) / REAL(8) field, leading separator blank included.  One conversion: the d
// significant digits of the E form are also the digits of the F form (same value, same number of significant
// digits, the exponent taken after rounding), so the F form is assembled from them.
std::string list_real(double x, int kind) {
    const int w = kind == 4 ? 16 : 25, d = kind == 4 ? 9 : 17, n = kind == 4 ? 4 : 5;
    char buf[128];
    if (x == 0.0) {
        snprintf(buf, sizeof buf, "%.*f", d - 1, 0.0);
        std::string b = buf;
        return " " + std::string(w - n - b.size(), ' ') + b + std::string(n, ' ');
    }
    sci_chars(buf, sizeof buf, x, d - 1);
    char* pe = strchr(buf, 'e');
    const int ex = atoi(pe + 1);
    const int k = ex + 1;
    if (k >= 0 && k <= d) {
        const bool neg = buf[0] == '-';
        const char* m = buf + (neg ? 1 : 0);  // "D.DDDD..."
        char digits[32];
        digits[0] = m[0];
        memcpy(digits + 1, m + 2, d - 1);
        std::string b;
        b.reserve(w);
        if (neg) b.push_back('-');
        if (k == 0)
            b.push_back('0');
        else
            b.append(digits, k);
        b.push_back('.');
        b.append(digits + k, d - k);
        if (static_cast<int>(b.size()) > w - n) b = b.substr(0, w - n);
        return " " + std::string(w - n - b.size(), ' ') + b + std::string(n, ' ');
    }
    std::string m(buf, pe - buf);
    char eb[16];
    snprintf(eb, sizeof eb, "E%c%0*d", ex >= 0 ? '+' : '-', n - 2, std::abs(ex));
    std::string b = m + eb;
    return " " + std::string(b.size() < static_cast<size_t>(w) ? w - b.size() : 0, ' ') + b;
}

bool read_lines(const char* path, std::vector<std::string>* lines) {
    std::ifstream f(path);
    if (!f) return false;
    std::string ln;
    while (std::getline(f, ln)) lines->push_back(ln);
    return true;
}

struct Out {
    FILE* f = nullptr;
    ~Out() {
        if (f) fclose(f);
    }
    bool open(const char* p) {
        f = fopen(p, "w");
        return f != nullptr;
    }
    void put(const std::string& s) { fputs(s.c_str(), f); }
};

const double kF32_13_6 = static_cast<double>(13.6f);

struct RelBlock {
    std::string name;
    double es, de, ue, vc, rmt;
    int lsm, nz, jri;
    std::vector<double> zp;
    // derived
    int n_header = 0, n = 0;
    double es_h = 0, de_h = 0;
    std::vector<double> e_l0, e_l, e_ev;
};

// energy bookkeeping of PHSH_REL (libphsh.f:4625-4667, 4674-4693)
void rel_energies(RelBlock& b, int cap, bool asbuilt) {
    double ES = b.es, DE = b.de, UE = b.ue;
    if (!(DE > 0.0)) {
        ES = -0.5;
        DE = 0.025;
        UE = 1.0;
    }
    const double nd = (UE - ES) / DE + 0.5;
    int N = (nd > -2.0e9 && nd < 2.0e9) ? static_cast<int>(nd) : (nd > 0 ? 2000000000 : 0);  // int() like the Fortran
    N = N + 1;
    b.n_header = N;
    b.es_h = ES;
    b.de_h = DE;
    ES = ES / kF32_13_6;
    DE = DE / kF32_13_6;
    if (cap > 0 && N > cap) N = cap;
    if (N < 0) N = 0;
    b.n = N;
    b.e_l0.resize(N);
    b.e_l.resize(N);
    b.e_ev.resize(N);
    double E = ES;
    for (int j = 0; j < N; ++j) {
        b.e_l0[j] = E;
        E = E * kF32_13_6;
        b.e_ev[j] = E;
        E = E / kF32_13_6;
        E = E + DE;
    }
    E = ES;
    for (int j = 0; j < N; ++j) {
        b.e_l[j] = E;
        if (!asbuilt) {
            E = E * kF32_13_6;
            E = E / kF32_13_6;
        }
        E = E + DE;
    }
}

phsh_b200_file_opts default_opts(const phsh_b200_file_opts* o) {
    phsh_b200_file_opts d;
    d.flags = 0;
    d.max_energies = 250;
    d.max_jri = 340;
    d.device = 0;
    return o ? *o : d;
}

}  // namespace

extern "C" {

static int phsh_b200_rel_file_impl(const char* mufftin, const char* phasout, const char* dataph, const char* inpdat,
                       const phsh_b200_file_opts* opts_in) {
    const phsh_b200_file_opts o = default_opts(opts_in);
    if (!mufftin || !phasout || !dataph || !inpdat) {
        set_error("NULL path");
        return PHSH_B200_EINVAL;
    }
    phsh::Trace trace("rel_file");
    std::vector<std::string> lines;
    if (!read_lines(mufftin, &lines)) {  // Fortran: open(status='old') failure aborts the process; we report
        set_error("cannot open mufftin file '%s'", mufftin);
        return PHSH_B200_EIO;
    }
    const bool asbuilt = (o.flags & PHSH_B200_ASBUILT) != 0;
    std::vector<RelBlock> blocks;
    std::string inp;  // inpdat echo
    inp += "1\n\n\n" + std::string(60, ' ') + "INPUT data\n\n\n";
    size_t i = 0;
    while (i + 2 < lines.size() + 0 && i < lines.size()) {
        RelBlock b;
        if (i + 2 >= lines.size()) break;  // read ... end=999
        b.name = pad(lines[i], 16).substr(0, 16);
        const std::string h = pad(lines[i + 1], 59);
        const std::string h2 = pad(lines[i + 2], 18);
        if (!fread_field(h.substr(0, 12), 4, &b.es) || !fread_field(h.substr(12, 12), 4, &b.de) ||
            !fread_field(h.substr(24, 12), 4, &b.ue) || !iread_field(h.substr(40, 3), &b.lsm) ||
            !fread_field(h.substr(47, 12), 4, &b.vc) || !iread_field(h2.substr(0, 4), &b.nz) ||
            !fread_field(h2.substr(4, 10), 6, &b.rmt) || !iread_field(h2.substr(14, 4), &b.jri))
            break;  // err=999: silent stop like the Fortran
        inp += dfmt(b.es, 12, 4, 'D') + dfmt(b.de, 12, 4, 'D') + dfmt(b.ue, 12, 4, 'D') + std::string(10, ' ');
        {
            char t[16];
            snprintf(t, sizeof t, "%3d\n", b.lsm);
            inp += t;
            snprintf(t, sizeof t, "%4d", b.nz);
            inp += t;
            inp += ffmt(b.rmt, 10, 6);
            snprintf(t, sizeof t, "%4d", b.jri);
            inp += t;
            inp += "  ";
            for (int k = 0; k < 6; ++k) inp += ffmt(0.0, 10, 6);
            inp += "\n";
        }
        if (b.jri <= 0 || (o.max_jri > 0 && b.jri > o.max_jri)) break;  // libphsh.f:4609 goto 999
        const int nrec = (b.jri + 4) / 5;
        if (i + 3 + nrec > lines.size()) break;
        bool ok = true;
        for (int r = 0; r < nrec && ok; ++r) {
            const std::string ln = pad(lines[i + 3 + r], 70);
            for (int k = 0; k < 5 && static_cast<int>(b.zp.size()) < b.jri; ++k) {
                double v;
                ok = fread_field(ln.substr(14 * k, 14), 6, &v);
                if (!ok) break;
                b.zp.push_back(v);
            }
        }
        if (!ok) break;
        for (size_t k = 0; k < b.zp.size(); k += 5) {
            for (size_t m = k; m < std::min(k + 5, b.zp.size()); ++m) inp += dfmt(b.zp[m], 14, 6, 'D');
            inp += "\n";
        }
        if (b.jri < 7 || b.lsm < 0 || b.lsm > 63) {
            set_error("block '%s': JRI=%d / LSM=%d not supported (need JRI>=7, 0<=LSM<=63)", b.name.c_str(), b.jri, b.lsm);
            return PHSH_B200_EINVAL;
        }
        {
            const double de = b.de > 0.0 ? b.de : 0.025, span = b.de > 0.0 ? b.ue - b.es : 1.5;
            if (!(o.max_energies > 0) && !(span / de < 1.0e7)) {
                set_error("block '%s': %g energies requested with the 250-energy cap lifted", b.name.c_str(), span / de);
                return PHSH_B200_EINVAL;
            }
        }
        rel_energies(b, o.max_energies, asbuilt);
        blocks.push_back(std::move(b));
        i += 3 + nrec;
    }
    inp += "\n\n" + std::string(56, ' ') + "end OF INPUT data\n";
    trace.mark("read + parse mufftin, inpdat echo");

    // batch blocks that share (N, LSM) and the energy grid into one launch
    std::vector<std::vector<double>> jf(blocks.size());
    std::vector<char> done(blocks.size(), 0);
    for (size_t a = 0; a < blocks.size(); ++a) {
        if (done[a]) continue;
        std::vector<size_t> grp;
        for (size_t c = a; c < blocks.size(); ++c)
            if (!done[c] && blocks[c].lsm == blocks[a].lsm && blocks[c].e_l0 == blocks[a].e_l0 &&
                blocks[c].e_l == blocks[a].e_l)
                grp.push_back(c);
        const int P = static_cast<int>(grp.size()), NE = blocks[a].n, L1 = blocks[a].lsm + 1;
        long stride = 0;
        for (size_t c : grp) stride = std::max<long>(stride, blocks[c].jri);
        stride = (stride + 1) & ~1L;
        std::vector<double> pot(static_cast<size_t>(P) * stride, 0.0), out(static_cast<size_t>(P) * NE * L1);
        std::vector<int> jri(P);
        for (int p = 0; p < P; ++p) {
            std::copy(blocks[grp[p]].zp.begin(), blocks[grp[p]].zp.end(), pot.begin() + static_cast<size_t>(p) * stride);
            jri[p] = blocks[grp[p]].jri;
        }
        if (NE > 0) {
            phsh_b200_rel_opts ro;
            ro.xs = -9.03065133;
            ro.dx = 3.125e-2;
            ro.lsm = blocks[a].lsm;
            ro.flags = o.flags & (PHSH_B200_ASBUILT | PHSH_B200_FAST | PHSH_B200_REAL32);
            const int rc = phsh_b200_rel_batch_host(pot.data(), stride, jri.data(), P, blocks[a].e_l0.data(),
                                                    blocks[a].e_l.data(), 0, NE, &ro, out.data(), nullptr, nullptr,
                                                    o.device);
            if (rc) return rc;
        }
        for (int p = 0; p < P; ++p) {
            jf[grp[p]].assign(out.begin() + static_cast<size_t>(p) * NE * L1,
                              out.begin() + static_cast<size_t>(p + 1) * NE * L1);
            done[grp[p]] = 1;
        }
    }

    trace.mark("device (H2D, kernels, D2H)");
    Out fp, fd, fi;
    if (!fp.open(phasout) || !fd.open(dataph) || !fi.open(inpdat)) {
        set_error("cannot open an output file for writing ('%s', '%s' or '%s')", phasout, dataph, inpdat);
        return PHSH_B200_EIO;
    }
    fi.put(inp);
    trace.mark("open outputs, write inpdat");
    for (size_t bi = 0; bi < blocks.size(); ++bi) {
        const RelBlock& b = blocks[bi];
        const int L1 = b.lsm + 1;
        char t[64];
        fp.put("RELATIVISTIC PHASE SHIFTS FOR " + b.name + "\n");
        snprintf(t, sizeof t, "%5d%5d\n", b.n_header, b.lsm);
        fp.put(ffmt(b.es_h, 10, 4) + ffmt(b.de_h, 9, 4) + t);
        for (int ie = 0; ie < b.n; ++ie) {
            // format (F9.4,8F8.5) with format reversion: every record restarts with an F9.4 field
            std::vector<double> items;
            items.push_back(b.e_ev[ie]);
            for (int l = 0; l < L1; ++l) items.push_back(jf[bi][static_cast<size_t>(ie) * L1 + l]);
            for (size_t k = 0; k < items.size(); k += 9) {
                std::string rec = ffmt(items[k], 9, 4);
                for (size_t m = k + 1; m < std::min(k + 9, items.size()); ++m) rec += ffmt(items[m], 8, 5);
                fp.put(rec + "\n");
            }
        }
        std::vector<std::string> etxt(b.n);  // the energy column repeats for every l
        for (int ie = 0; ie < b.n; ++ie) etxt[ie] = list_real(b.e_ev[ie], 8);
        for (int kk = 0; kk < 8; ++kk) {  // nl=8 plotted phase shifts (libphsh.f:4707-4716)
            snprintf(t, sizeof t, "\"L=%2d\n", kk);
            fd.put(t);
            for (int ie = 0; ie < b.n; ++ie) {
                const double v = kk < L1 ? jf[bi][static_cast<size_t>(ie) * L1 + kk] : 0.0;
                fd.put(etxt[ie] + list_real(static_cast<double>(static_cast<float>(v)), 4) + "\n");
            }
            fd.put("\n");
        }
    }
    trace.mark("format phasout + dataph");
    return 0;
}

static int phsh_b200_cav_file_impl(const char* mufftin, const char* phasout, const char* dataph, const char* zph,
                       const phsh_b200_file_opts* opts_in) {
    const phsh_b200_file_opts o = default_opts(opts_in);
    if (!mufftin || !phasout || !dataph || !zph) {
        set_error("NULL path");
        return PHSH_B200_EINVAL;
    }
    std::vector<std::string> lines;
    if (!read_lines(mufftin, &lines)) {
        set_error("cannot open mufftin file '%s'", mufftin);
        return PHSH_B200_EIO;
    }
    const bool asbuilt = (o.flags & PHSH_B200_ASBUILT) != 0;
    struct Atom {
        std::string name16;
        double z, rmt, mtz;
        int ntab;
        std::vector<double> rx, v;
    };
    std::vector<Atom> atoms;
    size_t i = 0;
    int NRat = 0;
    if (lines.empty() || !iread_field(pad(lines[0], 4).substr(0, 4), &NRat)) {
        set_error("'%s': cannot read the atom count (I4)", mufftin);
        return PHSH_B200_EIO;
    }
    i = 1;
    for (int k = 0; k < NRat; ++k) {
        Atom a;
        if (i + 2 >= lines.size()) {
            set_error("'%s': unexpected end of file in atom %d", mufftin, k + 1);
            return PHSH_B200_EIO;
        }
        a.name16 = pad(lines[i], 16).substr(0, 16);
        const std::string h = pad(lines[i + 1], 24);
        if (!fread_field(h.substr(0, 8), 4, &a.z) || !fread_field(h.substr(8, 8), 4, &a.rmt) ||
            !fread_field(h.substr(16, 8), 4, &a.mtz) || !iread_field(pad(lines[i + 2], 4).substr(0, 4), &a.ntab)) {
            set_error("'%s': bad header of atom %d", mufftin, k + 1);
            return PHSH_B200_EIO;
        }
        if (a.ntab < 5 || a.ntab > 250 || i + 3 + a.ntab > lines.size()) {
            set_error("'%s': NTAB=%d of atom %d outside [5,250] or table truncated", mufftin, a.ntab, k + 1);
            return PHSH_B200_EIO;
        }
        for (int r = 0; r < a.ntab; ++r) {
            const std::string ln = pad(lines[i + 3 + r], 28);
            double rr, vv;
            if (!fread_field(ln.substr(0, 14), 5, &rr) || !fread_field(ln.substr(14, 14), 5, &vv)) {
                set_error("'%s': bad table row %d of atom %d", mufftin, r + 1, k + 1);
                return PHSH_B200_EIO;
            }
            // the Fortran holds these in REAL*4
            a.rx.push_back(static_cast<double>(static_cast<float>(rr)));
            a.v.push_back(static_cast<double>(static_cast<float>(vv)));
        }
        a.rmt = static_cast<double>(static_cast<float>(a.rmt));
        atoms.push_back(std::move(a));
        i += 3 + atoms.back().ntab;
    }
    // fixed energy grid of PHSH_CAV (phsh2.f:118-122, 136-150), REAL*4 arithmetic
    const float emin = 1.f, emax = 12.f, estep = .25f;
    const int ianz = static_cast<int>((emax - emin) / estep + 1.01f);
    const int NL = 12;
    std::vector<float> e_ha;
    {
        float E = emin;
        do {
            E = 2.f * E;
            E = 0.5f * E;
            e_ha.push_back(E);
            E = E + estep;
        } while (E <= emax);
    }
    const int NE = static_cast<int>(e_ha.size());
    std::vector<double> e_ryd(NE);
    for (int k = 0; k < NE; ++k) e_ryd[k] = static_cast<double>(2.f * e_ha[k]);
    const int P = static_cast<int>(atoms.size());
    const long stride = 250;
    std::vector<double> V(static_cast<size_t>(P) * stride, 0.0), RX(static_cast<size_t>(P) * stride, 0.0), rmt(P);
    std::vector<int> ntab(P);
    for (int p = 0; p < P; ++p) {
        std::copy(atoms[p].v.begin(), atoms[p].v.end(), V.begin() + static_cast<size_t>(p) * stride);
        std::copy(atoms[p].rx.begin(), atoms[p].rx.end(), RX.begin() + static_cast<size_t>(p) * stride);
        rmt[p] = atoms[p].rmt;
        ntab[p] = atoms[p].ntab;
    }
    std::vector<double> phs(static_cast<size_t>(P) * NE * NL);
    if (P > 0) {
        const int rc = phsh_b200_cav_batch_host(V.data(), RX.data(), stride, ntab.data(), rmt.data(), P, e_ryd.data(), NE,
                                                NL, o.flags & PHSH_B200_ASBUILT, phs.data(), o.device);
        if (rc) return rc;
    }
    Out fp, fd, fz;
    if (!fp.open(phasout) || !fd.open(dataph) || !fz.open(zph)) {
        set_error("cannot open an output file for writing");
        return PHSH_B200_EIO;
    }
    Out& rows = asbuilt ? fd : fp;  // libphsh.f:3634-3646 writes title/grid/rows to the DATAPH unit
    fd.put("TitleText: DELTA(E)\n");
    char t[128];
    const int nl_header = (o.flags & 0x20) ? NL - 1 : NL;  // PHSH_B200_CONPHAS_HEADER
    for (int p = 0; p < P; ++p) {
        const Atom& a = atoms[p];
        snprintf(t, sizeof t, "PHASE SHIFTS FOR %s  ATOMIC NUMBER,%s\n", a.name16.c_str(), ffmt(a.z, 6, 1).c_str());
        fz.put(t);
        fz.put("MUFFIN-TIN RADIUS" + ffmt(a.rmt, 8, 4) + "       MUFFIN-TIN ZERO LEVEL" + ffmt(a.mtz / 2.0, 8, 4) +
               " HARTREES\n");
        // NAME is REAL*8(2) read with 2A8 and written with 2A4: the first four characters of each half
        rows.put("NON-RELATIVISTIC PHASE SHIFTS FOR " + a.name16.substr(0, 4) + a.name16.substr(8, 4) + "\n");
        snprintf(t, sizeof t, "  %3d  %3d\n", ianz, nl_header);
        rows.put(ffmt(emin, 9, 4) + ffmt(estep, 9, 4) + t);
        for (int ie = 0; ie < NE; ++ie) {
            fz.put("ENERGY" + ffmt(e_ha[ie], 8, 4) + " HARTREES\n");
            std::string zl;
            for (int l = 0; l < NL; ++l) {
                zl += ffmt(phs[(static_cast<size_t>(p) * NE + ie) * NL + l], 12, 5);
                if (l % 10 == 9 || l == NL - 1) zl += "\n";
            }
            fz.put(zl);
            std::vector<double> items;
            items.push_back(static_cast<double>(e_ha[ie] * 27.21f));
            for (int l = 0; l < NL; ++l) items.push_back(phs[(static_cast<size_t>(p) * NE + ie) * NL + l]);
            for (size_t k = 0; k < items.size(); k += 9) {
                std::string rec = ffmt(items[k], 9, 4);
                for (size_t m = k + 1; m < std::min(k + 9, items.size()); ++m) rec += ffmt(items[m], 8, 4);
                rows.put(rec + "\n");
            }
        }
        for (int kk = 0; kk < NL; ++kk) {
            snprintf(t, sizeof t, "\"L=%2d\n", kk);
            fd.put(t);
            for (int ie = 0; ie < NE; ++ie)
                fd.put(list_real(e_ha[ie], 8) + list_real(phs[(static_cast<size_t>(p) * NE + ie) * NL + kk], 8) + "\n");
            fd.put("\n");
        }
    }
    return 0;
}

// tiny namelist reader: KEY=value pairs between '&NAME' and '&end' / '/'
static bool parse_namelist(const std::string& text, const char* name, std::map<std::string, double>* kv) {
    std::string up;
    for (char c : text) up.push_back(static_cast<char>(toupper(static_cast<unsigned char>(c))));
    const std::string tag = std::string("&") + name;
    const size_t a = up.find(tag);
    if (a == std::string::npos) return false;
    size_t b = up.find("&END", a + tag.size());
    const size_t b2 = up.find('/', a + tag.size());
    if (b == std::string::npos || (b2 != std::string::npos && b2 < b)) b = b2;
    if (b == std::string::npos) b = up.size();
    std::string body = up.substr(a + tag.size(), b - a - tag.size());
    for (char& c : body)
        if (c == ',') c = ' ';
    std::istringstream ss(body);
    std::string tok;
    while (ss >> tok) {
        const size_t eq = tok.find('=');
        if (eq == std::string::npos) continue;
        std::string key = tok.substr(0, eq), val = tok.substr(eq + 1);
        if (val.empty() && !(ss >> val)) break;
        for (char& c : val)
            if (c == 'D') c = 'E';
        (*kv)[key] = atof(val.c_str());
    }
    return true;
}

static int phsh_b200_wil_file_impl(const char* mufftin, const char* phasout, const char* dataph, const char* zph,
                       const phsh_b200_file_opts* opts_in) {
    const phsh_b200_file_opts o = default_opts(opts_in);
    if (!mufftin || !phasout || !dataph || !zph) {
        set_error("NULL path");
        return PHSH_B200_EINVAL;
    }
    std::vector<std::string> lines;
    if (!read_lines(mufftin, &lines)) {
        set_error("cannot open mufftin file '%s'", mufftin);
        return PHSH_B200_EIO;
    }
    size_t i = 0;
    std::map<std::string, double> nl2;
    while (i < lines.size() && !parse_namelist(lines[i], "NL2", &nl2)) ++i;
    if (i >= lines.size() || !nl2.count("NRR")) {
        set_error("'%s': namelist &NL2 NRR=... not found", mufftin);
        return PHSH_B200_EIO;
    }
    ++i;
    const int NRR = static_cast<int>(nl2["NRR"]);
    struct Atom {
        double z, rt;
        std::vector<double> rs, zs;
    };
    std::vector<Atom> atoms;
    double E1 = 4., E2 = 24.;
    int NE = 30, NL = 9, NR = 101, NEUO = 2;
    for (int k = 0; k < NRR; ++k) {
        std::map<std::string, double> nl16;
        while (i < lines.size() && !parse_namelist(lines[i], "NL16", &nl16)) ++i;
        if (i >= lines.size()) {
            set_error("'%s': namelist &NL16 of atom %d not found", mufftin, k + 1);
            return PHSH_B200_EIO;
        }
        ++i;
        Atom a;
        a.z = nl16.count("Z") ? nl16["Z"] : 0.0;
        a.rt = nl16.count("RT") ? nl16["RT"] : 0.0;
        if (nl16.count("E1")) E1 = nl16["E1"];
        if (nl16.count("E2")) E2 = nl16["E2"];
        if (nl16.count("NE")) NE = static_cast<int>(nl16["NE"]);
        if (nl16.count("NL")) NL = static_cast<int>(nl16["NL"]);
        if (nl16.count("NR")) NR = static_cast<int>(nl16["NR"]);
        if (nl16.count("NEUO")) NEUO = static_cast<int>(nl16["NEUO"]);
        if ((nl16.count("NEUI") && static_cast<int>(nl16["NEUI"]) != 1) ||
            (nl16.count("POTYP") && static_cast<int>(nl16["POTYP"]) != 2) || (nl16.count("IX") && nl16["IX"] < 1)) {
            set_error("'%s': only NEUI=1, POTYP=2, IX>=1 (what CAVPOT writes) are supported", mufftin);
            return PHSH_B200_EINVAL;
        }
        for (int r = 0; r < 200 && i < lines.size(); ++r, ++i) {  // phsh2.f:488-493
            const std::string ln = pad(lines[i], 28);
            double rr, zz;
            if (!fread_field(ln.substr(0, 14), 5, &rr) || !fread_field(ln.substr(14, 14), 5, &zz)) {
                set_error("'%s': bad table row in atom %d", mufftin, k + 1);
                return PHSH_B200_EIO;
            }
            a.rs.push_back(static_cast<double>(static_cast<float>(rr)));
            a.zs.push_back(static_cast<double>(static_cast<float>(zz)));
            if (rr < 0) {
                ++i;
                break;
            }
        }
        a.z = static_cast<double>(static_cast<float>(a.z));
        a.rt = static_cast<double>(static_cast<float>(a.rt));
        atoms.push_back(std::move(a));
    }
    if (NE < 1 || NL < 1 || NL > 15 || NR < 5 || NR > 201 || (o.max_energies > 0 && NE > 401)) {
        set_error("'%s': NE=%d NL=%d NR=%d outside the reference's limits (401, 15, 201)", mufftin, NE, NL, NR);
        return PHSH_B200_EINVAL;
    }
    // E(I) = E1 + FLOAT(I-1)*DE in REAL*4 (phsh2.f:351-353)
    std::vector<float> ef(NE);
    {
        const float DE = (static_cast<float>(E2) - static_cast<float>(E1)) / static_cast<float>(std::max(NE - 1, 1));
        for (int k = 1; k <= NE; ++k) ef[k - 1] = static_cast<float>(E1) + static_cast<float>(k - 1) * DE;
    }
    std::vector<double> e(NE);
    for (int k = 0; k < NE; ++k) e[k] = ef[k];
    const int P = static_cast<int>(atoms.size());
    const long stride = 202;
    std::vector<double> rs(static_cast<size_t>(P) * stride, 0.0), zs(static_cast<size_t>(P) * stride, 0.0), Z(P), RT(P);
    std::vector<int> nrows(P);
    for (int p = 0; p < P; ++p) {
        std::copy(atoms[p].rs.begin(), atoms[p].rs.end(), rs.begin() + static_cast<size_t>(p) * stride);
        std::copy(atoms[p].zs.begin(), atoms[p].zs.end(), zs.begin() + static_cast<size_t>(p) * stride);
        nrows[p] = static_cast<int>(atoms[p].rs.size());
        Z[p] = atoms[p].z;
        RT[p] = atoms[p].rt;
    }
    std::vector<double> del(static_cast<size_t>(P) * NE * NL);
    if (P > 0) {
        const int rc = phsh_b200_wil_batch_host(rs.data(), zs.data(), stride, nrows.data(), Z.data(), RT.data(), P,
                                                e.data(), NE, NL, NR, 0, del.data(), o.device);
        if (rc) return rc;
    }
    Out fp, fd, fz;
    if (!fp.open(phasout) || !fd.open(dataph) || !fz.open(zph)) {
        set_error("cannot open an output file for writing");
        return PHSH_B200_EIO;
    }
    char t[64];
    snprintf(t, sizeof t, "&NL2\n IP=1,\n NRR=%d,\n /\n", NRR);
    fz.put(t);
    std::vector<float> eo(ef);
    if (NEUO == 2)
        for (auto& x : eo) x = 0.5f * x;  // phsh2.f:391
    fd.put("TitleText: DELTA(E)\n");
    for (int p = 0; p < P; ++p)
        for (int kk = 0; kk < NL; ++kk) {
            snprintf(t, sizeof t, "\"L=%2d\n", kk);
            fd.put(t);
            for (int ie = 0; ie < NE; ++ie)
                fd.put(list_real(eo[ie], 4) +
                       list_real(static_cast<double>(static_cast<float>(del[(static_cast<size_t>(p) * NE + ie) * NL + kk])), 4) +
                       "\n");
            fd.put("\n");
        }
    fp.put(" BE CAREFUL ABOUT THE ORDER OF THE ELEMENTS\n");
    for (int ie = 0; ie < NE; ++ie) {
        fp.put(ffmt(eo[ie], 7, 4) + "\n");
        for (int p = 0; p < P; ++p)
            for (int l = 0; l < NL; l += 10) {
                std::string rec;
                for (int m = l; m < std::min(l + 10, NL); ++m)
                    rec += ffmt(del[(static_cast<size_t>(p) * NE + ie) * NL + m], 7, 4);
                fp.put(rec + "\n");
            }
    }
    return 0;
}

// exception barrier: nothing may propagate through the C ABI
#define PHSH_GUARDED(call)                                                \
    try {                                                                 \
        return call;                                                      \
    } catch (const std::bad_alloc&) {                                     \
        set_error("out of host memory while processing the input file");  \
        return PHSH_B200_ENOMEM;                                          \
    } catch (const std::exception& e) {                                   \
        set_error("unexpected error: %s", e.what());                      \
        return PHSH_B200_EINVAL;                                          \
    }

int phsh_b200_rel_file(const char* mufftin, const char* phasout, const char* dataph, const char* inpdat,
                       const phsh_b200_file_opts* opts) {
    PHSH_GUARDED(phsh_b200_rel_file_impl(mufftin, phasout, dataph, inpdat, opts))
}
int phsh_b200_cav_file(const char* mufftin, const char* phasout, const char* dataph, const char* zph,
                       const phsh_b200_file_opts* opts) {
    PHSH_GUARDED(phsh_b200_cav_file_impl(mufftin, phasout, dataph, zph, opts))
}
int phsh_b200_wil_file(const char* mufftin, const char* phasout, const char* dataph, const char* zph,
                       const phsh_b200_file_opts* opts) {
    PHSH_GUARDED(phsh_b200_wil_file_impl(mufftin, phasout, dataph, zph, opts))
}

#include "phsh_cavpot_files.inc"

}  // extern "C"
