// phsh_conphas_files.cpp -- file layer of conphas (phsh3): split `phasout`, parse the sections, remove the pi jumps
// on the GPU and write the LEED-program formats.
//
// Replaces, byte for byte (paths relative to the reference checkout):
//   Conphas.split_phasout      phaseshifts/conphas.py:204-257
//   Conphas.load_data          phaseshifts/conphas.py:171-202
//   Conphas.calculate          phaseshifts/conphas.py:374-523   (unwrap :437-456 runs in conphas_kernel)
//   Conphas._write_viper_output phaseshifts/conphas.py:346-371
// including the quirks of the Python (energies taken from the first input file, extra rows appended to the
// first n_phases rows, the header's N capped at 250).
#include <cctype>
#include <charconv>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <ctime>
#include <fstream>
#include <new>
#include <stdexcept>
#include <string>
#include <vector>

#include <pwd.h>
#include <unistd.h>

#include "phsh_common.h"

namespace {

using phsh::set_error;
const double HARTREE = 27.21;  // conphas.py:66

bool read_all_lines(const char* path, std::vector<std::string>* lines) {
    std::ifstream f(path, std::ios::binary);
    if (!f) return false;
    std::string ln;
    while (std::getline(f, ln)) lines->push_back(ln + "\n");  // keep the newline like readlines()
    if (!lines->empty()) {
        // getline drops a final newline that may not exist; restore the file's own ending
        f.clear();
        f.seekg(0, std::ios::end);
        const std::streamoff n = f.tellg();
        if (n > 0) {
            f.seekg(n - 1);
            char c = 0;
            f.get(c);
            if (c != '\n') lines->back().pop_back();
        }
    }
    return true;
}

std::string dash_spaced(const std::string& s) {  // str.replace("-", " -")
    std::string o;
    for (char c : s) {
        if (c == '-') o.push_back(' ');
        o.push_back(c);
    }
    return o;
}

std::vector<std::string> split_ws(const std::string& s) {
    std::vector<std::string> t;
    size_t i = 0;
    while (i < s.size()) {
        while (i < s.size() && isspace(static_cast<unsigned char>(s[i]))) ++i;
        size_t j = i;
        while (j < s.size() && !isspace(static_cast<unsigned char>(s[j]))) ++j;
        if (j > i) t.push_back(s.substr(i, j - i));
        i = j;
    }
    return t;
}

bool to_double(const std::string& s, double* v) {
    // plain decimal tokens (the whole phasout) go through from_chars; anything else (leading '+', hex, inf/nan spellings,
    // out-of-range) keeps strtod's behaviour
    const char* b = s.c_str();
    const char* e = b + s.size();
    if (!s.empty() && s[0] != '+') {
        const auto r = std::from_chars(b, e, *v);
        if (r.ec == std::errc() && r.ptr == e) return true;
    }
    char* end = nullptr;
    *v = strtod(b, &end);
    return end != b && *end == '\0';
}

std::string basename_of(const std::string& p) {
    const size_t k = p.find_last_of('/');
    return k == std::string::npos ? p : p.substr(k + 1);
}
std::string dirname_of(const std::string& p) {
    const size_t k = p.find_last_of('/');
    if (k == std::string::npos) return "";
    return k == 0 ? "/" : p.substr(0, k);
}
std::string strip_ext(const std::string& b) {  // os.path.splitext(b)[0]
    const size_t k = b.find_last_of('.');
    if (k == std::string::npos || k == 0) return b;
    // a leading-dots-only prefix is not an extension
    bool alldots = true;
    for (size_t i = 0; i < k; ++i)
        if (b[i] != '.') alldots = false;
    return alldots ? b : b.substr(0, k);
}

std::string fmt(const char* f, double v) {
    char b[64];
    snprintf(b, sizeof b, f, v);
    return b;
}

// "%W.Pf" through std::to_chars (prints what printf prints -- the correctly rounded decimal -- several times faster)
inline void put_fixed(std::string* out, double v, int width, int prec) {
    char b[400];
    const auto r = std::to_chars(b, b + sizeof b - 1, v, std::chars_format::fixed, prec);
    int n;
    if (r.ec != std::errc()) {
        n = snprintf(b, sizeof b, "%*.*f", width, prec, v);
        out->append(b, n);
        return;
    }
    n = static_cast<int>(r.ptr - b);
    if (n < width) out->append(width - n, ' ');
    out->append(b, n);
}

struct Section {
    int n_phases = 0, lmf = 0;
    std::vector<std::vector<double>> phas;  // rows as the Python builds them (may be longer than lmf+1)
};

}  // namespace

extern "C" {

static int split_phasout_impl(const char* phasout, const char* const* names, int n_names, char* written,
                             int written_cap) {
    if (!phasout) {
        set_error("NULL path");
        return PHSH_B200_EINVAL;
    }
    std::vector<std::string> lines;
    if (!read_all_lines(phasout, &lines)) {
        set_error("Cannot open file '%s'", phasout);
        return PHSH_B200_EIO;
    }
    std::vector<size_t> idx;  // lines whose first non-blank character is a letter (conphas.py:224-230)
    for (size_t i = 0; i < lines.size(); ++i) {
        std::string t;
        for (char c : lines[i])
            if (c != ' ') t.push_back(c);
        if (!t.empty() && isalpha(static_cast<unsigned char>(t[0]))) idx.push_back(i);
    }
    std::string list;
    for (size_t s = 0; s < idx.size(); ++s) {
        std::string name;
        if (static_cast<int>(s) < n_names && names && names[s]) {
            name = names[s];
        } else {  // guessed: last blank-separated word of the title before any '#', + ".ph" (conphas.py:232-234)
            std::string t = lines[idx[s]].substr(0, lines[idx[s]].find('#'));
            const size_t k = t.find_last_of(' ');
            t = (k == std::string::npos) ? t : t.substr(k + 1);
            std::string u;
            for (char c : t)
                if (c != '\n' && c != '\r') u.push_back(c);
            name = u + ".ph";
        }
        FILE* f = fopen(name.c_str(), "wb");
        if (!f) {
            set_error("cannot write '%s'", name.c_str());
            return PHSH_B200_EIO;
        }
        const size_t end = (s + 1 < idx.size()) ? idx[s + 1] : lines.size();
        for (size_t i = idx[s]; i < end; ++i) fputs(lines[i].c_str(), f);
        fclose(f);
        list += name + "\n";
    }
    if (written && written_cap > 0) {
        strncpy(written, list.c_str(), static_cast<size_t>(written_cap) - 1);
        written[written_cap - 1] = '\0';
    }
    return static_cast<int>(idx.size());
}

// Conphas.load_data (conphas.py:171-202): the title line is skipped, the second line holds initial energy, step,
// n_phases and lmf, and everything after it is one stream of numbers in which a '-' also starts a new number.
static int parse_phase_file(const char* path, double* e0, double* de, int* n_phases, int* lmf, std::vector<double>* data) {
    std::vector<std::string> lines;
    if (!read_all_lines(path, &lines) || lines.size() < 2) {
        set_error("cannot read phase shift file '%s'", path);
        return PHSH_B200_EIO;
    }
    const std::vector<std::string> h = split_ws(dash_spaced(lines[1]));
    double dn, dl;
    if (h.size() < 4 || !to_double(h[0], e0) || !to_double(h[1], de) || !to_double(h[2], &dn) || !to_double(h[3], &dl) ||
        h[2].find_first_not_of("+-0123456789") != std::string::npos ||
        h[3].find_first_not_of("+-0123456789") != std::string::npos) {
        set_error("'%s': bad header line (initial energy, step, n_phases, lmf)", path);
        return PHSH_B200_EIO;
    }
    if (!(dn >= -2.0e9 && dn <= 2.0e9) || !(dl >= -1.0 && dl <= 1.0e4)) {
        set_error("'%s': header values out of range (n_phases=%g, lmf=%g)", path, dn, dl);
        return PHSH_B200_EIO;
    }
    *n_phases = static_cast<int>(dn);
    *lmf = static_cast<int>(dl);
    std::string joined;  // "".join(line.replace("-", " -").replace("\n", "")) -- no separator between lines
    for (size_t k = 2; k < lines.size(); ++k) {
        const std::string t = dash_spaced(lines[k]);
        for (char c : t)
            if (c != '\n') joined.push_back(c);
    }
    data->clear();
    for (const std::string& tok : split_ws(joined)) {
        double v;
        if (!to_double(tok, &v)) {
            set_error("'%s': could not convert string to float: '%s'", path, tok.c_str());
            return PHSH_B200_EIO;
        }
        data->push_back(v);
    }
    return 0;
}

static int load_phase_file_impl(const char* path, double* header, double* data, long data_cap, long* n_data) {
    if (!path || !header || !n_data || data_cap < 0 || (data_cap > 0 && !data)) {
        set_error("bad arguments");
        return PHSH_B200_EINVAL;
    }
    double e0 = 0, de = 0;
    int n_phases = 0, lmf = 0;
    std::vector<double> v;
    const int rc = parse_phase_file(path, &e0, &de, &n_phases, &lmf, &v);
    if (rc) return rc;
    header[0] = e0;
    header[1] = de;
    header[2] = n_phases;
    header[3] = lmf;
    *n_data = static_cast<long>(v.size());
    for (long k = 0; k < data_cap && k < *n_data; ++k) data[k] = v[k];
    return 0;
}

static int conphas_files_impl(const char* const* inputs, int n_inputs, const char* output_file, const char* format,
                              int lmax, const phsh_b200_conphas_opts* opts) {
    if (!inputs || n_inputs < 0 || !output_file || lmax < 0) {
        set_error("bad arguments");
        return PHSH_B200_EINVAL;
    }
    phsh::Trace trace("conphas_files");
    const int device = opts ? opts->device : 0;
    std::string fmtname = format ? format : "";
    for (char& c : fmtname) c = static_cast<char>(tolower(static_cast<unsigned char>(c)));
    std::vector<Section> secs(n_inputs);
    std::vector<std::vector<std::vector<double>>> conpha(n_inputs);
    std::vector<double> energy;
    int neng = 0, n_phases = 0;
    for (int i = 0; i < n_inputs; ++i) {
        Section& S = secs[i];
        std::vector<double> data;
        {
            double e0, de;
            const int prc = parse_phase_file(inputs[i], &e0, &de, &S.n_phases, &S.lmf, &data);
            if (prc) return prc;
        }
        n_phases = S.n_phases;
        if (n_phases > 250) n_phases = 250;  // conphas.py:412-413
        S.n_phases = n_phases;
        S.phas.clear();
        size_t index = 0;
        while (index < data.size()) {  // conphas.py:420-431, including the append-to-the-first-rows quirk
            if (n_phases <= 0) break;
            for (int ie = 0; ie < n_phases; ++ie) {
                if (index >= data.size()) {
                    set_error("Malformatted or incomplete input");
                    return PHSH_B200_EIO;
                }
                energy.push_back(data[index++]);
                S.phas.emplace_back();
                for (int l = 0; l <= S.lmf; ++l) {
                    if (index >= data.size()) {
                        set_error("Malformatted or incomplete input");
                        return PHSH_B200_EIO;
                    }
                    S.phas[ie].push_back(data[index++]);
                }
            }
        }
        neng += n_phases;
        if (n_phases > 0 && (static_cast<int>(S.phas.size()) < n_phases || lmax > S.lmf)) {
            set_error("'%s': lmax=%d exceeds the %d phases per energy of the file (IndexError in the reference)",
                      inputs[i], lmax, S.lmf + 1);
            return PHSH_B200_EINVAL;
        }
        // unwrap columns 0..min(lmax,18) on the GPU (conphas.py:437-456)
        const int L = S.lmf + 1;
        std::vector<double> in(static_cast<size_t>(n_phases) * L), out(in.size());
        for (int ie = 0; ie < n_phases; ++ie)
            for (int l = 0; l < L; ++l) in[static_cast<size_t>(ie) * L + l] = S.phas[ie][l];
        trace.mark("read + parse input");
        if (n_phases > 0) {
            const int rc = phsh_b200_conphas_host(in.data(), 1, n_phases, L, lmax > 18 ? 18 : lmax, out.data(), device);
            if (rc) return rc;
        }
        trace.mark("device (H2D, conphas_kernel, D2H)");
        conpha[i].assign(n_phases, std::vector<double>(L));
        for (int ie = 0; ie < n_phases; ++ie)
            for (int l = 0; l < L; ++l) conpha[i][ie][l] = out[static_cast<size_t>(ie) * L + l];
        // dataph_<name>.d next to the input (conphas.py:458-475)
        const std::string in_path = inputs[i];
        const std::string root = dirname_of(in_path);
        const std::string dataph = (root.empty() ? std::string() : (root == "/" ? root : root + "/")) + "dataph_" +
                                   strip_ext(basename_of(in_path)) + ".d";
        FILE* f = fopen(dataph.c_str(), "w");
        if (!f) {
            set_error("cannot write '%s'", dataph.c_str());
            return PHSH_B200_EIO;
        }
        {
            std::string txt;
            txt.reserve(static_cast<size_t>(lmax + 1) * (n_phases + 2) * 24);
            std::vector<std::string> ecol(n_phases);
            for (int ii = 0; ii < n_phases; ++ii) put_fixed(&ecol[ii], energy[ii] / HARTREE, 9, 7);
            for (int kk = 0; kk <= lmax; ++kk) {
                txt += "\"L = " + std::to_string(kk) + "\n";
                for (int ii = 0; ii < n_phases; ++ii) {
                    txt += ecol[ii];
                    txt.push_back('\t');
                    put_fixed(&txt, conpha[i][ii][kk], 9, 7);
                    txt.push_back('\n');
                }
                txt.push_back('\n');
            }
            fwrite(txt.data(), 1, txt.size(), f);
        }
        fclose(f);
        trace.mark("write dataph_<name>.d");
    }
    // conphas.py:434 re-runs `conpha = deepcopy(phas)` for every input file, which throws away the unwrapping of
    // all earlier files: only the last input is written unwrapped.  Reproduced unless asked otherwise.
    if (!(opts && opts->unwrap_all_inputs))
        for (int ii = 0; ii + 1 < n_inputs; ++ii)
            for (size_t ie = 0; ie < conpha[ii].size(); ++ie)
                for (size_t l = 0; l < conpha[ii][ie].size(); ++l) conpha[ii][ie][l] = secs[ii].phas[ie][l];
    for (int ii = 0; ii < n_inputs; ++ii)
        if (static_cast<int>(conpha[ii].size()) < n_phases) {
            set_error("'%s' holds fewer energies (%d) than the last input (%d) (IndexError in the reference)", inputs[ii],
                      static_cast<int>(conpha[ii].size()), n_phases);
            return PHSH_B200_EINVAL;
        }
    FILE* f = fopen(output_file, "w");
    if (!f) {
        set_error("cannot write '%s'", output_file);
        return PHSH_B200_EIO;
    }
    char tbuf[64];
    time_t now = time(nullptr);
    struct tm g;
    gmtime_r(&now, &g);
    if (fmtname == "viper" || fmtname == "viperleed") {
        const double d0[4] = {-10.17, -0.08, -74.19, 19.18};  // conphas.py:321
        const double* c = (opts && opts->has_v0) ? opts->v0 : d0;
        std::string ts;
        if (opts && opts->timestamp) {
            ts = opts->timestamp;
        } else {
            strftime(tbuf, sizeof tbuf, "%y%m%d-%H%M%S", &g);
            ts = tbuf;
        }
        fprintf(f, "%d %8.2f %8.2f %8.2f %8.2f phaseshifts %s\n", n_inputs, c[0], c[1], c[2], c[3], ts.c_str());
        for (int ie = 0; ie < n_phases; ++ie) {
            {
                std::string erow;
                put_fixed(&erow, energy[ie] / HARTREE, 7, 4);
                erow.push_back('\n');
                fwrite(erow.data(), 1, erow.size(), f);
            }
            for (int ii = 0; ii < n_inputs; ++ii)
                for (int l0 = 0; l0 <= lmax; l0 += 10) {
                    std::string row;
                    for (int l = l0; l <= lmax && l < l0 + 10; ++l) {
                        if (l > l0) row += " ";
                        put_fixed(&row, conpha[ii][ie][l], 7, 4);
                    }
                    while (!row.empty() && isspace(static_cast<unsigned char>(row.back()))) row.pop_back();
                    fprintf(f, "%s\n", row.c_str());
                }
        }
    } else if (fmtname == "curve") {
        fprintf(f, "# phase shift curve: %s\n#energy ", basename_of(output_file).c_str());
        for (int l = 0; l <= lmax; ++l) fprintf(f, "l=%d ", l);
        fputs("\n", f);
        for (int ie = 0; ie < n_phases; ++ie) {
            std::string row;
            put_fixed(&row, energy[ie] / HARTREE, 7, 4);
            row.push_back(' ');
            for (int ii = 0; ii < n_inputs; ++ii)
                for (int l = 0; l <= lmax; ++l) {
                    put_fixed(&row, conpha[ii][ie][l], 7, 4);
                    row.push_back(' ');
                }
            row.push_back('\n');
            fwrite(row.data(), 1, row.size(), f);
        }
    } else {
        if (fmtname == "cleed") {
            std::string user, ts;
            if (opts && opts->user) {
                user = opts->user;
            } else {
                const char* u = getenv("LOGNAME");
                if (!u) u = getenv("USER");
                if (!u) u = getenv("LNAME");
                if (!u) u = getenv("USERNAME");
                if (!u) {
                    struct passwd* pw = getpwuid(getuid());
                    u = pw ? pw->pw_name : "unknown";
                }
                user = u;
            }
            if (opts && opts->timestamp) {
                ts = opts->timestamp;
            } else {
                strftime(tbuf, sizeof tbuf, "%Y-%m-%d at %H:%M:%S", &g);
                ts = tbuf;
            }
            fprintf(f, "%d %d neng lmax (calculated by %s on %s)\n", neng, lmax, user.c_str(), ts.c_str());
        }
        std::string txt;
        txt.reserve(static_cast<size_t>(n_phases) * (8 + 7 * (lmax + 2) * std::max(n_inputs, 1)));
        for (int ie = 0; ie < n_phases; ++ie) {
            put_fixed(&txt, energy[ie] / HARTREE, 7, 4);
            txt.push_back('\n');
            for (int ii = 0; ii < n_inputs; ++ii)
                for (int l = 0; l <= lmax; ++l) put_fixed(&txt, conpha[ii][ie][l], 7, 4);
            txt.push_back('\n');
        }
        fwrite(txt.data(), 1, txt.size(), f);
    }
    fclose(f);
    trace.mark("write output file");
    return 0;
}

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

int phsh_b200_split_phasout(const char* phasout, const char* const* names, int n_names, char* written,
                            int written_cap) {
    PHSH_GUARDED(split_phasout_impl(phasout, names, n_names, written, written_cap))
}
int phsh_b200_load_phase_file(const char* path, double* header, double* data, long data_cap, long* n_data) {
    PHSH_GUARDED(load_phase_file_impl(path, header, data, data_cap, n_data))
}
int phsh_b200_conphas_files(const char* const* inputs, int n_inputs, const char* output_file, const char* format,
                            int lmax, const phsh_b200_conphas_opts* opts) {
    PHSH_GUARDED(conphas_files_impl(inputs, n_inputs, output_file, format, lmax, opts))
}

}  // extern "C"
