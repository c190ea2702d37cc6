// Amber ff14SB parser and molecular topology builder (host, C++17, no GPU needed).
//
// Product-side replacement for the reference's pure-Python get_amber_parameters()
// (src/molearn/loss_functions/torch_protein_energy_utils.py:86-189, cards :192-356,
// read_lib_file :19-83) and the v=5 topology part of get_convolutions() (:420-504, 578-607,
// 659-745, 813-830).  The contract is bit-exact agreement of the index lists (values AND order)
// and of the fp64 parameter arrays with the reference, so the quirks of that code are kept:
// fixed-column slicing of the parameter cards, the "negative period continues the series"
// rule, charges keyed by Amber TYPE (the last atom of a type in a residue wins), the inverted
// vdW equivalence table, tracker insertion order driving the angle / torsion enumeration.
#include "topology.hpp"
#include "../../include/molearn_b200.h"

#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <map>
#include <set>
#include <sstream>
#include <stdexcept>
#include <unordered_map>

namespace {

struct KeyError : std::runtime_error { using std::runtime_error::runtime_error; };
struct ParseError : std::runtime_error { using std::runtime_error::runtime_error; };
struct IoError : std::runtime_error { using std::runtime_error::runtime_error; };

// ---- small Python-flavoured string helpers ------------------------------------------------
std::vector<std::string> split_ws(const std::string& s) {
    std::vector<std::string> out;
    size_t i = 0, n = s.size();
    while (i < n) {
        while (i < n && std::isspace(static_cast<unsigned char>(s[i]))) ++i;
        size_t j = i;
        while (j < n && !std::isspace(static_cast<unsigned char>(s[j]))) ++j;
        if (j > i) out.emplace_back(s.substr(i, j - i));
        i = j;
    }
    return out;
}
std::string slice(const std::string& s, size_t a, size_t b) {  // s[a:b]
    if (a >= s.size()) return std::string();
    return s.substr(a, std::min(b, s.size()) - a);
}
std::string strip(const std::string& s) {
    size_t a = 0, b = s.size();
    while (a < b && std::isspace(static_cast<unsigned char>(s[a]))) ++a;
    while (b > a && std::isspace(static_cast<unsigned char>(s[b - 1]))) --b;
    return s.substr(a, b - a);
}
bool blank(const std::string& s) { return strip(s).empty(); }
double to_double(const std::string& s) {
    char* end = nullptr;
    double v = std::strtod(s.c_str(), &end);
    if (end == s.c_str() || *end != '\0') throw ParseError("could not convert string to float: '" + s + "'");
    return v;
}
long to_int(const std::string& s) {
    char* end = nullptr;
    long v = std::strtol(s.c_str(), &end, 10);
    if (end == s.c_str() || *end != '\0') throw ParseError("invalid literal for int(): '" + s + "'");
    return v;
}
std::string unquote(const std::string& s) {
    if (s.size() < 2 || s.front() != '"' || s.back() != '"')
        throw ParseError("expected a quoted name but got " + s);
    return s.substr(1, s.size() - 2);
}

// line iterator that mirrors iterating a Python file object (lines keep their '\n')
struct Lines {
    std::vector<std::string> v;
    size_t pos = 0;
    explicit Lines(const std::string& path) {
        std::ifstream f(path, std::ios::binary);
        if (!f) throw IoError("ERROR: file " + path + " not found!");
        std::string all((std::istreambuf_iterator<char>(f)), std::istreambuf_iterator<char>());
        size_t a = 0;
        while (a < all.size()) {
            size_t b = all.find('\n', a);
            if (b == std::string::npos) { v.emplace_back(all.substr(a)); break; }
            v.emplace_back(all.substr(a, b - a + 1));
            a = b + 1;
        }
    }
    bool next(std::string& out) {
        if (pos >= v.size()) return false;
        out = v[pos++];
        return true;
    }
};

using Key2 = std::pair<std::string, std::string>;
struct Key3 { std::string a, b, c; bool operator<(const Key3& o) const { return std::tie(a, b, c) < std::tie(o.a, o.b, o.c); } };
struct Key4 { std::string a, b, c, d; bool operator<(const Key4& o) const { return std::tie(a, b, c, d) < std::tie(o.a, o.b, o.c, o.d); } };
struct Series { std::vector<double> factor, barrier, phase, period; };

struct ForceField {
    // amino12.lib
    std::map<std::string, std::map<std::string, std::string>> atom_type;            // [res][pdb name] -> type
    std::map<std::string, std::map<std::string, double>> charge;                    // [res][TYPE] -> charge
    std::map<std::string, std::map<std::string, std::vector<std::string>>> conn;    // [res][pdb name] -> names
    // parm10.dat + frcmod.ff14SB
    std::map<Key2, std::pair<double, double>> bond;    // (K, r0)
    std::map<Key3, std::pair<double, double>> angle;   // (K, theta0)
    std::map<Key4, Series> torsion;
    std::vector<std::pair<std::string, std::vector<std::string>>> equiv;  // insertion ordered
    std::map<std::string, std::pair<double, double>> vdw;                 // (R*, eps)
};

// ---- amino12.lib (utils.py:19-83) ------------------------------------------------------------
void parse_lib(const std::string& path, ForceField& ff) {
    Lines L(path);
    const auto& lines = L.v;
    std::map<std::string, std::map<long, std::string>> by_index;
    for (size_t k = 0; k < lines.size(); ++k) {
        auto tok = split_ws(lines[k]);
        if (tok.size() == 3 && tok[0] == "!!index" && tok[1] == "array" && tok[2] == "str") {
            for (size_t m = k + 1; m < lines.size(); ++m) {
                const std::string& l2 = lines[m];
                if (l2.empty() || l2[0] != ' ') break;
                auto t2 = split_ws(l2);
                if (t2.empty()) break;
                if (t2.size() != 1 && t2[0].size() != 5) break;
                std::string res = unquote(t2[0]);
                ff.atom_type[res]; ff.charge[res]; ff.conn[res]; by_index[res];
            }
            break;
        }
    }
    for (size_t k = 0; k < lines.size(); ++k) {
        const std::string& line = lines[k];
        if (slice(line, 0, 7) != "!entry.") continue;
        std::string res = slice(line, 7, 10);
        if (slice(line, 10, 22) == ".unit.atoms ") {
            if (!ff.atom_type.count(res)) throw KeyError("KeyError: '" + res + "'");
            for (size_t m = k + 1; m < lines.size(); ++m) {
                const std::string& l2 = lines[m];
                if (l2.empty() || l2[0] != ' ') break;
                auto t = split_ws(l2);
                if (t.size() != 8) throw ParseError("amino12.lib: expected 8 columns in atoms table: " + l2);
                std::string pdb = unquote(t[0]), amber = unquote(t[1]);
                ff.atom_type[res][pdb] = amber;
                ff.charge[res][amber] = to_double(t[7]);   // keyed by TYPE (utils.py:66)
                by_index[res][to_int(t[5])] = pdb;
                ff.conn[res][pdb] = {};
            }
        } else if (slice(line, 10, 29) == ".unit.connectivity ") {
            for (size_t m = k + 1; m < lines.size(); ++m) {
                const std::string& l2 = lines[m];
                if (l2.empty() || l2[0] != ' ') break;
                auto t = split_ws(l2);
                if (t.size() != 3) break;
                auto& idx = by_index[res];
                auto ia = idx.find(to_int(t[0])), ib = idx.find(to_int(t[1]));
                if (ia == idx.end() || ib == idx.end()) throw KeyError("KeyError: connectivity index in " + res);
                ff.conn[res][ia->second].push_back(ib->second);
                ff.conn[res][ib->second].push_back(ia->second);
            }
        }
    }
}

// ---- parameter cards (utils.py:192-356) ------------------------------------------------------
template <class F> void until_blank(Lines& L, F&& f) {
    std::string line;
    while (L.next(line)) {
        if (line == "\n" || blank(line)) return;
        f(line);
    }
}
void card_skip(Lines& L) { until_blank(L, [](const std::string&) {}); }
void card_bonds(Lines& L, ForceField& ff) {  // :217-233
    until_blank(L, [&](const std::string& line) {
        auto c = split_ws(slice(line, 5, 25));
        if (c.size() != 2) throw ParseError("Expected 2 floats but got " + slice(line, 6, 26));
        ff.bond[{strip(slice(line, 0, 2)), strip(slice(line, 3, 5))}] = {to_double(c[0]), to_double(c[1])};
    });
}
void card_angles(Lines& L, ForceField& ff) {  // :237-256
    until_blank(L, [&](const std::string& line) {
        auto c = split_ws(slice(line, 8, 28));
        if (c.size() != 2) throw ParseError("Expected 2 floats but got " + slice(line, 6, 26));
        ff.angle[Key3{strip(slice(line, 0, 2)), strip(slice(line, 3, 5)), strip(slice(line, 6, 8))}] =
            {to_double(c[0]), to_double(c[1])};
    });
}
void card_torsions(Lines& L, ForceField& ff) {  // :259-293
    until_blank(L, [&](const std::string& line) {
        Key4 key{strip(slice(line, 0, 2)), strip(slice(line, 3, 5)), strip(slice(line, 6, 8)), strip(slice(line, 9, 11))};
        auto c = split_ws(slice(line, 11, 55));
        if (c.size() != 4) throw ParseError("I wanted four values here?");
        double f = static_cast<double>(to_int(c[0])), b = to_double(c[1]), ph = to_double(c[2]), pe = to_double(c[3]);
        auto it = ff.torsion.find(key);
        if (it != ff.torsion.end()) {
            double last = it->second.period.back();
            if (last > 0) {          // previous series complete: this line starts a replacement
                it->second = Series{{f}, {b}, {ph}, {pe}};
            } else if (last < 0) {   // negative period: more terms follow
                it->second.factor.push_back(f); it->second.barrier.push_back(b);
                it->second.phase.push_back(ph); it->second.period.push_back(pe);
            }
        } else {
            ff.torsion[key] = Series{{f}, {b}, {ph}, {pe}};
        }
    });
}
void card_equiv(Lines& L, ForceField& ff) {  // :341-347
    until_blank(L, [&](const std::string& line) {
        auto c = split_ws(line);
        for (auto& e : ff.equiv)
            if (e.first == c[0]) { e.second = c; return; }
        ff.equiv.emplace_back(c[0], c);
    });
}
void card_vdw(Lines& L, ForceField& ff) {  // :350-356
    until_blank(L, [&](const std::string& line) {
        auto c = split_ws(line);
        if (c.size() < 3) throw ParseError("vdW card: expected symbol R eps: " + line);
        ff.vdw[c[0]] = {to_double(c[1]), to_double(c[2])};
    });
}

void parse_forcefield(const std::string& dir, ForceField& ff) {  // utils.py:86-189
    parse_lib(dir + "/amino12.lib", ff);
    std::string line;
    {
        Lines L(dir + "/parm10.dat");
        L.next(line);     // title
        card_skip(L);     // card 2 masses: parsed by the reference, unused by the energy
        L.next(line);     // card 3 hydrophilic symbols, one line
        card_bonds(L, ff);
        card_angles(L, ff);
        card_torsions(L, ff);
        card_skip(L);     // card 7 impropers: unused
        card_skip(L);     // card 8 H-bond 10-12: unused
        card_equiv(L, ff);
        while (L.next(line)) {
            auto t = split_ws(line);
            if (t.size() > 1 && t[1] == "RE") card_vdw(L, ff);
        }
    }
    {
        Lines L(dir + "/frcmod.ff14SB");
        L.next(line);
        while (L.next(line)) {  // :158-177 - every test is applied to the same header line
            std::string tag = slice(line, 0, 4);
            if (tag == "MASS") card_skip(L);
            if (tag == "BOND") card_bonds(L, ff);
            if (tag == "ANGL") card_angles(L, ff);
            if (tag == "DIHE") card_torsions(L, ff);
            if (tag == "IMPR") card_skip(L);
            if (tag == "HBON") card_skip(L);
            if (tag == "NONB") card_vdw(L, ff);
        }
    }
    const double deg = M_PI / 180.0;  // np.deg2rad multiplies by this fp64 constant (:180-184)
    for (auto& kv : ff.angle) kv.second.second *= deg;
    for (auto& kv : ff.torsion)
        for (auto& p : kv.second.phase) p *= deg;
}

// ---- name normalisation (utils.py:420-487) --------------------------------------------------
struct Ren3 { const char* a; const char* r; const char* b; };
struct Ren2 { const char* a; const char* b; };
const Ren3 kFixHPre[] = {{"HB1", "MET", "HB3"}, {"HG1", "MET", "HG3"}, {"HB1", "ASN", "HB3"}};
const Ren2 kFixHAny[] = {{"HN", "H"}, {"1HD2", "HD21"}, {"2HD2", "HD22"}, {"1HG2", "HG21"}, {"2HG2", "HG22"},
    {"3HG2", "HG23"}, {"3HG1", "HG13"}, {"1HG1", "HG11"}, {"2HG1", "HG12"}, {"1HD1", "HD11"}, {"2HD1", "HD12"},
    {"3HD1", "HD13"}, {"3HD2", "HD23"}, {"1HH1", "HH11"}, {"2HH1", "HH12"}, {"1HH2", "HH21"}, {"2HH2", "HH22"},
    {"1HE2", "HE21"}, {"2HE2", "HE22"}};
const Ren3 kFixHPost[] = {{"HG11", "ILE", "HG13"}, {"CD", "ILE", "CD1"}, {"HD1", "ILE", "HD11"}, {"HD2", "ILE", "HD12"},
    {"HD3", "ILE", "HD13"}, {"HB1", "PHE", "HB3"}, {"HB1", "GLU", "HB3"}, {"HG1", "GLU", "HG3"}, {"HB1", "LEU", "HB3"},
    {"HB1", "ARG", "HB3"}, {"HG1", "ARG", "HG3"}, {"HD1", "ARG", "HD3"}, {"HB1", "ASP", "HB3"}, {"HA1", "GLY", "HA3"},
    {"HB1", "LYS", "HB3"}, {"HG1", "LYS", "HG3"}, {"HD1", "LYS", "HD3"}, {"HE1", "LYS", "HE3"}, {"HB1", "TYR", "HB3"},
    {"HB1", "HIP", "HB3"}, {"HB1", "SER", "HB3"}, {"HG1", "SER", "HG"}, {"HB1", "PRO", "HB3"}, {"HG1", "PRO", "HG3"},
    {"HD1", "PRO", "HD3"}, {"HB1", "LEU", "HB3"}, {"HB1", "GLN", "HB3"}, {"HG1", "GLN", "HG3"}, {"HB1", "TRP", "HB3"}};

void normalise(std::vector<std::string>& names, std::vector<std::string>& res,
               const std::vector<int32_t>& resids, bool fix_h) {
    const size_t n = names.size();
    for (auto& a : names) if (a == "OXT") a = "O";
    for (auto& r : res) if (r == "HSD") r = "HID";
    for (auto& r : res) if (r == "HSE") r = "HIE";
    // grouped by resid VALUE over the whole array (np.unique, :425-433)
    std::map<int32_t, std::vector<size_t>> groups;
    for (size_t i = 0; i < n; ++i) groups[resids[i]].push_back(i);
    for (auto& g : groups) {
        bool all_his = true, hd1 = false, he2 = false;
        for (size_t i : g.second) all_his = all_his && res[i] == "HIS";
        if (!all_his) continue;
        for (size_t i : g.second) { hd1 = hd1 || names[i] == "HD1"; he2 = he2 || names[i] == "HE2"; }
        const char* nw = (hd1 && he2) ? "HIP" : hd1 ? "HID" : he2 ? "HIE" : nullptr;
        if (nw) for (size_t i : g.second) res[i] = nw;
    }
    for (auto& r : res) if (r == "HIS") r = "HIE";
    if (fix_h) {
        for (const auto& q : kFixHPre)
            for (size_t i = 0; i < n; ++i) if (names[i] == q.a && res[i] == q.r) names[i] = q.b;
        for (const auto& q : kFixHAny)
            for (size_t i = 0; i < n; ++i) if (names[i] == q.a) names[i] = q.b;
        for (const auto& q : kFixHPost)
            for (size_t i = 0; i < n; ++i) if (names[i] == q.a && res[i] == q.r) names[i] = q.b;
    }
}

template <class M, class K> const typename M::mapped_type& at_or_key_error(const M& m, const K& k, const std::string& what) {
    auto it = m.find(k);
    if (it == m.end()) throw KeyError("KeyError: " + what);
    return it->second;
}

void build(const std::string& dir, mlp_topology& T, bool fix_h) {
    ForceField ff;
    parse_forcefield(dir, ff);
    normalise(T.names, T.resnames, T.resids, fix_h);
    const int n = T.n_atoms;
    std::vector<std::string> pn(n);
    for (int i = 0; i < n; ++i) pn[i] = (T.names[i] == "H2" || T.names[i] == "H3") ? "H" : T.names[i];  // :490
    T.types.resize(n); T.charges.resize(n); T.vdw_R.resize(n); T.vdw_e.resize(n);
    std::unordered_map<std::string, std::string> inv_equiv;  // :498-502, later lines overwrite earlier ones
    for (const auto& e : ff.equiv)
        for (const auto& s : e.second) inv_equiv[s] = e.first;
    for (int i = 0; i < n; ++i) {
        const auto& rt = at_or_key_error(ff.atom_type, T.resnames[i], "'" + T.resnames[i] + "'");
        T.types[i] = at_or_key_error(rt, pn[i], "'" + pn[i] + "' (residue " + T.resnames[i] + ")");           // :489
        T.charges[i] = at_or_key_error(at_or_key_error(ff.charge, T.resnames[i], T.resnames[i]), T.types[i], T.types[i]);  // :493
        auto ie = inv_equiv.find(T.types[i]);
        const std::string& vt = ie == inv_equiv.end() ? T.types[i] : ie->second;
        const auto& rv = at_or_key_error(ff.vdw, vt, "'" + vt + "' (vdW)");
        T.vdw_R[i] = static_cast<float>(rv.first);   // torch.tensor(list of python floats) -> fp32
        T.vdw_e[i] = static_cast<float>(rv.second);
    }

    // bonds, v=5 (:578-607)
    std::vector<std::vector<int32_t>> tracker(n);
    std::vector<std::pair<const std::string*, int32_t>> current, previous;
    bool have_resid = false;
    int32_t cur_resid = 0;
    auto emit = [&](int32_t i2, int32_t i1) {
        tracker[i1].push_back(i2);
        tracker[i2].push_back(i1);
        T.bonds.push_back(i2);
        T.bonds.push_back(i1);
    };
    auto contains = [](const std::vector<std::string>& v, const std::string& s) {
        return std::find(v.begin(), v.end(), s) != v.end();
    };
    for (int32_t i1 = 0; i1 < n; ++i1) {
        const std::string& a1 = pn[i1];
        const auto& cres = at_or_key_error(ff.conn, T.resnames[i1], "'" + T.resnames[i1] + "'");
        auto c1 = cres.find(a1);
        if (c1 == cres.end()) throw KeyError("AssertionError: atom '" + a1 + "' not in template of " + T.resnames[i1]);
        if (!have_resid || T.resids[i1] != cur_resid) {
            // the reference starts from current_resid=-9999; a first residue numbered -9999 would
            // not roll, which only matters for the (empty) previous list
            previous.swap(current);
            current.clear();
            cur_resid = T.resids[i1];
            have_resid = true;
        }
        if (a1 == "N")
            for (const auto& p : previous)
                if (*p.first == "C") emit(p.second, i1);
        for (const auto& p : current) {
            auto c2 = cres.find(*p.first);
            if (contains(c1->second, *p.first) && c2 != cres.end() && contains(c2->second, a1)) emit(p.second, i1);
        }
        current.emplace_back(&pn[i1], i1);
    }

    // angles, torsions, 1-4 (:659-687)
    for (int32_t a1 = 0; a1 < n; ++a1)
        for (int32_t a2 : tracker[a1])
            for (int32_t a3 : tracker[a2]) {
                if (a3 > a1) { T.angles.push_back(a1); T.angles.push_back(a2); T.angles.push_back(a3); }
                if (a3 != a1)
                    for (int32_t a4 : tracker[a3])
                        if (a4 > a1 && a2 != a4) {
                            T.torsions.insert(T.torsions.end(), {a1, a2, a3, a4});
                            T.pairs14.push_back(a1); T.pairs14.push_back(a4);
                        }
            }

    // parameters (:705-745)
    const size_t nb = T.bonds.size() / 2, na = T.angles.size() / 3, nt = T.torsions.size() / 4;
    T.bond_para.resize(nb * 2);
    for (size_t k = 0; k < nb; ++k) {
        const std::string &t0 = T.types[T.bonds[2 * k]], &t1 = T.types[T.bonds[2 * k + 1]];
        auto it = ff.bond.find({t0, t1});
        if (it == ff.bond.end()) it = ff.bond.find({t1, t0});
        if (it == ff.bond.end()) throw KeyError("KeyError: bond ('" + t1 + "', '" + t0 + "')");
        T.bond_para[2 * k] = it->second.second;      // r0
        T.bond_para[2 * k + 1] = it->second.first;   // K
    }
    T.angle_para.resize(na * 2);
    for (size_t k = 0; k < na; ++k) {
        const std::string &t0 = T.types[T.angles[3 * k]], &t1 = T.types[T.angles[3 * k + 1]], &t2 = T.types[T.angles[3 * k + 2]];
        auto it = ff.angle.find(Key3{t0, t1, t2});
        if (it == ff.angle.end()) it = ff.angle.find(Key3{t2, t1, t0});
        if (it == ff.angle.end()) throw KeyError("KeyError: angle ('" + t2 + "', '" + t1 + "', '" + t0 + "')");
        T.angle_para[2 * k] = it->second.second;     // theta0 (rad)
        T.angle_para[2 * k + 1] = it->second.first;  // K
    }
    std::vector<const Series*> found(nt);
    size_t pmax = 0;
    for (size_t k = 0; k < nt; ++k) {
        const std::string &t0 = T.types[T.torsions[4 * k]], &t1 = T.types[T.torsions[4 * k + 1]],
                          &t2 = T.types[T.torsions[4 * k + 2]], &t3 = T.types[T.torsions[4 * k + 3]];
        const Key4 keys[4] = {{t0, t1, t2, t3}, {t3, t2, t1, t0}, {"X", t2, t1, "X"}, {"X", t1, t2, "X"}};
        const Series* s = nullptr;
        for (const auto& key : keys) {
            auto it = ff.torsion.find(key);
            if (it != ff.torsion.end()) { s = &it->second; break; }
        }
        if (!s) throw KeyError("KeyError: torsion ('" + t0 + "', '" + t1 + "', '" + t2 + "', '" + t3 + "') cannot be found");
        found[k] = s;
        pmax = std::max(pmax, s->barrier.size());
    }
    T.P = static_cast<int32_t>(pmax);
    T.torsion_para.assign(nt * 4 * pmax, 0.0);
    for (size_t k = 0; k < nt; ++k) {
        double* p = &T.torsion_para[k * 4 * pmax];
        for (size_t s = 0; s < pmax; ++s) p[s] = 1.0;  // factor defaults to 1 so barrier/factor is finite
        const Series& S = *found[k];
        for (size_t s = 0; s < S.barrier.size(); ++s) {
            p[0 * pmax + s] = S.factor[s];
            p[1 * pmax + s] = S.barrier[s];
            p[2 * pmax + s] = S.phase[s];
            p[3 * pmax + s] = std::fabs(S.period[s]);
        }
    }

    // exclusions: unique unordered 1-2, 1-3, 1-4 pairs (:818-830 with pair semantics)
    std::set<std::pair<int32_t, int32_t>> ex;
    auto add = [&](int32_t a, int32_t b) { if (a != b) ex.insert({std::min(a, b), std::max(a, b)}); };
    for (size_t k = 0; k < nb; ++k) add(T.bonds[2 * k], T.bonds[2 * k + 1]);
    for (size_t k = 0; k < na; ++k) add(T.angles[3 * k], T.angles[3 * k + 2]);
    for (size_t k = 0; k < nt; ++k) add(T.torsions[4 * k], T.torsions[4 * k + 3]);
    T.excl.reserve(ex.size() * 2);
    for (const auto& p : ex) { T.excl.push_back(p.first); T.excl.push_back(p.second); }
}

thread_local std::string g_error;

}  // namespace

void mlp_set_error(const std::string& msg) { g_error = msg; }

extern "C" {

const char* mlp_last_error(void) { return g_error.c_str(); }
int mlp_abi_version(void) { return MLP_ABI_VERSION; }

int mlp_topology_build(const char* param_dir, int32_t n_atoms, const char* const* atom_names,
                       const char* const* res_names, const int32_t* resids, uint32_t flags,
                       mlp_topology** out) {
    if (!param_dir || !out || n_atoms < 0 || (n_atoms > 0 && (!atom_names || !res_names || !resids))) {
        mlp_set_error("mlp_topology_build: bad argument");
        return MLP_E_ARG;
    }
    *out = nullptr;
    auto* T = new mlp_topology();
    try {
        T->n_atoms = n_atoms;
        T->names.resize(n_atoms); T->resnames.resize(n_atoms); T->resids.assign(resids, resids + n_atoms);
        for (int i = 0; i < n_atoms; ++i) { T->names[i] = atom_names[i]; T->resnames[i] = res_names[i]; }
        build(param_dir, *T, (flags & MLP_TOPO_FIX_H) != 0);
    } catch (const KeyError& e) { mlp_set_error(e.what()); delete T; return MLP_E_KEY;
    } catch (const ParseError& e) { mlp_set_error(e.what()); delete T; return MLP_E_PARSE;
    } catch (const IoError& e) { mlp_set_error(e.what()); delete T; return MLP_E_IO;
    } catch (const std::exception& e) { mlp_set_error(e.what()); delete T; return MLP_E_ARG; }
    *out = T;
    return MLP_OK;
}

int mlp_topology_counts(const mlp_topology* t, int64_t counts[8]) {
    if (!t || !counts) { mlp_set_error("mlp_topology_counts: null argument"); return MLP_E_ARG; }
    counts[0] = t->n_atoms; counts[1] = t->bonds.size() / 2; counts[2] = t->angles.size() / 3;
    counts[3] = t->torsions.size() / 4; counts[4] = t->P; counts[5] = t->excl.size() / 2;
    counts[6] = t->pairs14.size() / 2; counts[7] = 0;
    return MLP_OK;
}

int mlp_topology_export(const mlp_topology* t, int32_t* bonds, double* bond_para, int32_t* angles,
                        double* angle_para, int32_t* torsions, double* torsion_para, int32_t* excl) {
    if (!t) { mlp_set_error("mlp_topology_export: null topology"); return MLP_E_ARG; }
    auto cp = [](auto* dst, const auto& v) { if (dst && !v.empty()) std::memcpy(dst, v.data(), v.size() * sizeof(v[0])); };
    cp(bonds, t->bonds); cp(bond_para, t->bond_para); cp(angles, t->angles); cp(angle_para, t->angle_para);
    cp(torsions, t->torsions); cp(torsion_para, t->torsion_para); cp(excl, t->excl);
    return MLP_OK;
}

int mlp_topology_pairs14(const mlp_topology* t, int32_t* pairs14) {
    if (!t || !pairs14) { mlp_set_error("mlp_topology_pairs14: null argument"); return MLP_E_ARG; }
    if (!t->pairs14.empty()) std::memcpy(pairs14, t->pairs14.data(), t->pairs14.size() * sizeof(int32_t));
    return MLP_OK;
}

int mlp_topology_params14(const mlp_topology* t, float* vdw_14R, float* vdw_14e, float* q1q2_14) {
    if (!t) { mlp_set_error("mlp_topology_params14: null topology"); return MLP_E_ARG; }
    // utils.py:833-836, in the reference's fp32 tensor arithmetic: 0.5*(R_i+R_j), sqrt(e_i+e_j) (a SUM under the
    // root, as written there), q_i*q_j/permittivity with permittivity = 1
    const size_t n = t->pairs14.size() / 2;
    for (size_t k = 0; k < n; ++k) {
        const int i = t->pairs14[2 * k], j = t->pairs14[2 * k + 1];
        if (vdw_14R) vdw_14R[k] = 0.5f * (t->vdw_R[i] + t->vdw_R[j]);
        if (vdw_14e) vdw_14e[k] = std::sqrt(t->vdw_e[i] + t->vdw_e[j]);
        if (q1q2_14) q1q2_14[k] = static_cast<float>(t->charges[i]) * static_cast<float>(t->charges[j]);
    }
    return MLP_OK;
}

int mlp_topology_atoms(const mlp_topology* t, char* names, char* resnames, char* types, double* charges,
                       float* vdw_R, float* vdw_eps) {
    if (!t) { mlp_set_error("mlp_topology_atoms: null topology"); return MLP_E_ARG; }
    auto put = [&](char* dst, const std::vector<std::string>& v) {
        if (!dst) return;
        std::memset(dst, 0, static_cast<size_t>(t->n_atoms) * 8);
        for (int i = 0; i < t->n_atoms; ++i) std::strncpy(dst + 8 * i, v[i].c_str(), 7);
    };
    put(names, t->names); put(resnames, t->resnames); put(types, t->types);
    if (charges) std::memcpy(charges, t->charges.data(), t->charges.size() * sizeof(double));
    if (vdw_R) std::memcpy(vdw_R, t->vdw_R.data(), t->vdw_R.size() * sizeof(float));
    if (vdw_eps) std::memcpy(vdw_eps, t->vdw_e.data(), t->vdw_e.size() * sizeof(float));
    return MLP_OK;
}

void mlp_topology_destroy(mlp_topology* t) { delete t; }

}  // extern "C"
