// qxb200 -- the file seam (B3): run a `.qx` / `.jld2` / `.yml` triple with no host-language help.
//
//   QXContexts.execute(dsl_file, input_file, param_file, output_file; ...)   /root/reference/bin/qxrun.jl:83-87
//   parameter file schema                                                    /root/reference/src/outputs.jl:47-78
//   data file = one JLD2 dataset per data label                              /root/reference/src/compute_graph/tensor_cache.jl:90-106
//
// Everything here sits ABOVE the public C ABI: the graph is built, compiled and executed through the same
// qxb_graph_* / qxb_amplitudes entry points a Julia `ccall` binding uses, so the file path cannot drift
// from the in-process one.  Pure host code; the compute calls fail with QXB_ERR_CUDA without a GPU.
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <fstream>
#include <memory>
#include <sstream>

#include "../../include/qxb200.h"
#include "qxb_ir.h"
#include "qxb_jld2.h"

using namespace qxb;

// ===================================================================== YAML subset
// What YAML.jl writes for output_params_dict (outputs.jl:54-77): nested block mappings, block sequences of
// scalars (possibly at the indentation of their key), plain / single- / double-quoted scalars, `~`/`null`,
// flow sequences / flow mappings of scalars.  Anchors, tags, multi-line scalars and multi-document streams are rejected.
namespace {

struct YNode {
    enum Kind { SCALAR, MAP, SEQ } kind = SCALAR;
    std::string s;
    bool null = false, quoted = false;
    std::vector<std::pair<std::string, YNode>> map;
    std::vector<YNode> seq;
    const YNode* get(const std::string& k) const {
        if (kind != MAP) return nullptr;
        for (auto& kv : map) if (kv.first == k) return &kv.second;
        return nullptr;
    }
};

struct YLine { int indent; std::string text; int lineno; };

[[noreturn]] void yerr(const std::string& path, int lineno, const std::string& what) {
    throw Error(QXB_ERR_ARG, path + ":" + std::to_string(lineno) + ": " + what);
}

struct YParser {
    std::string path;
    std::vector<YLine> lines;

    static std::string strip_comment(const std::string& in) {
        char q = 0;
        for (size_t i = 0; i < in.size(); ++i) {
            char c = in[i];
            if (q) { if (c == q) q = 0; else if (q == '"' && c == '\\') ++i; }
            else if (c == '\'' || c == '"') q = c;
            else if (c == '#' && (i == 0 || in[i - 1] == ' ' || in[i - 1] == '\t')) return in.substr(0, i);
        }
        return in;
    }
    static std::string trim(const std::string& s) {
        size_t a = s.find_first_not_of(" \t\r"), b = s.find_last_not_of(" \t\r");
        return a == std::string::npos ? "" : s.substr(a, b - a + 1);
    }
    void load(const std::string& text) {
        std::istringstream is(text);
        std::string ln;
        int no = 0;
        bool seen_doc = false;
        while (std::getline(is, ln)) {
            ++no;
            std::string body = strip_comment(ln);
            std::string t = trim(body);
            if (t.empty()) continue;
            if (t == "---") { if (seen_doc || !lines.empty()) yerr(path, no, "multi-document streams are not supported"); seen_doc = true; continue; }
            if (t == "...") break;
            if (t[0] == '%') continue;
            size_t ind = body.find_first_not_of(' ');
            if (body[ind] == '\t') yerr(path, no, "tab used for indentation");
            lines.push_back({(int)ind, t, no});
        }
    }
    YNode scalar(const std::string& t, int no) {
        YNode n;
        if (t.empty() || t == "~" || t == "null" || t == "Null" || t == "NULL") { n.null = true; return n; }
        if (t[0] == '&' || t[0] == '*' || t[0] == '!' || t[0] == '|' || t[0] == '>')
            yerr(path, no, "anchors, tags and block scalars are not supported");
        if (t[0] == '\'' || t[0] == '"') {
            char q = t[0];
            std::string out;
            size_t i = 1;
            for (; i < t.size(); ++i) {
                if (q == '\'' && t[i] == '\'') { if (i + 1 < t.size() && t[i + 1] == '\'') { out += '\''; ++i; continue; } break; }
                if (q == '"' && t[i] == '"') break;
                if (q == '"' && t[i] == '\\' && i + 1 < t.size()) {
                    char e = t[++i];
                    out += e == 'n' ? '\n' : e == 't' ? '\t' : e == '0' ? '\0' : e;
                    continue;
                }
                out += t[i];
            }
            if (i + 1 != t.size()) yerr(path, no, "malformed quoted scalar");
            n.s = out; n.quoted = true;
            return n;
        }
        if (t[0] == '[') {
            if (t.back() != ']') yerr(path, no, "flow sequence must close on the same line");
            n.kind = YNode::SEQ;
            std::string cur;
            char q = 0;
            std::string inner = t.substr(1, t.size() - 2);
            for (size_t i = 0; i <= inner.size(); ++i) {
                char c = i < inner.size() ? inner[i] : ',';
                if (q) { cur += c; if (c == q) q = 0; continue; }
                if (c == '\'' || c == '"') { q = c; cur += c; continue; }
                if (c == '[' || c == '{') yerr(path, no, "nested flow collections are not supported");
                if (c == ',') { std::string item = trim(cur); if (!item.empty() || i < inner.size()) n.seq.push_back(scalar(item, no)); cur.clear(); continue; }
                cur += c;
            }
            return n;
        }
        if (t[0] == '{') {                                   // flow mapping of scalars: {a: 1, b: 'x'}
            if (t.back() != '}') yerr(path, no, "flow mapping must close on the same line");
            n.kind = YNode::MAP;
            std::string cur;
            char q = 0;
            std::string inner = t.substr(1, t.size() - 2);
            for (size_t i = 0; i <= inner.size(); ++i) {
                char c = i < inner.size() ? inner[i] : ',';
                if (q) { cur += c; if (c == q) q = 0; continue; }
                if (c == '\'' || c == '"') { q = c; cur += c; continue; }
                if (c == '[' || c == '{') yerr(path, no, "nested flow collections are not supported");
                if (c == ',') {
                    std::string item = trim(cur), k, v;
                    cur.clear();
                    if (item.empty()) { if (i < inner.size()) yerr(path, no, "empty entry in flow mapping"); continue; }
                    if (!split_key(item, k, v)) yerr(path, no, "expected 'key: value' in flow mapping");
                    if (n.get(k)) yerr(path, no, "duplicate key '" + k + "'");
                    n.map.push_back({k, scalar(v, no)});
                    continue;
                }
                cur += c;
            }
            return n;
        }
        n.s = t;
        return n;
    }
    // splits "key: value" at the first ": " / trailing ":" outside quotes; false if the line is no mapping entry
    static bool split_key(const std::string& t, std::string& key, std::string& val) {
        char q = 0;
        for (size_t i = 0; i < t.size(); ++i) {
            char c = t[i];
            if (q) { if (c == q) q = 0; continue; }
            if ((c == '\'' || c == '"') && i == 0) { q = c; continue; }
            if (c == ':' && (i + 1 == t.size() || t[i + 1] == ' ')) {
                key = trim(t.substr(0, i));
                if (key.size() >= 2 && (key[0] == '\'' || key[0] == '"') && key.back() == key[0]) key = key.substr(1, key.size() - 2);
                val = trim(t.substr(i + 1));
                return true;
            }
        }
        return false;
    }
    static bool is_item(const std::string& t) { return t == "-" || (t.size() > 1 && t[0] == '-' && t[1] == ' '); }

    YNode block(size_t& i, int indent) {
        YNode n;
        if (is_item(lines[i].text)) {
            n.kind = YNode::SEQ;
            while (i < lines.size() && lines[i].indent == indent && is_item(lines[i].text)) {
                std::string rest = trim(lines[i].text.substr(1));
                int no = lines[i].lineno;
                ++i;
                if (rest.empty()) {
                    if (i < lines.size() && lines[i].indent > indent) n.seq.push_back(block(i, lines[i].indent));
                    else { YNode z; z.null = true; n.seq.push_back(z); }
                } else {
                    std::string k, v;
                    if (rest[0] != '\'' && rest[0] != '"' && rest[0] != '[' && split_key(rest, k, v))
                        yerr(path, no, "mappings inside sequences are not supported");
                    n.seq.push_back(scalar(rest, no));
                }
            }
            if (i < lines.size() && lines[i].indent > indent) yerr(path, lines[i].lineno, "unexpected indentation");
            return n;
        }
        n.kind = YNode::MAP;
        while (i < lines.size() && lines[i].indent == indent && !is_item(lines[i].text)) {
            std::string k, v;
            int no = lines[i].lineno;
            if (!split_key(lines[i].text, k, v)) yerr(path, no, "expected 'key: value'");
            if (n.get(k)) yerr(path, no, "duplicate key '" + k + "'");
            ++i;
            if (!v.empty()) { n.map.push_back({k, scalar(v, no)}); continue; }
            if (i < lines.size() && (lines[i].indent > indent || (lines[i].indent == indent && is_item(lines[i].text))))
                n.map.push_back({k, block(i, lines[i].indent)});
            else { YNode z; z.null = true; n.map.push_back({k, z}); }
        }
        if (i < lines.size() && lines[i].indent > indent) yerr(path, lines[i].lineno, "unexpected indentation");
        return n;
    }
    YNode parse(const std::string& text) {
        load(text);
        if (lines.empty()) { YNode z; z.null = true; return z; }
        size_t i = 0;
        YNode root = block(i, lines[0].indent);
        if (i != lines.size()) yerr(path, lines[i].lineno, "unexpected dedent / content after the document");
        return root;
    }
};

std::string read_text(const std::string& path) {
    std::ifstream in(path, std::ios::binary);
    if (!in) throw Error(QXB_ERR_ARG, "cannot open '" + path + "'");
    std::ostringstream ss;
    ss << in.rdbuf();
    return ss.str();
}

int64_t to_int(const std::string& path, const YNode* n, const char* key) {
    if (!n || n->kind != YNode::SCALAR || n->null) throw Error(QXB_ERR_ARG, path + ": output.params." + key + " missing");
    char* end = nullptr;
    long long v = strtoll(n->s.c_str(), &end, 10);
    if (end == n->s.c_str() || *end) throw Error(QXB_ERR_ARG, path + ": output.params." + key + " is not an integer: '" + n->s + "'");
    return v;
}

struct Params {
    qxb_params p{};
    std::vector<std::string> bitstrings;
};

// splitmix64: the documented stream of the native Uniform method (the reference draws from Julia's
// MersenneTwister, simulation.jl:24-28, which no other runtime reproduces; List files carry the strings)
uint64_t splitmix64(uint64_t& s) {
    uint64_t z = (s += 0x9e3779b97f4a7c15ull);
    z = (z ^ (z >> 30)) * 0xbf58476d1ce4e5b9ull;
    z = (z ^ (z >> 27)) * 0x94d049bb133111ebull;
    return z ^ (z >> 31);
}

Params read_params(const std::string& path) {
    YParser yp;
    yp.path = path;
    YNode root = yp.parse(read_text(path));
    const YNode* out = root.get("output");
    if (!out || out->kind != YNode::MAP) throw Error(QXB_ERR_ARG, path + ": no 'output' mapping (outputs.jl:54-77)");
    const YNode* method = out->get("method");
    const YNode* pr = out->get("params");
    if (!method || method->kind != YNode::SCALAR || !pr || pr->kind != YNode::MAP)
        throw Error(QXB_ERR_ARG, path + ": output.method / output.params missing");
    Params r;
    qxb_params& p = r.p;
    p.M = 0.0001;
    const YNode* seed = pr->get("seed");
    if (seed && seed->kind == YNode::SCALAR && !seed->null) { p.has_seed = 1; p.seed = to_int(path, seed, "seed"); }
    if (method->s == "List") {
        p.method = QXB_METHOD_LIST;
        const YNode* bs = pr->get("bitstrings");
        if (!bs || bs->kind != YNode::SEQ) throw Error(QXB_ERR_ARG, path + ": List method needs output.params.bitstrings");
        for (const YNode& b : bs->seq) {
            if (b.kind != YNode::SCALAR || b.null) throw Error(QXB_ERR_ARG, path + ": bitstrings must be scalars");
            r.bitstrings.push_back(b.s);
        }
        p.num_qubits = r.bitstrings.empty() ? 0 : (int64_t)r.bitstrings[0].size();
        p.num_samples = pr->get("num_samples") ? to_int(path, pr->get("num_samples"), "num_samples") : (int64_t)r.bitstrings.size();
    } else if (method->s == "Uniform" || method->s == "Rejection") {
        p.method = method->s == "Uniform" ? QXB_METHOD_UNIFORM : QXB_METHOD_REJECTION;
        p.num_qubits = to_int(path, pr->get("num_qubits"), "num_qubits");
        p.num_samples = to_int(path, pr->get("num_samples"), "num_samples");
        if (p.num_qubits < 0 || p.num_samples < 0) throw Error(QXB_ERR_ARG, path + ": negative num_qubits / num_samples");
        if (const YNode* m = pr->get("M")) if (m->kind == YNode::SCALAR && !m->null) p.M = strtod(m->s.c_str(), nullptr);
        if (const YNode* fm = pr->get("fix_M")) p.fix_M = fm->s == "true" || fm->s == "True" || fm->s == "TRUE";
        if (p.method == QXB_METHOD_UNIFORM) {
            uint64_t s = p.has_seed ? (uint64_t)p.seed : 0x5851f42d4c957f2dull;
            for (int64_t i = 0; i < p.num_samples; ++i) {
                std::string b((size_t)p.num_qubits, '0');
                uint64_t word = 0;
                for (int64_t q = 0; q < p.num_qubits; ++q) {
                    if (q % 64 == 0) word = splitmix64(s);
                    b[(size_t)q] = '0' + (char)((word >> (q % 64)) & 1);
                }
                r.bitstrings.push_back(b);
            }
        }
    } else {
        throw Error(QXB_ERR_ARG, path + ": output method \"" + method->s + "\" not supported");   // outputs.jl:74
    }
    for (const std::string& b : r.bitstrings) {
        if ((int64_t)b.size() != p.num_qubits) throw Error(QXB_ERR_ARG, path + ": bitstrings of different lengths");
        for (char c : b)
            if (c != '0' && c != '1' && c != '+' && c != '-')
                throw Error(QXB_ERR_ARG, path + ": bitstring '" + b + "' has characters outside 0 1 + - (basics.md:55-63)");
    }
    p.n_bitstrings = (int64_t)r.bitstrings.size();
    return r;
}

template <typename F>
int guard(F&& f) {
    try {
        f();
        return QXB_OK;
    } catch (const Error& e) {
        set_last_error(e.what());
        return e.code;
    } catch (const std::exception& e) {
        set_last_error(e.what());
        return QXB_ERR_ARG;
    }
}

// forwards an error of a nested ABI call (its message is already the last error)
void ok(int rc) {
    if (rc != QXB_OK) throw Error(rc, qxb_last_error());
}

std::string stem_of(const std::string& p) {
    size_t slash = p.find_last_of('/');
    size_t dot = p.find_last_of('.');
    if (dot == std::string::npos || (slash != std::string::npos && dot < slash)) return p;
    return p.substr(0, dot);
}

double now() {
    return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count();
}

}  // namespace

struct qxb_jld2 {
    jld2::File file;
};

// ===================================================================== C ABI
extern "C" {

int qxb_jld2_open(const char* path, qxb_jld2** f) {
    return guard([&] {
        if (!path || !f) throw Error(QXB_ERR_ARG, "null argument");
        std::unique_ptr<qxb_jld2> h(new qxb_jld2());
        h->file = jld2::read_file(path);
        *f = h.release();
    });
}

void qxb_jld2_close(qxb_jld2* f) { delete f; }

int qxb_jld2_count(const qxb_jld2* f, int* n, int* checksum_failures) {
    return guard([&] {
        if (!f || !n) throw Error(QXB_ERR_ARG, "null argument");
        *n = (int)f->file.datasets.size();
        if (checksum_failures) *checksum_failures = f->file.checksum_failures;
    });
}

int qxb_jld2_info(const qxb_jld2* f, int i, const char** name, int* elem_kind, int* elem_size, int* rank, int64_t* dims) {
    return guard([&] {
        if (!f || i < 0 || i >= (int)f->file.datasets.size()) throw Error(QXB_ERR_ARG, "dataset index out of range");
        const jld2::Dataset& d = f->file.datasets[i];
        if (d.dims.size() > QXB_JLD2_MAX_RANK) throw Error(QXB_ERR_UNSUPP, "dataset '" + d.name + "' has more than 32 dimensions");
        if (name) *name = d.name.c_str();
        if (elem_kind) *elem_kind = d.kind;
        if (elem_size) *elem_size = d.elem_size;
        if (rank) *rank = (int)d.dims.size();
        if (dims) for (size_t k = 0; k < d.dims.size(); ++k) dims[k] = d.dims[k];
    });
}

int qxb_jld2_read(const qxb_jld2* f, int i, void* out, int as_c64) {
    return guard([&] {
        if (!f || !out || i < 0 || i >= (int)f->file.datasets.size()) throw Error(QXB_ERR_ARG, "bad read arguments");
        const jld2::Dataset& d = f->file.datasets[i];
        if (as_c64) {
            std::vector<std::complex<double>> v = jld2::as_c64(d);
            if (!v.empty()) memcpy(out, v.data(), v.size() * sizeof(v[0]));
        } else {
            if (d.kind == jld2::EK_OTHER) throw Error(QXB_ERR_UNSUPP, "dataset '" + d.name + "' has a datatype this reader does not decode");
            if (!d.raw.empty()) memcpy(out, d.raw.data(), d.raw.size());
        }
    });
}

int qxb_jld2_write(const char* path, int n, const char* const* names, const int* elem_kinds, const int* elem_sizes,
                   const int* ranks, const int64_t* const* dims, const void* const* data, int commit_types) {
    return guard([&] {
        if (!path || n < 0 || (n > 0 && (!names || !elem_kinds || !elem_sizes || !ranks || !dims || !data)))
            throw Error(QXB_ERR_ARG, "bad write arguments");
        std::vector<jld2::WriteArray> arrays((size_t)n);
        for (int i = 0; i < n; ++i) {
            jld2::WriteArray& a = arrays[i];
            a.name = names[i]; a.kind = elem_kinds[i]; a.elem_size = elem_sizes[i]; a.data = data[i];
            if (ranks[i] < 0 || ranks[i] > QXB_JLD2_MAX_RANK) throw Error(QXB_ERR_ARG, "bad rank");
            a.dims.assign(dims[i], dims[i] + ranks[i]);
            static const int want[] = {16, 8, 8, 4};
            if (a.kind >= 0 && a.kind <= jld2::EK_F32 && a.elem_size != want[a.kind]) throw Error(QXB_ERR_ARG, "element size does not match the element kind");
            if (a.elem_size <= 0) throw Error(QXB_ERR_ARG, "bad element size");
        }
        jld2::write_file(path, arrays, commit_types != 0);
    });
}

uint32_t qxb_debug_lookup3(const void* data, size_t n, uint32_t initval) {
    return jld2::lookup3((const uint8_t*)data, n, initval);
}

int qxb_graph_load_jld2(qxb_graph* g, const char* path, int* n_set) {
    return guard([&] {
        if (!g || !path) throw Error(QXB_ERR_ARG, "null argument");
        jld2::File f = jld2::read_file(path);
        int n = 0;
        for (const jld2::Dataset& d : f.datasets) {
            if (d.kind == jld2::EK_OTHER || d.kind == jld2::EK_STRING) continue;
            if (d.count() == 0) continue;
            std::vector<std::complex<double>> v = jld2::as_c64(d);
            ok(qxb_graph_set_data(g, d.name.c_str(), v.data(), d.dims.data(), (int)d.dims.size()));
            ++n;
        }
        if (n_set) *n_set = n;
    });
}

int qxb_params_read(const char* yml_path, qxb_params* p, char* bitstrings, int64_t buflen) {
    return guard([&] {
        if (!yml_path || !p) throw Error(QXB_ERR_ARG, "null argument");
        Params r = read_params(yml_path);
        *p = r.p;
        if (!bitstrings) return;
        int64_t need = r.p.n_bitstrings * (r.p.num_qubits + 1);
        if (buflen < need) throw Error(QXB_ERR_ARG, "bitstring buffer too small: need " + std::to_string(need) + " bytes");
        for (int64_t i = 0; i < r.p.n_bitstrings; ++i)
            memcpy(bitstrings + i * (r.p.num_qubits + 1), r.bitstrings[(size_t)i].c_str(), (size_t)r.p.num_qubits + 1);
    });
}

int qxb_execute_files(const char* dsl_file, const char* input_file, const char* param_file, const char* output_file,
                      int dtype, int64_t max_amplitudes, int64_t max_slices, int replan_candidates,
                      int64_t* n_amplitudes, double* seconds) {
    return qxb_execute_files_multi(dsl_file, input_file, param_file, output_file, dtype, max_amplitudes, max_slices,
                                   replan_candidates, 1, 1, n_amplitudes, seconds);
}

int qxb_execute_files_multi(const char* dsl_file, const char* input_file, const char* param_file, const char* output_file,
                            int dtype, int64_t max_amplitudes, int64_t max_slices, int replan_candidates,
                            int n_devices, int sub_comm_size, int64_t* n_amplitudes, double* seconds) {
    qxb_multi* multi = nullptr;
    qxb_graph* g = nullptr;
    int rc = guard([&] {
        if (!dsl_file) throw Error(QXB_ERR_ARG, "no DSL file given");
        const std::string stem = stem_of(dsl_file);
        const std::string input = input_file && *input_file ? input_file : stem + ".jld2";    // qxrun.jl:25
        const std::string param = param_file && *param_file ? param_file : stem + ".yml";     // qxrun.jl:21
        double t0 = now();
        std::string text = read_text(dsl_file);
        Params pr = read_params(param);
        ok(qxb_graph_create(&g, dtype));
        ok(qxb_graph_parse_dsl(g, text.data(), text.size()));
        ok(qxb_graph_load_jld2(g, input.c_str(), nullptr));
        double t1 = now();

        std::vector<std::string>& bs = pr.bitstrings;
        if (max_amplitudes >= 0 && (int64_t)bs.size() > max_amplitudes) bs.resize((size_t)max_amplitudes);   // qxrun.jl:32-35
        const int64_t n_amp = (int64_t)bs.size();
        int n_out = 0, root_rank = 0;
        ok(qxb_graph_num_outputs(g, &n_out));
        ok(qxb_graph_root_dims(g, &root_rank, nullptr));
        if (root_rank != 0)
            throw Error(QXB_ERR_UNSUPP, "the program saves a tensor (open network): the file runner writes one amplitude per "
                                        "bitstring; use qxb_amplitudes with qxb_graph_root_dims for open networks");
        if (n_amp > 0 && n_out != (int)pr.p.num_qubits)
            throw Error(QXB_ERR_ARG, "the program has " + std::to_string(n_out) + " outputs but the parameter file gives " +
                                     std::to_string(pr.p.num_qubits) + "-qubit bitstrings");
        std::vector<uint8_t> bits((size_t)(n_amp * n_out));
        for (int64_t a = 0; a < n_amp; ++a)
            for (int q = 0; q < n_out; ++q) {
                char c = bs[(size_t)a][(size_t)q];
                bits[(size_t)(a * n_out + q)] = c == '0' ? 0 : c == '1' ? 1 : c == '+' ? 2 : 3;
            }
        if (replan_candidates > 0)      // plan for the batch the run will use (Rejection: candidate batches of 1024)
            ok(qxb_graph_replan(g, replan_candidates, pr.p.method == QXB_METHOD_REJECTION ? 1024 : (n_amp > 0 ? n_amp : 1), nullptr, nullptr));
        // several GPUs (qxrun -m): one replica per device for the List / Uniform methods; the rejection sampler is
        // sequential in its acceptance bound and stays on one device
        const bool use_multi = n_devices != 1 && pr.p.method != QXB_METHOD_REJECTION;
        if (use_multi) ok(qxb_multi_create(&multi, g, n_devices, nullptr, nullptr));
        else ok(qxb_graph_compile(g, nullptr));
        int64_t n_slices = 0;
        ok(qxb_graph_num_slices(g, &n_slices));
        if (max_slices >= 0 && max_slices < n_slices) n_slices = max_slices;                  // qxrun.jl:36-39
        double t2 = now();

        const size_t es = dtype == QXB_C32 ? 8 : 16;
        std::vector<uint8_t> amps((size_t)n_amp * es);
        double final_M = pr.p.M;
        int64_t drawn = 0;
        if (pr.p.method == QXB_METHOD_REJECTION) {
            // Rejection sampling from the circuit's output distribution (docs/src/features.md:68-84, outputs.jl:57-62):
            // draw a uniform bitstring, accept with probability p 2^n / M; unless fix_M, a larger ratio raises M
            // ("empirical supremum" sampling).  Candidates go down in batches of 1024 -- one qxb_amplitudes call each.
            // Stream: the library's splitmix64 (the reference: Julia's MersenneTwister, not reproducible elsewhere).
            if (n_out != (int)pr.p.num_qubits)
                throw Error(QXB_ERR_ARG, "the program has " + std::to_string(n_out) + " outputs but the parameter file says num_qubits = " +
                                         std::to_string(pr.p.num_qubits));
            if (!(pr.p.M > 0.0))
                throw Error(QXB_ERR_ARG, "rejection sampling needs M > 0 in the parameter file (the acceptance probability is p 2^n / M)");
            int64_t want = pr.p.num_samples;
            if (max_amplitudes >= 0 && want > max_amplitudes) want = max_amplitudes;
            uint64_t rs = pr.p.has_seed ? (uint64_t)pr.p.seed : 0x5851f42d4c957f2dull;
            const int64_t batch = 1024;
            const double N = std::ldexp(1.0, n_out);
            std::vector<uint8_t> cand((size_t)(batch * std::max(n_out, 1))), camp((size_t)batch * es);
            bs.clear(); amps.clear();
            for (int b = 0; b < 10000 && (int64_t)bs.size() < want; ++b) {
                uint64_t word = 0;
                for (int64_t i = 0; i < batch * n_out; ++i) {
                    if (i % 64 == 0) word = splitmix64(rs);
                    cand[(size_t)i] = (uint8_t)((word >> (i % 64)) & 1);
                }
                ok(qxb_amplitudes(g, cand.data(), batch, 0, n_slices, camp.data()));
                for (int64_t i = 0; i < batch && (int64_t)bs.size() < want; ++i) {
                    double re, im;
                    if (dtype == QXB_C32) { float v[2]; memcpy(v, &camp[(size_t)i * 8], 8); re = v[0]; im = v[1]; }
                    else { double v[2]; memcpy(v, &camp[(size_t)i * 16], 16); re = v[0]; im = v[1]; }
                    const double ratio = (re * re + im * im) * N;
                    const double u = (double)(splitmix64(rs) >> 11) / 9007199254740992.0;
                    ++drawn;
                    if (!pr.p.fix_M && ratio > final_M) final_M = ratio;
                    if (u < ratio / final_M) {
                        std::string str((size_t)n_out, '0');
                        for (int q = 0; q < n_out; ++q) str[(size_t)q] = (char)('0' + cand[(size_t)(i * n_out + q)]);
                        bs.push_back(str);
                        amps.insert(amps.end(), camp.begin() + (size_t)i * es, camp.begin() + (size_t)(i + 1) * es);
                    }
                }
            }
        } else if (n_amp > 0) {
            if (multi) ok(qxb_multi_amplitudes(multi, bits.data(), n_amp, 0, n_slices, sub_comm_size, amps.data()));
            else ok(qxb_amplitudes(g, bits.data(), n_amp, 0, n_slices, amps.data()));
        }
        const int64_t n_res = (int64_t)bs.size();
        double t3 = now();

        if (output_file && *output_file) {
            std::vector<char> names((size_t)(n_res * (n_out > 0 ? n_out : 1)));
            for (int64_t a = 0; a < n_res; ++a) memcpy(&names[(size_t)(a * n_out)], bs[(size_t)a].data(), (size_t)n_out);
            std::vector<jld2::WriteArray> arrays(2);
            arrays[0].name = "bitstrings"; arrays[0].kind = jld2::EK_STRING; arrays[0].elem_size = n_out > 0 ? n_out : 1;
            arrays[0].dims = {n_res}; arrays[0].data = names.data();
            arrays[1].name = "amplitudes"; arrays[1].kind = dtype == QXB_C32 ? jld2::EK_C32 : jld2::EK_C64;
            arrays[1].elem_size = (int)es; arrays[1].dims = {n_res}; arrays[1].data = amps.data();
            const double drawn_d = (double)drawn;
            if (pr.p.method == QXB_METHOD_REJECTION) {       // what the sampler ended with (the Python harness writes the same)
                jld2::WriteArray m; m.name = "M"; m.kind = jld2::EK_F64; m.elem_size = 8; m.data = &final_M;
                jld2::WriteArray d; d.name = "drawn"; d.kind = jld2::EK_F64; d.elem_size = 8; d.data = &drawn_d;
                arrays.push_back(m); arrays.push_back(d);
            }
            jld2::write_file(output_file, arrays, false);
        }
        double t4 = now();
        if (n_amplitudes) *n_amplitudes = n_res;
        if (seconds) { seconds[0] = t1 - t0; seconds[1] = t2 - t1; seconds[2] = t3 - t2; seconds[3] = t4 - t3; }
    });
    if (multi) qxb_multi_destroy(multi);
    if (g) qxb_graph_destroy(g);
    return rc;
}

}  // extern "C"
