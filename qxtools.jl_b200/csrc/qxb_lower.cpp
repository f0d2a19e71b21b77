// qxb200 -- DSL parser and the planner-side lowering of every `ncon` into a
// bit-level batched contraction (batch / M / N / K bit groups + address maps).
//
// Semantics followed: /root/reference/docs/src/users_guide.md:93-164 (DSL),
// /root/reference/src/compute_graph/compute_graph.jl:39-58 (views apply to leaves
// and chain), :61-74 (labels local to each command), tensor_cache.jl:52-53
// (column-major leaves).
//
// The B200-first idea: a `view` does not slice.  The sliced mode stays in the
// leaf as address bits tagged with its slice variable, and those bits ride
// through every contraction as batch bits -- so ONE launch of a node covers every
// slice assignment the node actually depends on (and every bitstring, if it
// depends on the outputs), and nodes that depend on few variables are computed
// only that many times.  Variables the caller's slice range pins ("fixed") become
// plain offsets into the leaf.
#include "qxb_ir.h"

#include <algorithm>
#include <climits>
#include <cmath>
#include <cstring>
#include <sstream>

#include "../../include/qxb200.h"

namespace qxb {

// sanity bounds on what a program may declare (a mode of 2^30 entries is already beyond any leaf or bond; without them
// a corrupt extent reaches the bit arithmetic: found by tests/test_fuzz_inputs.py)
static const int64_t kMaxExtent = int64_t(1) << 30;
static const int64_t kMaxOutputs = 65536;

// ---------------------------------------------------------------------- parse
static std::vector<int64_t> parse_list(const std::string& tok, bool labels) {
    std::vector<int64_t> out;
    if (labels && tok == "0") return out;      // scalar placeholder (users_guide.md:138-144)
    size_t i = 0;
    while (i < tok.size()) {
        size_t j = tok.find(',', i);
        if (j == std::string::npos) j = tok.size();
        std::string t = tok.substr(i, j - i);
        if (t.empty()) throw Error(QXB_ERR_ARG, "empty entry in list '" + tok + "'");
        char* end = nullptr;
        long long v = strtoll(t.c_str(), &end, 10);
        if (*end) throw Error(QXB_ERR_ARG, "bad integer '" + t + "'");
        out.push_back(v);
        i = j + 1;
    }
    return out;
}

static int64_t parse_int(const std::string& t) {
    char* end = nullptr;
    long long v = strtoll(t.c_str(), &end, 10);
    if (t.empty() || *end) throw Error(QXB_ERR_ARG, "bad integer '" + t + "'");
    return v;
}

void add_cmd(Program& p, const Cmd& c) {
    p.cmds.push_back(c);
    p.analysed = false;
}

void parse_dsl(Program& p, const char* text, size_t n) {
    std::string s(text, n);
    std::istringstream in(s);
    std::string line;
    bool first = true;
    int lineno = 0;
    while (std::getline(in, line)) {
        ++lineno;
        if (!line.empty() && line.back() == '\r') line.pop_back();
        if (first) {
            first = false;
            if (line.rfind("# version:", 0) != 0)
                throw Error(QXB_ERR_ARG, "first line of a .qx file must be '# version: x.y.z'");
            continue;
        }
        size_t k = line.find_first_not_of(" \t");
        if (k == std::string::npos || line[k] == '#') continue;
        std::istringstream ls(line);
        std::vector<std::string> t;
        std::string w;
        while (ls >> w) t.push_back(w);
        auto need = [&](size_t m) {
            if (t.size() != m)
                throw Error(QXB_ERR_ARG, "line " + std::to_string(lineno) + ": '" + t[0] + "' takes " +
                                             std::to_string(m - 1) + " arguments");
        };
        Cmd c;
        if (t[0] == "load") {
            need(4); c.kind = CMD_LOAD; c.name = t[1]; c.label = t[2]; c.dims = parse_list(t[3], false);
        } else if (t[0] == "output") {
            need(4); c.kind = CMD_OUTPUT; c.name = t[1]; c.idx = parse_int(t[2]); c.dim = parse_int(t[3]);
        } else if (t[0] == "view") {
            need(6); c.kind = CMD_VIEW; c.name = t[1]; c.a = t[2]; c.label = t[3];
            c.idx = parse_int(t[4]); c.dim = parse_int(t[5]);
        } else if (t[0] == "ncon") {
            need(7); c.kind = CMD_NCON; c.name = t[1]; c.cl = parse_list(t[2], true);
            c.a = t[3]; c.al = parse_list(t[4], true); c.b = t[5]; c.bl = parse_list(t[6], true);
        } else if (t[0] == "save") {
            need(3); c.kind = CMD_SAVE; c.name = t[1]; c.a = t[2];
        } else {
            throw Error(QXB_ERR_ARG, "line " + std::to_string(lineno) + ": unknown instruction '" + t[0] + "'");
        }
        add_cmd(p, c);
    }
    if (first) throw Error(QXB_ERR_ARG, "empty .qx program");
}

// -------------------------------------------------------------------- analyse
static int var_number(const std::string& sym) {
    if (sym.size() < 2 || sym[0] != 'v') return -1;
    for (size_t i = 1; i < sym.size(); ++i)
        if (sym[i] < '0' || sym[i] > '9') return -1;
    return atoi(sym.c_str() + 1);
}

void analyse(Program& p) {
    p.defs.clear(); p.by_name.clear(); p.vars.clear(); p.var_by_sym.clear();
    p.n_outputs = 0; p.root = -1;
    // slice symbols v1..vk in numeric order (compute_graph.jl:42)
    std::map<int, std::pair<std::string, int64_t>> found;
    for (const Cmd& c : p.cmds) {
        if (c.kind != CMD_VIEW) continue;
        int num = var_number(c.label);
        if (num < 0) throw Error(QXB_ERR_ARG, "slice symbol '" + c.label + "' is not of the form v<N>");
        if (c.dim < 1 || c.dim > kMaxExtent) throw Error(QXB_ERR_ARG, "view " + c.name + ": bad bond dimension");
        auto it = found.find(num);
        if (it == found.end()) found[num] = {c.label, c.dim};
        else if (it->second.second != c.dim)
            throw Error(QXB_ERR_ARG, "inconsistent extent for slice symbol " + c.label);
    }
    for (auto& kv : found) {
        p.var_by_sym[kv.second.first] = (int)p.vars.size();
        p.vars.push_back(SliceVar{kv.second.first, kv.second.second, ceil_log2(kv.second.second)});
    }
    auto define = [&](TensorDef&& d) {
        if (p.by_name.count(d.name)) throw Error(QXB_ERR_ARG, "symbol '" + d.name + "' defined twice");
        p.by_name[d.name] = (int)p.defs.size();
        p.defs.push_back(std::move(d));
    };
    auto lookup = [&](const std::string& n, const std::string& who) -> int {
        auto it = p.by_name.find(n);
        if (it == p.by_name.end()) throw Error(QXB_ERR_ARG, who + ": symbol '" + n + "' used before definition");
        return it->second;
    };
    for (const Cmd& c : p.cmds) {
        switch (c.kind) {
        case CMD_LOAD: {
            TensorDef d; d.kind = T_LOAD; d.name = c.name; d.data_label = c.label;
            for (int64_t e : c.dims) {
                if (e < 1 || e > kMaxExtent) throw Error(QXB_ERR_ARG, "load " + c.name + ": bad dimension");
                d.modes.push_back(Mode{e, e, ceil_log2(e), -1});
            }
            d.leaf = (int)p.defs.size();
            define(std::move(d));
            break;
        }
        case CMD_OUTPUT: {
            if (c.idx < 1 || c.idx > kMaxOutputs) throw Error(QXB_ERR_ARG, "output " + c.name + ": index is 1-based (and at most 65536)");
            if (c.dim < 2 || c.dim > kMaxExtent) throw Error(QXB_ERR_ARG, "output " + c.name + ": dimension must be >= 2");
            TensorDef d; d.kind = T_OUTPUT; d.name = c.name; d.out_idx = c.idx; d.amp = true;
            d.modes.push_back(Mode{c.dim, c.dim, ceil_log2(c.dim), -1});
            d.leaf = (int)p.defs.size();
            p.n_outputs = std::max<int>(p.n_outputs, (int)c.idx);
            define(std::move(d));
            break;
        }
        case CMD_VIEW: {
            int t = lookup(c.a, "view " + c.name);
            const TensorDef& src = p.defs[t];
            if (src.kind == T_NCON)
                throw Error(QXB_ERR_UNSUPP, "view " + c.name + ": views of contraction results are not supported "
                                            "(compute_graph.jl:41-56 only views leaves)");
            if (c.idx < 1 || c.idx > (int64_t)src.modes.size())
                throw Error(QXB_ERR_ARG, "view " + c.name + ": mode " + std::to_string(c.idx) + " out of range");
            TensorDef d; d.kind = T_VIEW; d.name = c.name; d.modes = src.modes; d.leaf = src.leaf;
            d.vars = src.vars; d.amp = src.amp; d.data_label = src.data_label; d.out_idx = src.out_idx;
            Mode& m = d.modes[c.idx - 1];
            if (m.var != -1) throw Error(QXB_ERR_ARG, "view " + c.name + ": mode already sliced");
            if (m.ext != c.dim)
                throw Error(QXB_ERR_ARG, "view " + c.name + ": bond dimension " + std::to_string(c.dim) +
                                             " does not match mode extent " + std::to_string(m.ext));
            int v = p.var_by_sym.at(c.label);
            for (const Mode& o : d.modes)
                if (o.var == v) throw Error(QXB_ERR_UNSUPP, "view " + c.name + ": slice symbol applied twice to one tensor");
            m.var = v; m.ext = 1;
            d.vars.insert(v);
            p.defs[t].uses++;
            define(std::move(d));
            break;
        }
        case CMD_NCON: {
            int a = lookup(c.a, "ncon " + c.name), b = lookup(c.b, "ncon " + c.name);
            const TensorDef &A = p.defs[a], &B = p.defs[b];
            if (c.al.size() != A.modes.size() || c.bl.size() != B.modes.size())
                throw Error(QXB_ERR_ARG, "ncon " + c.name + ": label count does not match tensor rank");
            std::map<int64_t, Mode> ext;
            auto scan = [&](const std::vector<int64_t>& ls, const TensorDef& T) {
                std::set<int64_t> seen;
                for (size_t i = 0; i < ls.size(); ++i) {
                    if (ls[i] < 1) throw Error(QXB_ERR_ARG, "ncon " + c.name + ": labels must be positive");
                    if (!seen.insert(ls[i]).second)
                        throw Error(QXB_ERR_UNSUPP, "ncon " + c.name + ": repeated label inside one tensor");
                    auto it = ext.find(ls[i]);
                    if (it == ext.end()) ext[ls[i]] = T.modes[i];
                    else if (it->second.ext != T.modes[i].ext)
                        throw Error(QXB_ERR_ARG, "ncon " + c.name + ": extent mismatch on label " + std::to_string(ls[i]));
                }
            };
            scan(c.al, A); scan(c.bl, B);
            TensorDef d; d.kind = T_NCON; d.name = c.name; d.a = a; d.b = b;
            d.cl = c.cl; d.al = c.al; d.bl = c.bl;
            std::set<int64_t> seen;
            for (int64_t l : c.cl) {
                if (!seen.insert(l).second) throw Error(QXB_ERR_ARG, "ncon " + c.name + ": duplicate output label");
                auto it = ext.find(l);
                if (it == ext.end()) throw Error(QXB_ERR_ARG, "ncon " + c.name + ": output label not on any input");
                d.modes.push_back(it->second);
            }
            d.vars = A.vars; d.vars.insert(B.vars.begin(), B.vars.end());
            d.amp = A.amp || B.amp;
            p.defs[a].uses++; p.defs[b].uses++;
            define(std::move(d));
            break;
        }
        case CMD_SAVE: {
            int t = lookup(c.a, "save");
            if (p.root != -1) throw Error(QXB_ERR_UNSUPP, "more than one save instruction");
            p.root = t;
            break;
        }
        }
    }
    if (p.root < 0) throw Error(QXB_ERR_ARG, "program has no save instruction");
    p.analysed = true;
}

int64_t num_slices(const Program& p) {
    int64_t n = 1;
    for (const SliceVar& v : p.vars) n *= v.dim;
    return n;
}

void slice_values(const Program& p, int64_t s, int64_t* out) {
    for (size_t i = 0; i < p.vars.size(); ++i) { out[i] = s % p.vars[i].dim; s /= p.vars[i].dim; }
}

// ---------------------------------------------------------------------- lower
namespace {
const int NONE = INT_MIN;
struct Item { int nbits; int posA, posB; int keyC; int posC; };

std::vector<Seg> make_segs(const std::vector<std::pair<int, std::pair<int, int>>>& parts) {
    // parts: (src, (dst, len)) sorted by src; merge runs contiguous in both spaces
    std::vector<Seg> out;
    for (auto& pr : parts) {
        int src = pr.first, dst = pr.second.first, len = pr.second.second;
        if (!out.empty() && out.back().src + out.back().len == src && out.back().dst + out.back().len == dst)
            out.back().len = (uint8_t)(out.back().len + len);
        else
            out.push_back(Seg{(uint8_t)src, (uint8_t)dst, (uint8_t)len, 0});
    }
    return out;
}
int find_entry(const LTensor& t, int key) {
    for (const LayEntry& e : t.lay) if (e.key == key) return e.pos;
    return -1;
}
}  // namespace

Lowered lower(const Program& p, uint64_t free_mask, bool early_sum) {
    if (!p.analysed) throw Error(QXB_ERR_STATE, "program not analysed");
    const int k = (int)p.vars.size();
    if (k > 63) throw Error(QXB_ERR_UNSUPP, "more than 63 slice variables");
    free_mask &= low_mask(k);
    auto is_free = [&](int v) { return ((free_mask >> v) & 1ull) != 0; };
    Lowered L; L.free_mask = free_mask; L.n_free = __builtin_popcountll(free_mask);
    std::vector<int> producer;                    // per LTensor: op that writes it (-1: leaf)
    std::vector<int> lt_of_def(p.defs.size(), -1);
    // Early summation of a batched slice variable: v is summed (becomes a K index) at
    // the node whose subtree holds ALL leaf uses that depend on v -- sum-product
    // reordering, exact because nothing outside that subtree depends on v.  Needs a
    // tree: if any contraction result is used twice, fall back to summing at the root.
    std::vector<int> total_uses(k, 0);
    for (const TensorDef& d : p.defs) {
        if (d.kind != T_NCON) continue;
        if (d.uses > 1) early_sum = false;
        for (int o : {d.a, d.b})
            if (p.defs[o].kind != T_NCON)
                for (int v : p.defs[o].vars) total_uses[v]++;
    }
    L.early_sum = early_sum;
    std::vector<std::map<int, int>> use_cnt;      // per LTensor: var -> leaf uses inside its subtree

    auto leaf_tensor = [&](int di) -> int {
        if (lt_of_def[di] >= 0) return lt_of_def[di];
        const TensorDef& d = p.defs[di];
        const TensorDef& base = p.defs[d.leaf];
        LTensor t; t.def = di; t.is_leaf = true;
        int pos = 0;
        for (size_t m = 0; m < d.modes.size(); ++m) {
            const Mode& md = d.modes[m];
            if (md.var == -1) {
                if (md.nbits > 0) t.lay.push_back(LayEntry{(int)m, md.nbits, pos});
            } else if (is_free(md.var)) {
                if (md.full_ext != p.vars[md.var].dim)
                    throw Error(QXB_ERR_ARG, "view on " + d.name + ": extent does not match the slice variable");
                if (md.nbits > 0) t.lay.push_back(LayEntry{~md.var, md.nbits, pos});
            } else {
                t.fixed.push_back({md.var, pos});
            }
            pos += md.nbits;
        }
        t.span_bits = pos;
        if (pos > 40) throw Error(QXB_ERR_UNSUPP, "leaf tensor " + d.name + " is too large");
        if (base.kind == T_OUTPUT) {
            t.is_output_leaf = true; t.amp = true; t.phase = PH_CHUNK; t.out_idx = base.out_idx;
            t.persistent = true;
            L.output_leaves.push_back((int)L.tensors.size());
        } else {
            t.data_label = base.data_label;
            t.phase = d.vars.empty() ? PH_CONST : PH_BLOCK;
        }
        lt_of_def[di] = (int)L.tensors.size();
        L.tensors.push_back(std::move(t));
        std::map<int, int> cnt;
        for (int v : d.vars) cnt[v] = 1;
        use_cnt.push_back(std::move(cnt));
        producer.push_back(-1);
        return lt_of_def[di];
    };

    for (size_t di = 0; di < p.defs.size(); ++di) {
        const TensorDef& d = p.defs[di];
        if (d.kind != T_NCON) continue;
        int ia_t = p.defs[d.a].kind == T_NCON ? lt_of_def[d.a] : leaf_tensor(d.a);
        int ib_t = p.defs[d.b].kind == T_NCON ? lt_of_def[d.b] : leaf_tensor(d.b);
        if (ia_t < 0 || ib_t < 0) throw Error(QXB_ERR_STATE, "ncon " + d.name + ": operand not lowered");
        const TensorDef &DA = p.defs[d.a], &DB = p.defs[d.b];
        std::vector<Item> items;
        // ordinary labels
        std::set<int64_t> labels(d.al.begin(), d.al.end());
        labels.insert(d.bl.begin(), d.bl.end());
        for (int64_t l : labels) {
            int ia = -1, ib = -1, ic = -1;
            for (size_t i = 0; i < d.al.size(); ++i) if (d.al[i] == l) ia = (int)i;
            for (size_t i = 0; i < d.bl.size(); ++i) if (d.bl[i] == l) ib = (int)i;
            for (size_t i = 0; i < d.cl.size(); ++i) if (d.cl[i] == l) ic = (int)i;
            int nb = 0;
            int posA = -1, posB = -1;
            if (ia >= 0 && DA.modes[ia].var == -1 && DA.modes[ia].nbits > 0) {
                nb = DA.modes[ia].nbits; posA = find_entry(L.tensors[ia_t], ia);
                if (posA < 0) throw Error(QXB_ERR_STATE, "ncon " + d.name + ": lost mode of A");
            }
            if (ib >= 0 && DB.modes[ib].var == -1 && DB.modes[ib].nbits > 0) {
                if (nb && nb != DB.modes[ib].nbits) throw Error(QXB_ERR_ARG, "ncon " + d.name + ": extent mismatch");
                nb = DB.modes[ib].nbits; posB = find_entry(L.tensors[ib_t], ib);
                if (posB < 0) throw Error(QXB_ERR_STATE, "ncon " + d.name + ": lost mode of B");
            }
            if (nb == 0) continue;                  // sliced or extent-1 mode: no address bits
            items.push_back(Item{nb, posA, posB, ic >= 0 ? ic : NONE, -1});
        }
        // free slice variables ride along as batch bits
        std::set<int> vs(DA.vars.begin(), DA.vars.end());
        vs.insert(DB.vars.begin(), DB.vars.end());
        std::map<int, int> cntC = use_cnt[ia_t];
        for (auto& kv : use_cnt[ib_t]) cntC[kv.first] += kv.second;
        for (int v : vs) {
            if (!is_free(v) || p.vars[v].nbits == 0) continue;
            int posA = find_entry(L.tensors[ia_t], ~v), posB = find_entry(L.tensors[ib_t], ~v);
            if (posA < 0 && posB < 0) continue;          // already summed further down
            const bool done = early_sum && cntC[v] == total_uses[v];
            items.push_back(Item{p.vars[v].nbits, posA, posB, done ? NONE : ~v, -1});
        }
        // layout of C: follow the bigger operand's bit order, then the other's extra bits
        const bool a_big = L.tensors[ia_t].span_bits >= L.tensors[ib_t].span_bits;
        std::vector<int> cidx, kidx;
        for (size_t i = 0; i < items.size(); ++i) (items[i].keyC == NONE ? kidx : cidx).push_back((int)i);
        auto rank = [&](const Item& it) {
            int pb = a_big ? it.posA : it.posB, po = a_big ? it.posB : it.posA;
            return pb >= 0 ? std::make_pair(0, pb) : std::make_pair(1, po);
        };
        std::sort(cidx.begin(), cidx.end(), [&](int x, int y) { return rank(items[x]) < rank(items[y]); });
        std::sort(kidx.begin(), kidx.end(), [&](int x, int y) { return rank(items[x]) < rank(items[y]); });
        LTensor C; C.def = (int)di;
        LOp op; op.name = d.name;
        int pc = 0;
        std::vector<std::pair<int, std::pair<int, int>>> pa, pb;
        for (int i : cidx) {
            Item& it = items[i];
            it.posC = pc;
            C.lay.push_back(LayEntry{it.keyC, it.nbits, pc});
            if (it.posA >= 0) pa.push_back({pc, {it.posA, it.nbits}});
            if (it.posB >= 0) pb.push_back({pc, {it.posB, it.nbits}});
            if (it.posA >= 0 && it.posB >= 0) op.n_batch += it.nbits;
            else if (it.posA >= 0) op.n_m += it.nbits;
            else op.n_n += it.nbits;
            pc += it.nbits;
        }
        op.nC = pc; C.span_bits = pc;
        op.segA = make_segs(pa); op.segB = make_segs(pb);
        pa.clear(); pb.clear();
        int pk = 0;
        for (int i : kidx) {
            Item& it = items[i];
            if (it.posA >= 0) pa.push_back({pk, {it.posA, it.nbits}});
            if (it.posB >= 0) pb.push_back({pk, {it.posB, it.nbits}});
            pk += it.nbits;
        }
        op.nK = pk;
        op.segKA = make_segs(pa); op.segKB = make_segs(pb);
        if (op.nC > 46 || op.nK > 30) throw Error(QXB_ERR_UNSUPP, "ncon " + d.name + ": tensor too large");
        C.amp = L.tensors[ia_t].amp || L.tensors[ib_t].amp;
        C.phase = C.amp ? PH_CHUNK : (d.vars.empty() ? PH_CONST : PH_BLOCK);
        op.phase = C.phase;
        op.a = ia_t; op.b = ib_t; op.c = (int)L.tensors.size();
        op.macs_per_amp = std::ldexp(1.0, op.nC + op.nK);
        op.elems_a = std::ldexp(1.0, (int)[&] { int s = 0; for (auto& e : L.tensors[ia_t].lay) s += e.nbits; return s; }());
        op.elems_b = std::ldexp(1.0, (int)[&] { int s = 0; for (auto& e : L.tensors[ib_t].lay) s += e.nbits; return s; }());
        op.elems_c = std::ldexp(1.0, op.nC);
        const int oi = (int)L.ops.size();
        for (int t : {ia_t, ib_t}) {
            LTensor& T = L.tensors[t];
            if (T.first_use < 0) T.first_use = oi;
            T.last_use = oi;
            if (T.phase != op.phase) T.persistent = true;
        }
        for (int t : {ia_t, ib_t})
            if (producer[t] >= 0 && L.ops[producer[t]].phase != PH_CONST) op.deps.push_back(producer[t]);
        lt_of_def[di] = op.c;
        L.tensors.push_back(std::move(C));
        use_cnt.push_back(std::move(cntC));
        producer.push_back(oi);
        L.ops.push_back(std::move(op));
    }
    const TensorDef& R = p.defs[p.root];
    if (R.kind != T_NCON) throw Error(QXB_ERR_UNSUPP, "the saved tensor must be the result of a contraction");
    L.root = lt_of_def[p.root];
    L.tensors[L.root].persistent = true;
    // A closed network saves a scalar.  An open one (convert_to_tnc(...; no_output=true), the case the reference's own
    // contract_tn! tests use, test/test_contraction_planning.jl:58-61) saves a tensor: its modes stay address bits
    // of the root and are gathered into Julia order by reduce_root_open.
    L.root_modes.assign(R.modes.size(), Lowered::RootMode{1, 0, 0});
    for (size_t m = 0; m < R.modes.size(); ++m) {
        // a sliced mode that stays open would make the slices different ELEMENTS of the result, not terms of a sum
        if (R.modes[m].var >= 0)
            throw Error(QXB_ERR_UNSUPP, "mode " + std::to_string(m + 1) + " of the saved tensor is sliced (" +
                                        p.vars[R.modes[m].var].sym + "): open indices cannot be slice bonds");
        L.root_modes[m].ext = R.modes[m].ext;
    }
    for (const LayEntry& e : L.tensors[L.root].lay) {
        if (e.key >= 0) {
            if (e.key >= (int)R.modes.size()) throw Error(QXB_ERR_STATE, "root layout names a mode the saved tensor does not have");
            L.root_modes[e.key].nbits = e.nbits;
            L.root_modes[e.key].pos = e.pos;
        } else {
            L.root_vars.push_back(~e.key);
        }
    }
    L.root_elems = 1;
    for (const auto& rm : L.root_modes) {
        if (rm.ext > (int64_t(1) << rm.nbits)) throw Error(QXB_ERR_STATE, "root mode wider than its address bits");
        L.root_elems *= rm.ext;
    }
    if (L.root_modes.size() > 32) throw Error(QXB_ERR_UNSUPP, "the saved tensor has more than 32 modes");
    L.root_scale = 1.0;
    for (int v = 0; v < k; ++v)
        if (is_free(v) && !R.vars.count(v)) L.root_scale *= (double)p.vars[v].dim;
    return L;
}

// ---------------------------------------------------------------- memory plan
namespace {
struct Arena {
    // first-fit free list over [0, inf); offsets in elements
    std::vector<std::pair<int64_t, int64_t>> free_;   // (offset, size), sorted by offset
    int64_t top = 0, peak = 0;
    int64_t alloc(int64_t n) {
        for (size_t i = 0; i < free_.size(); ++i) {
            if (free_[i].second >= n) {
                int64_t off = free_[i].first;
                free_[i].first += n; free_[i].second -= n;
                if (free_[i].second == 0) free_.erase(free_.begin() + i);
                return off;
            }
        }
        // grow: extend a trailing free block if it touches the top
        if (!free_.empty() && free_.back().first + free_.back().second == top) {
            int64_t off = free_.back().first;
            top = off + n; free_.pop_back();
            peak = std::max(peak, top);
            return off;
        }
        int64_t off = top; top += n; peak = std::max(peak, top);
        return off;
    }
    void release(int64_t off, int64_t n) {
        auto it = std::lower_bound(free_.begin(), free_.end(), std::make_pair(off, (int64_t)0));
        it = free_.insert(it, {off, n});
        size_t i = it - free_.begin();
        if (i + 1 < free_.size() && free_[i].first + free_[i].second == free_[i + 1].first) {
            free_[i].second += free_[i + 1].second; free_.erase(free_.begin() + i + 1);
        }
        if (i > 0 && free_[i - 1].first + free_[i - 1].second == free_[i].first) {
            free_[i - 1].second += free_[i].second; free_.erase(free_.begin() + i);
        }
    }
};
inline int64_t unit_elems(const LTensor& t) { return std::max<int64_t>(2, int64_t(1) << t.span_bits); }
}  // namespace

void plan_memory(Lowered& L) {
    // Chunk-phase tensors all carry the amplitude axis, so their offsets are planned
    // per amplitude row and scaled by the batch size at launch time.
    //
    // The step runs as a dependency graph, not a serial stream: a buffer that reuses
    // freed arena space must wait for the last reader of whatever lived there
    // (write-after-read edge).  Small tensors are never recycled, so the hundreds of
    // tiny nodes stay free of false dependencies; only the big ones share space.
    const int64_t kNoReuseBelow = 512;       // elements per amplitude row
    Arena ar[3];
    struct Freed { int phase; int64_t off, size; int last_reader; };
    std::vector<Freed> freed;
    std::vector<int> deferred;               // operands of a fused launch waiting for its last op
    for (int ti : L.output_leaves) {
        LTensor& t = L.tensors[ti];
        t.offset = ar[PH_CHUNK].alloc(unit_elems(t));
    }
    for (size_t oi = 0; oi < L.ops.size(); ++oi) {
        LOp& op = L.ops[oi];
        LTensor& C = L.tensors[op.c];
        const int64_t n = unit_elems(C);
        C.offset = ar[op.phase].alloc(n);
        for (const Freed& f : freed)
            if (f.phase == (int)op.phase && f.off < C.offset + n && C.offset < f.off + f.size &&
                std::find(op.deps.begin(), op.deps.end(), f.last_reader) == op.deps.end())
                op.deps.push_back(f.last_reader);
        // Inside a fused launch nothing is released before its last op: one launch works through all rows of the batch,
        // so a result row written early (any op's C that leaves the launch) must not land on an operand row another CTA has
        // still to read.  The deferred operands go back to the arena at the last op, after its C has its place.
        const bool in_fused = L.fused_first >= 0 && (int)oi >= L.fused_first && (int)oi <= L.fused_last;
        for (int ti : {op.a, op.b}) {
            LTensor& T = L.tensors[ti];
            if (T.is_leaf || T.persistent || T.offset < 0) continue;
            if (T.last_use == (int)oi && T.phase == op.phase && unit_elems(T) >= kNoReuseBelow) {
                if (in_fused && (int)oi < L.fused_last) { deferred.push_back(ti); T.last_use = -2; continue; }
                ar[T.phase].release(T.offset, unit_elems(T));
                freed.push_back(Freed{(int)T.phase, T.offset, unit_elems(T), (int)oi});
                T.last_use = -2;          // released once, even if it is both operands
            }
        }
        if (in_fused && (int)oi == L.fused_last) {
            for (int ti : deferred) {
                LTensor& T = L.tensors[ti];
                ar[T.phase].release(T.offset, unit_elems(T));
                freed.push_back(Freed{(int)T.phase, T.offset, unit_elems(T), (int)oi});
            }
            deferred.clear();
        }
    }
    L.const_elems = ar[PH_CONST].peak;
    L.block_elems = ar[PH_BLOCK].peak;
    L.chunk_elems_per_amp = ar[PH_CHUNK].peak;
}

double lowered_cost_bytes(const Lowered& L, double n_amp, double elem_bytes) {
    double b = 0;
    for (const LOp& op : L.ops) {
        if (op.phase == PH_CONST) continue;
        b += op.elems_a * (L.tensors[op.a].amp ? n_amp : 1) + op.elems_b * (L.tensors[op.b].amp ? n_amp : 1) +
             op.elems_c * (L.tensors[op.c].amp ? n_amp : 1);
    }
    return b * elem_bytes;
}

std::vector<int> partition_vars(const Program& p, int n_parts, bool early_sum) {
    const int k = (int)p.vars.size();
    uint64_t mask = low_mask(k);
    std::vector<int> chosen;
    int64_t remaining = n_parts;
    while (remaining > 1) {
        int best = -1; double best_cost = 0;
        for (int v = 0; v < k; ++v) {
            if (!((mask >> v) & 1ull) || p.vars[v].dim < 2 || remaining % p.vars[v].dim) continue;
            Lowered L = lower(p, mask & ~(1ull << v), early_sum);
            const double c = lowered_cost_bytes(L, 1024.0, 1.0);
            if (best < 0 || c < best_cost) { best = v; best_cost = c; }
        }
        if (best < 0) return {};              // the extents do not factor n_parts: caller falls back to ranges
        chosen.push_back(best);
        mask &= ~(1ull << best);
        remaining /= p.vars[best].dim;
    }
    return chosen;
}

// ------------------------------------------------------------------- describe
static void seg_json(std::ostringstream& o, const std::vector<Seg>& s) {
    o << "[";
    for (size_t i = 0; i < s.size(); ++i)
        o << (i ? "," : "") << "[" << (int)s[i].src << "," << (int)s[i].dst << "," << (int)s[i].len << "]";
    o << "]";
}

std::string describe_json(const Program& p, const Lowered& L) {
    std::ostringstream o;
    o << "{\"early_sum\":" << (L.early_sum ? "true" : "false") << ",\"n_free\":" << L.n_free << ",\"free_mask\":" << L.free_mask << ",\"n_slice_vars\":" << p.vars.size() << ",\"n_outputs\":" << p.n_outputs
      << ",\"slice_vars\":[";
    for (size_t i = 0; i < p.vars.size(); ++i)
        o << (i ? "," : "") << "{\"sym\":\"" << p.vars[i].sym << "\",\"dim\":" << p.vars[i].dim << "}";
    o << "],\"root_modes\":[";
    for (size_t i = 0; i < L.root_modes.size(); ++i)
        o << (i ? "," : "") << "[" << L.root_modes[i].ext << "," << L.root_modes[i].nbits << "," << L.root_modes[i].pos << "]";
    o << "],\"root_scale\":" << L.root_scale << ",\"root\":" << L.root << ",\"arena_elems\":{\"const\":" << L.const_elems
      << ",\"block\":" << L.block_elems << ",\"chunk_per_amp\":" << L.chunk_elems_per_amp << "},\"tensors\":[";
    static const char* ph[] = {"const", "block", "chunk"};
    for (size_t i = 0; i < L.tensors.size(); ++i) {
        const LTensor& t = L.tensors[i];
        o << (i ? "," : "") << "{\"name\":\"" << p.defs[t.def].name << "\",\"span_bits\":" << t.span_bits
          << ",\"amp\":" << (t.amp ? "true" : "false") << ",\"phase\":\"" << ph[t.phase] << "\",\"leaf\":"
          << (t.is_leaf ? "true" : "false") << ",\"output_leaf\":" << (t.is_output_leaf ? "true" : "false")
          << ",\"out_idx\":" << t.out_idx << ",\"data_label\":\"" << t.data_label << "\",\"offset\":" << t.offset
          << ",\"persistent\":" << (t.persistent ? "true" : "false") << ",\"fixed\":[";
        for (size_t j = 0; j < t.fixed.size(); ++j)
            o << (j ? "," : "") << "[" << t.fixed[j].first << "," << t.fixed[j].second << "]";
        o << "],\"lay\":[";
        for (size_t j = 0; j < t.lay.size(); ++j)
            o << (j ? "," : "") << "[" << t.lay[j].key << "," << t.lay[j].nbits << "," << t.lay[j].pos << "]";
        o << "]}";
    }
    o << "],\"ops\":[";
    for (size_t i = 0; i < L.ops.size(); ++i) {
        const LOp& op = L.ops[i];
        o << (i ? "," : "") << "{\"name\":\"" << op.name << "\",\"a\":" << op.a << ",\"b\":" << op.b << ",\"c\":" << op.c
          << ",\"phase\":\"" << ph[op.phase] << "\",\"nC\":" << op.nC
          << ",\"nK\":" << op.nK << ",\"batch_bits\":" << op.n_batch << ",\"m_bits\":" << op.n_m
          << ",\"n_bits\":" << op.n_n << ",\"a_bits\":" << std::ilogb(op.elems_a) << ",\"b_bits\":" << std::ilogb(op.elems_b)
          << ",\"amp\":" << (L.tensors[op.c].amp ? "true" : "false") << ",\"a_amp\":" << (L.tensors[op.a].amp ? "true" : "false")
          << ",\"b_amp\":" << (L.tensors[op.b].amp ? "true" : "false") << ",\"segA\":";
        seg_json(o, op.segA); o << ",\"segB\":"; seg_json(o, op.segB);
        o << ",\"segKA\":"; seg_json(o, op.segKA); o << ",\"segKB\":"; seg_json(o, op.segKB);
        o << ",\"deps\":[";
        for (size_t j = 0; j < op.deps.size(); ++j) o << (j ? "," : "") << op.deps[j];
        o << "]}";
    }
    o << "]}";
    return o.str();
}

}  // namespace qxb
