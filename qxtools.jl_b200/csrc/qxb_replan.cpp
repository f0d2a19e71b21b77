// qxb200 -- batch-aware re-planning of the ncon tree (planner-side lowering).
//
// QXTools plans for one slice of one bitstring: contraction_scheme
// (/root/reference/src/contraction_planning.jl:219-299) takes the sliced hyper-edges out
// of the line graph and orders the rest.  This executor batches the sliced hyper-edges and
// the bitstrings, so that order drags batch bits through every large intermediate.  Here
// the tensor network behind the program is recovered (index identity = labels matched in
// an ncon), new elimination orders are drawn (min-fill on the line graph of the UNSLICED
// network, contraction_planning.jl:47-63,127-176, seeded random tie-breaking), turned into
// pairwise plans (:386-448, smallest batched result first inside a hyper-edge) and scored
// with the executor's own cost model (bytes moved by the lowered program).  Leaves and
// views are kept verbatim; only the ncon statements are re-derived -- an exact
// re-association, the program's value is unchanged.
#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cmath>
#include <map>
#include <random>
#include <set>
#include <sstream>

#include "../../include/qxb200.h"
#include "qxb_ir.h"
#include "qxb_treeopt.h"

namespace qxb {

namespace {

struct UF {
    std::vector<int> p;
    int make() { p.push_back((int)p.size()); return (int)p.size() - 1; }
    int find(int x) { while (p[x] != x) { p[x] = p[p[x]]; x = p[x]; } return x; }
    void unite(int a, int b) { a = find(a); b = find(b); if (a != b) p[b] = a; }
};

struct Net {
    std::vector<std::string> names;            // leaf tensors (top of their view chains)
    std::vector<std::vector<int>> ids;         // classes per leaf (deduplicated), AMP appended for outputs
    std::vector<std::vector<int>> modes;       // class per DSL mode (for label emission)
    std::vector<double> dim;                   // extent per class (AMP last)
    int amp = -1;
};

// min-fill elimination order over classes [0, n) (AMP excluded by the caller)
std::vector<int> min_fill_order(std::vector<std::set<int>> adj, std::mt19937_64* rng) {
    const int n = (int)adj.size();
    std::vector<char> alive(n, 1);
    auto fill = [&](int v) {
        std::vector<int> nb(adj[v].begin(), adj[v].end());
        int f = 0;
        for (size_t i = 0; i < nb.size(); ++i)
            for (size_t j = i + 1; j < nb.size(); ++j)
                if (!adj[nb[i]].count(nb[j])) ++f;
        return f;
    };
    std::vector<int> cost(n, 0);
    for (int v = 0; v < n; ++v) cost[v] = fill(v);
    std::vector<int> order;
    for (int step = 0; step < n; ++step) {
        int best = -1;
        for (int v = 0; v < n; ++v) if (alive[v] && (best < 0 || cost[v] < best)) best = cost[v];
        std::vector<int> cands;
        size_t mind = SIZE_MAX;
        for (int v = 0; v < n; ++v)
            if (alive[v] && cost[v] == best) mind = std::min(mind, adj[v].size());
        for (int v = 0; v < n; ++v)
            if (alive[v] && cost[v] == best && adj[v].size() == mind) cands.push_back(v);
        int v = cands[0];
        if (rng && cands.size() > 1) v = cands[(size_t)((*rng)() % cands.size())];
        std::set<int> nb = adj[v];
        alive[v] = 0;
        order.push_back(v);
        std::set<int> touched(nb.begin(), nb.end());
        for (int a : nb) adj[a].erase(v);
        for (int a : nb)
            for (int b : nb)
                if (a != b && adj[a].insert(b).second) { touched.insert(adj[a].begin(), adj[a].end()); }
        for (int a : nb) touched.insert(adj[a].begin(), adj[a].end());
        adj[v].clear();
        for (int a : touched) if (alive[a]) cost[a] = fill(a);
    }
    return order;
}

struct PlanStep { int a, b; };                 // indices into the growing tensor list

// elimination order -> pairwise plan over tensor ids [0, n_leaves) + intermediates
void plan_from_order(const Net& net, const std::vector<int>& order, std::vector<PlanStep>& plan,
                     std::vector<std::vector<int>>& t_ids, int& root) {
    const int nl = (int)net.ids.size();
    t_ids.assign(net.ids.begin(), net.ids.end());
    std::vector<char> alive(nl, 1);
    std::map<int, std::set<int>> owners;
    for (int t = 0; t < nl; ++t) for (int i : t_ids[t]) owners[i].insert(t);
    auto size = [&](const std::vector<int>& ids) { double r = 1; for (int i : ids) r *= net.dim[i]; return r; };
    auto result = [&](int a, int b) {
        std::set<int> res(t_ids[a].begin(), t_ids[a].end());
        res.insert(t_ids[b].begin(), t_ids[b].end());
        for (int i : t_ids[a]) {
            if (i == net.amp) continue;
            if (std::find(t_ids[b].begin(), t_ids[b].end(), i) == t_ids[b].end()) continue;
            const std::set<int>& ow = owners[i];
            bool others = false;
            for (int o : ow) if (o != a && o != b) { others = true; break; }
            if (!others) res.erase(i);
        }
        return std::vector<int>(res.begin(), res.end());
    };
    auto contract = [&](int a, int b) {
        std::vector<int> res = result(a, b);
        const int c = (int)t_ids.size();
        plan.push_back(PlanStep{a, b});
        for (int i : t_ids[a]) owners[i].erase(a);
        for (int i : t_ids[b]) owners[i].erase(b);
        for (int i : res) owners[i].insert(c);
        alive[a] = alive[b] = 0;
        alive.push_back(1);
        t_ids.push_back(std::move(res));
        return c;
    };
    std::vector<int> full(order);
    if (net.amp >= 0) full.push_back(net.amp);
    for (int ix : full) {
        std::vector<int> group(owners[ix].begin(), owners[ix].end());
        std::sort(group.begin(), group.end(), [&](int x, int y) {
            const double sx = size(t_ids[x]), sy = size(t_ids[y]);
            return sx != sy ? sx < sy : x < y;
        });
        while (group.size() > 1) {
            double best = -1; size_t bx = 0, by = 1;
            for (size_t x = 0; x < group.size(); ++x)
                for (size_t y = x + 1; y < group.size(); ++y) {
                    const double s = size(result(group[x], group[y]));
                    if (best < 0 || s < best) { best = s; bx = x; by = y; }
                }
            const int c = contract(group[bx], group[by]);
            std::vector<int> g2;
            for (size_t k = 0; k < group.size(); ++k) if (k != bx && k != by) g2.push_back(group[k]);
            g2.push_back(c);
            group.swap(g2);
        }
    }
    std::vector<int> rest;
    for (size_t t = 0; t < alive.size(); ++t) if (alive[t]) rest.push_back((int)t);
    while (rest.size() > 1) {                 // disconnected components: join the scalars
        const int c = contract(rest[0], rest[1]);
        rest.erase(rest.begin(), rest.begin() + 2);
        rest.insert(rest.begin(), c);
    }
    root = rest[0];
}


// ---- local refinement: tree rotations  (A.B).E -> (A.E).B  -------------------------------
// Y = ncon(D; X, E) with X = ncon(X; A, B) used only by Y.  If every label contracted at Y lives
// on A's side (none on B), E can be applied to A first: A' = ncon(A, E), D = ncon(A', B).  Exact
// (sum-product re-association); accepted only when the executor's cost model gets cheaper.
bool rotate(std::vector<Cmd>& cmds, size_t yi, int x_pos /*0: X is operand a, 1: operand b*/, bool toward_a,
            const std::map<std::string, int>& uses) {
    const Cmd& Y = cmds[yi];
    if (Y.kind != CMD_NCON) return false;
    const std::string& xn = x_pos == 0 ? Y.a : Y.b;
    const std::vector<int64_t>& xl = x_pos == 0 ? Y.al : Y.bl;
    const std::string& en = x_pos == 0 ? Y.b : Y.a;
    const std::vector<int64_t>& el = x_pos == 0 ? Y.bl : Y.al;
    size_t xi = cmds.size();
    for (size_t i = 0; i < yi; ++i) if (cmds[i].kind == CMD_NCON && cmds[i].name == xn) xi = i;
    if (xi == cmds.size()) return false;
    auto u = uses.find(xn);
    if (u == uses.end() || u->second != 1) return false;
    const Cmd& X = cmds[xi];
    if (X.cl.size() != xl.size()) return false;
    const std::string& an = toward_a ? X.a : X.b;
    const std::string& bn = toward_a ? X.b : X.a;
    const std::vector<int64_t>& al = toward_a ? X.al : X.bl;
    const std::vector<int64_t>& bl = toward_a ? X.bl : X.al;
    if (an == en || bn == en) return false;
    int64_t M = 1;
    for (int64_t l : Y.cl) M = std::max(M, l + 1);
    for (int64_t l : xl) M = std::max(M, l + 1);
    for (int64_t l : el) M = std::max(M, l + 1);
    std::map<int64_t, int64_t> tr;
    for (size_t j = 0; j < X.cl.size(); ++j) tr[X.cl[j]] = xl[j];
    auto T = [&](const std::vector<int64_t>& ls) {
        std::vector<int64_t> out;
        for (int64_t l : ls) { auto it = tr.find(l); out.push_back(it != tr.end() ? it->second : l + M); }
        return out;
    };
    const std::vector<int64_t> aU = T(al), bU = T(bl);
    auto in = [](const std::vector<int64_t>& v, int64_t x) { return std::find(v.begin(), v.end(), x) != v.end(); };
    for (int64_t l : xl)
        if (in(el, l) && !in(Y.cl, l) && in(bU, l)) return false;       // contracted at Y but also on B's side
    std::vector<int64_t> outp;
    for (int64_t l : aU) {
        const bool summed = in(el, l) && !in(Y.cl, l) && !in(bU, l);
        if (!summed) outp.push_back(l);
    }
    for (int64_t l : el)
        if (!in(aU, l) && (in(Y.cl, l) || in(bU, l)) && !in(outp, l)) outp.push_back(l);
    Cmd c1; c1.kind = CMD_NCON; c1.name = xn + "r"; c1.cl = outp; c1.a = an; c1.al = aU; c1.b = en; c1.bl = el;
    Cmd c2; c2.kind = CMD_NCON; c2.name = Y.name; c2.cl = Y.cl; c2.a = c1.name; c2.al = outp; c2.b = bn; c2.bl = bU;
    for (const Cmd& c : cmds) if (c.name == c1.name && c.kind != CMD_SAVE) return false;    // name clash
    std::vector<Cmd> out;
    for (size_t i = 0; i < cmds.size(); ++i) {
        if (i == xi) continue;
        if (i == yi) { out.push_back(c1); out.push_back(c2); continue; }
        out.push_back(cmds[i]);
    }
    cmds.swap(out);
    return true;
}

// number of batched (free) slice variables the plan is scored for; -1 = all (set by replan for its helpers)
thread_local int g_plan_n_free = -1;
uint64_t plan_mask(size_t n_vars) {
    const int k = (int)n_vars;
    return low_mask(g_plan_n_free < 0 || g_plan_n_free > k ? k : g_plan_n_free);
}

TreeCostModel cost_model_for(double elem_bytes) {
    TreeCostModel cm;
    cm.elem_bytes = elem_bytes;
    cm.bandwidth = 6.0e12;                                  // streaming kernels: 93 % of 6.45 TB/s (profiles/r1_summary.md)
    cm.flop_rate = elem_bytes <= 8 ? 40e12 : 27e12;         // GEMM kernels: c32 SIMT / 3xTF32 ~40, c64 DMMA 27 TFLOP/s
    cm.gemm_min_k_bits = elem_bytes <= 8 ? 4 : 3;           // gemm_kcb(): K chunk of the tensor-core kernels
    if (const char* e = getenv("QXB_MIN_LOB")) cm.thread_bits = std::min(8, std::max(5, atoi(e)));
    // A/B knobs: QXB_PLAN_L1_BW (TB/s; 0 = the single-rate model of profiles/r1p), QXB_PLAN_FLOP_RATE (TFLOP/s)
    if (const char* e = getenv("QXB_PLAN_FLOP_RATE")) if (atof(e) > 0) cm.flop_rate = atof(e) * 1e12;
    if (const char* e = getenv("QXB_PLAN_L1_BW")) cm.l1_bandwidth = atof(e) * 1e12;
    cm.shared_reread = getenv("QXB_PLAN_SHARED_REREAD") && atoi(getenv("QXB_PLAN_SHARED_REREAD")) != 0;   // experiment, off
    return cm;
}

// modelled seconds of a lowered program: per node max(bytes / bandwidth, flops / rate)
double lowered_cost_seconds(const Lowered& L, double n_amp, const TreeCostModel& cm) {
    double t = 0;
    for (const LOp& op : L.ops) {
        const double u = L.tensors[op.c].amp ? n_amp : 1;
        const bool rr = cm.shared_reread && L.tensors[op.c].amp;
        const double bytes = cm.elem_bytes * (op.elems_a * ((L.tensors[op.a].amp || rr) ? n_amp : 1) +
                                              op.elems_b * ((L.tensors[op.b].amp || rr) ? n_amp : 1) + op.elems_c * u);
        const double flops = 8.0 * op.macs_per_amp * u;
        (void)flops;
        const double compute = cm.compute_seconds(std::log2(std::max(1.0, op.macs_per_amp * u)), op.nC, op.n_m, op.n_n, op.nK);
        t += (op.phase == PH_CONST ? cm.const_weight : 1.0) * (std::max(bytes / cm.bandwidth, compute) + cm.launch_s);
    }
    return t;
}

double program_seconds(const std::vector<Cmd>& cmds, double n_amp, bool early_sum, const TreeCostModel& cm) {
    Program cand;
    cand.cmds = cmds;
    analyse(cand);
    Lowered L = lower(cand, plan_mask(cand.vars.size()), early_sum);
    return lowered_cost_seconds(L, n_amp, cm);
}

double program_cost(const std::vector<Cmd>& cmds, double n_amp, bool early_sum, double elem_bytes) {
    Program cand;
    cand.cmds = cmds;
    analyse(cand);
    Lowered L = lower(cand, plan_mask(cand.vars.size()), early_sum);
    return lowered_cost_bytes(L, n_amp, elem_bytes);
}

// Greedy descent over rotations of the largest nodes; returns the improved cost.
double refine_by_rotations(std::vector<Cmd>& cmds, double cost, double n_amp, bool early_sum, double elem_bytes,
                           int max_accept) {
    for (int acc = 0; acc < max_accept; ++acc) {
        Program cur;
        cur.cmds = cmds;
        analyse(cur);
        Lowered L = lower(cur, plan_mask(cur.vars.size()), early_sum);
        std::map<std::string, int> uses;
        for (const Cmd& c : cmds)
            if (c.kind == CMD_NCON) { uses[c.a]++; uses[c.b]++; } else if (c.kind == CMD_VIEW || c.kind == CMD_SAVE) uses[c.a]++;
        // candidates: the nodes that move the most bytes
        std::vector<std::pair<double, std::string>> big;
        for (const LOp& op : L.ops) {
            if (op.phase == PH_CONST) continue;
            const double w = op.elems_c * (L.tensors[op.c].amp ? n_amp : 1) + op.elems_a * (L.tensors[op.a].amp ? n_amp : 1) +
                             op.elems_b * (L.tensors[op.b].amp ? n_amp : 1);
            big.push_back({-w, op.name});
        }
        std::sort(big.begin(), big.end());
        if (big.size() > 48) big.resize(48);
        bool improved = false;
        for (auto& cand : big) {
            size_t yi = cmds.size();
            for (size_t i = 0; i < cmds.size(); ++i) if (cmds[i].kind == CMD_NCON && cmds[i].name == cand.second) yi = i;
            if (yi == cmds.size()) continue;
            for (int variant = 0; variant < 4 && !improved; ++variant) {
                std::vector<Cmd> trial = cmds;
                if (!rotate(trial, yi, variant & 1, (variant & 2) == 0, uses)) continue;
                double c;
                try { c = program_cost(trial, n_amp, early_sum, elem_bytes); } catch (const Error&) { continue; }
                if (c < cost * 0.998) { cmds.swap(trial); cost = c; improved = true; }
            }
            if (improved) break;
        }
        if (!improved) break;
    }
    return cost;
}

}  // namespace

bool replan(Program& prog, int candidates, uint64_t seed, double n_amp, bool early_sum, double* given_bytes,
            double* new_bytes, double elem_bytes, int n_free, double budget_bytes, int* n_free_out, double* seconds_out) {
    if (!prog.analysed) analyse(prog);
    const int n_vars = (int)prog.vars.size();
    const bool auto_free = n_free == -2;
    const bool autoslice = n_free == -3;
    g_plan_n_free = (auto_free || autoslice) ? -1 : n_free;
    struct Reset { ~Reset() { g_plan_n_free = -1; } } reset_guard;
    if (n_free_out) *n_free_out = g_plan_n_free < 0 ? n_vars : std::min(g_plan_n_free, n_vars);
    if (seconds_out) *seconds_out = INFINITY;
    const TreeCostModel cm = cost_model_for(elem_bytes);
    double base = INFINITY, base_time = INFINITY;      // a given plan the lowering rejects (tensor too large) costs infinity
    try {
        const Lowered L0 = lower(prog, plan_mask(prog.vars.size()), early_sum);
        base = lowered_cost_bytes(L0, n_amp, elem_bytes);
        base_time = lowered_cost_seconds(L0, n_amp, cm);
    } catch (const Error&) {
    }
    if (given_bytes) *given_bytes = base;
    if (new_bytes) *new_bytes = base;

    // ---- recover the network: classes of modes identified through ncon labels
    UF uf;
    std::vector<std::vector<int>> h(prog.defs.size());
    std::vector<char> is_operand(prog.defs.size(), 0);
    for (size_t di = 0; di < prog.defs.size(); ++di) {
        const TensorDef& d = prog.defs[di];
        if (d.kind == T_LOAD || d.kind == T_OUTPUT) {
            for (size_t m = 0; m < d.modes.size(); ++m) h[di].push_back(uf.make());
        } else if (d.kind == T_VIEW) {
            // the view's target is the previous link of the chain: same modes, same handles
            int tgt = -1;
            for (const Cmd& c : prog.cmds)
                if (c.kind == CMD_VIEW && c.name == d.name) tgt = prog.by_name.at(c.a);
            h[di] = h[tgt];
        } else {
            is_operand[d.a] = is_operand[d.b] = 1;
            std::map<int64_t, int> by;
            auto scan = [&](const std::vector<int64_t>& ls, int t) {
                for (size_t i = 0; i < ls.size(); ++i) {
                    auto it = by.find(ls[i]);
                    if (it == by.end()) by[ls[i]] = h[t][i]; else uf.unite(it->second, h[t][i]);
                }
            };
            scan(d.al, d.a); scan(d.bl, d.b);
            for (int64_t l : d.cl) h[di].push_back(by.at(l));
        }
    }
    if (prog.defs[prog.root].kind != T_NCON || !prog.defs[prog.root].modes.empty()) return false;
    for (const TensorDef& d : prog.defs) if (d.kind == T_NCON && d.uses > 1) return false;   // needs a tree

    Net net;
    std::map<int, int> cls;                    // uf root -> dense class id
    auto cls_of = [&](int handle) {
        const int r = uf.find(handle);
        auto it = cls.find(r);
        if (it != cls.end()) return it->second;
        const int id = (int)cls.size();
        cls[r] = id;
        return id;
    };
    std::vector<int> leaf_defs;
    for (size_t di = 0; di < prog.defs.size(); ++di) {
        const TensorDef& d = prog.defs[di];
        if (d.kind == T_NCON || !is_operand[di]) continue;
        leaf_defs.push_back((int)di);
        std::vector<int> modes, ids;
        for (size_t m = 0; m < d.modes.size(); ++m) {
            const int c = cls_of(h[di][m]);
            if (std::find(modes.begin(), modes.end(), c) != modes.end()) return false;   // repeated index in one tensor
            modes.push_back(c);
        }
        ids = modes;
        net.names.push_back(d.name);
        net.modes.push_back(modes);
        net.ids.push_back(ids);
    }
    const int ncls = (int)cls.size();
    net.dim.assign(ncls + 1, 1.0);
    for (size_t li = 0; li < leaf_defs.size(); ++li) {
        const TensorDef& d = prog.defs[leaf_defs[li]];
        for (size_t m = 0; m < d.modes.size(); ++m) net.dim[net.modes[li][m]] = (double)d.modes[m].full_ext;
    }
    std::vector<int> cls_var(ncls + 1, -1);            // slice variable a class is viewed with (-1: none)
    for (size_t li = 0; li < leaf_defs.size(); ++li) {
        const TensorDef& d = prog.defs[leaf_defs[li]];
        for (size_t m = 0; m < d.modes.size(); ++m) if (d.modes[m].var >= 0) cls_var[net.modes[li][m]] = d.modes[m].var;
    }
    bool any_out = false;
    for (size_t li = 0; li < leaf_defs.size(); ++li)
        if (prog.defs[prog.defs[leaf_defs[li]].leaf].kind == T_OUTPUT) { net.ids[li].push_back(ncls); any_out = true; }
    net.amp = any_out ? ncls : -1;
    net.dim[ncls] = std::max(2.0, n_amp);
    if (leaf_defs.size() < 3) return false;

    std::vector<std::set<int>> lg(ncls);
    for (const auto& modes : net.modes)
        for (int a : modes) for (int b : modes) if (a != b) lg[a].insert(b);

    // leaf statements verbatim, in program order
    std::vector<Cmd> leaf_cmds;
    std::set<std::string> taken;
    for (const Cmd& c : prog.cmds)
        if (c.kind == CMD_LOAD || c.kind == CMD_OUTPUT || c.kind == CMD_VIEW) { leaf_cmds.push_back(c); taken.insert(c.name); }
    std::string prefix = "R";
    while (true) {
        bool clash = false;
        for (const std::string& s : taken) if (s.rfind(prefix, 0) == 0) { clash = true; break; }
        if (!clash) break;
        prefix += "_";
    }
    std::string save_label = "output";
    for (const Cmd& c : prog.cmds) if (c.kind == CMD_SAVE) save_label = c.name;

    // ---- emit the ncon statements of a pairwise plan (labels like build_compute_graph assigns them)
    auto emit = [&](const std::vector<PlanStep>& plan, int root) {
        std::vector<std::string> names(net.names);
        std::vector<std::vector<int>> cur(net.modes);          // classes per mode of every tensor (DSL order)
        std::map<int, int> cnt;
        for (const auto& m : net.modes) for (int i : m) cnt[i]++;
        std::vector<Cmd> cmds(leaf_cmds);
        for (size_t s = 0; s < plan.size(); ++s) {
            const int a = plan[s].a, b = plan[s].b;
            const std::vector<int>&ia = cur[a], &ib = cur[b];
            std::map<int, int64_t> label;
            for (int i : ia) if (!label.count(i)) label[i] = (int64_t)label.size() + 1;
            for (int i : ib) if (!label.count(i)) label[i] = (int64_t)label.size() + 1;
            auto in = [](const std::vector<int>& v, int x) { return std::find(v.begin(), v.end(), x) != v.end(); };
            auto others = [&](int i) { return cnt[i] - (in(ia, i) ? 1 : 0) - (in(ib, i) ? 1 : 0); };
            std::vector<int> keep;
            for (int i : ia) if (others(i) > 0) keep.push_back(i);
            for (int i : ib) if (!in(keep, i) && others(i) > 0) keep.push_back(i);
            Cmd c; c.kind = CMD_NCON;
            c.name = prefix + std::to_string(s + 1);
            c.a = names[a]; c.b = names[b];
            for (int i : keep) c.cl.push_back(label[i]);
            for (int i : ia) c.al.push_back(label[i]);
            for (int i : ib) c.bl.push_back(label[i]);
            for (int i : ia) cnt[i]--;
            for (int i : ib) cnt[i]--;
            for (int i : keep) cnt[i]++;
            names.push_back(c.name);
            cur.push_back(keep);
            cmds.push_back(std::move(c));
        }
        Cmd sv; sv.kind = CMD_SAVE; sv.name = save_label; sv.a = names[root];
        cmds.push_back(sv);
        return cmds;
    };
    // ---- score with the executor's own lowering: modelled seconds (bytes for the streaming nodes, flops for
    //      GEMM-shaped ones); a candidate the lowering rejects is simply skipped
    double best = base_time, best_b = base;
    std::vector<Cmd> best_cmds;
    auto consider = [&](std::vector<Cmd>&& cmds) {
        Program cand;
        cand.cmds = cmds;
        try {
            analyse(cand);
            Lowered L = lower(cand, plan_mask(cand.vars.size()), early_sum);
            const double c = lowered_cost_seconds(L, n_amp, cm);
            if (c < best) { best = c; best_b = lowered_cost_bytes(L, n_amp, elem_bytes); best_cmds = std::move(cmds); return true; }
        } catch (const Error&) {
        }
        return false;
    };

    std::mt19937_64 rng(seed);
    std::vector<PlanStep> best_mf_plan; int best_mf_root = -1;
    const int n_mf = std::max(1, std::min(candidates, 32));
    for (int it = 0; it < n_mf; ++it) {
        std::vector<int> order = min_fill_order(lg, it == 0 ? nullptr : &rng);
        std::vector<PlanStep> plan;
        std::vector<std::vector<int>> t_ids;
        int root = -1;
        plan_from_order(net, order, plan, t_ids, root);
        if (consider(emit(plan, root))) { best_mf_plan = plan; best_mf_root = root; }
    }
    // ---- tree search on the network (greedy restarts + subtree reconfiguration), seeded with the best order-based tree
    auto make_tn = [&](int nf) {
        TreeNet tn;
        tn.ncls = ncls + 1;
        tn.amp = net.amp;
        tn.wbits.assign(ncls + 1, 0.0);
        for (int c = 0; c < ncls; ++c) {
            const bool fixed = cls_var[c] >= 0 && nf >= 0 && cls_var[c] >= nf;          // variables >= nf are fixed by the caller
            tn.wbits[c] = fixed ? 0.0 : std::ceil(std::log2(std::max(1.0, net.dim[c])) - 1e-9);
        }
        tn.wbits[ncls] = std::log2(std::max(2.0, n_amp));
        tn.total.assign(ncls + 1, 0);
        tn.leaf_ids = net.ids;
        for (const auto& ids : net.ids) for (int i : ids) tn.total[i]++;
        for (size_t li = 0; li < leaf_defs.size(); ++li) tn.leaf_var.push_back(prog.defs[leaf_defs[li]].vars.empty() ? 0 : 1);
        return tn;
    };
    TreeCostModel cm_search = cm;                          // the search may use extra tree builders
    if (autoslice) cm_search.bisection_restarts = 16;      // GPU-aware slicing: also seed with recursive-bisection trees
    if (getenv("QXB_TREEOPT_BISECT")) cm_search.bisection_restarts = atoi(getenv("QXB_TREEOPT_BISECT"));
    if (getenv("QXB_TREEOPT_PROBE")) {                     // planner experiments: search only, print the report
        const int rounds = atoi(getenv("QXB_TREEOPT_PROBE"));
        TreeNet tn = make_tn(g_plan_n_free);
        std::vector<std::pair<int, int>> tplan; int troot = -1; TreeReport rep;
        optimize_tree(tn, cm_search, std::max(1, candidates), rounds, seed, {}, {}, tplan, troot, &rep);
        fprintf(stderr, "[treeopt] restarts %d rounds %d n_free %d: %.4g s (2^%.2f flops, %.4g GB, largest node 2^%.0f)\n",
                candidates, rounds, g_plan_n_free, rep.seconds, std::log2(std::max(rep.flops, 1.0)), rep.bytes / 1e9, rep.max_bits);
        return false;
    }
    if (autoslice) {
        // GPU-aware slicing: search a tree for the network as it is, then add slice variables
        // (views on every leaf of the chosen index classes, like build_compute_graph does for a bond group,
        // compute_graph.jl:39-58) until the largest tensor fits the budget.  Exact: a sliced index is summed by
        // the executor's slice loop instead of inside a contraction.
        TreeNet tn = make_tn(-1);
        std::vector<std::vector<std::pair<int, int>>> seeds;
        std::vector<int> seed_roots;
        if (best_mf_root >= 0) {
            std::vector<std::pair<int, int>> sp;
            for (const PlanStep& st : best_mf_plan) sp.push_back({st.a, st.b});
            seeds.push_back(sp); seed_roots.push_back(best_mf_root);
        }
        std::vector<std::pair<int, int>> tplan;
        int troot = -1;
        TreeReport rep, rep2;
        optimize_tree(tn, cm_search, std::max(4, candidates), 32, seed ^ 0xD1B54A32D192ED03ull, seeds, seed_roots, tplan, troot, &rep);
        const double max_node_bits = std::floor(std::log2(std::max(budget_bytes, 3.0 * elem_bytes) / (3.0 * elem_bytes)));
        std::vector<char> sliceable(ncls + 1, 0);
        for (int c = 0; c < ncls; ++c) sliceable[c] = cls_var[c] < 0;
        const std::vector<int> chosen = slice_tree(tn, cm, tplan, troot, max_node_bits, sliceable, 48, seed, &rep2);
        if (getenv("QXB_REPLAN_VERBOSE")) {
            double blocks = 1;
            for (int c : chosen) blocks *= net.dim[c];
            fprintf(stderr, "[autoslice] unsliced tree %.3g s (2^%.1f flops, largest 2^%.0f); %zu indices sliced (%.3g slices): "
                            "%.3g s per slice, largest 2^%.0f -> %.3g s\n", rep.seconds, std::log2(std::max(rep.flops, 1.0)),
                    rep.max_bits, chosen.size(), blocks, rep2.seconds, rep2.max_bits, rep2.seconds * blocks);
        }
        if (rep2.max_bits > max_node_bits) return false;
        for (size_t j = 0; j < chosen.size(); ++j) {
            const int c = chosen[j];
            const std::string sym = "v" + std::to_string(n_vars + (int)j + 1);
            for (size_t li = 0; li < net.modes.size(); ++li) {
                for (size_t m = 0; m < net.modes[li].size(); ++m) {
                    if (net.modes[li][m] != c) continue;
                    Cmd v; v.kind = CMD_VIEW;
                    v.name = net.names[li] + "_s";
                    while (taken.count(v.name)) v.name += "_s";
                    taken.insert(v.name);
                    v.a = net.names[li]; v.label = sym; v.idx = (int64_t)m + 1; v.dim = (int64_t)net.dim[c];
                    leaf_cmds.push_back(v);
                    net.names[li] = v.name;
                }
            }
        }
        std::vector<PlanStep> plan;
        for (auto& st : tplan) plan.push_back(PlanStep{st.first, st.second});
        std::vector<Cmd> cmds = emit(plan, troot);
        Program cand;
        cand.cmds = cmds;
        analyse(cand);
        const Lowered L = lower(cand, low_mask(n_vars), early_sum);      // the new variables fixed (one slice per block)
        if (seconds_out) *seconds_out = lowered_cost_seconds(L, n_amp, cm);
        if (new_bytes) *new_bytes = lowered_cost_bytes(L, n_amp, elem_bytes);
        if (n_free_out) *n_free_out = n_vars;
        prog.cmds = std::move(cmds);
        prog.analysed = false;
        analyse(prog);
        return true;
    }
    if (auto_free) {
        // how many variables to batch: the count that minimises (blocks x modelled seconds of the best quick tree)
        // among those whose largest node (three operands live) fits the budget
        double best_total = INFINITY;
        int best_nf = -1;
        for (int nf = n_vars; nf >= 0; --nf) {
            TreeNet tn = make_tn(nf);
            std::vector<std::pair<int, int>> tplan; int troot = -1; TreeReport rep;
            optimize_tree(tn, cm, 6, 6, seed ^ (uint64_t)nf, {}, {}, tplan, troot, &rep);
            double blocks = 1;
            for (int v = nf; v < n_vars; ++v) blocks *= (double)prog.vars[v].dim;
            const bool fits = budget_bytes <= 0 || 3.0 * elem_bytes * std::exp2(rep.max_bits) <= budget_bytes;
            if (getenv("QXB_REPLAN_VERBOSE"))
                fprintf(stderr, "[replan] n_free %d: blocks %.3g, quick tree %.3g s/block (2^%.1f flops, %.3g GB, largest node 2^%.0f) -> %.3g s%s\n",
                        nf, blocks, rep.seconds, std::log2(std::max(rep.flops, 1.0)), rep.bytes / 1e9, rep.max_bits,
                        rep.seconds * blocks, fits ? "" : "  (over budget)");
            if (fits && rep.seconds * blocks < best_total) { best_total = rep.seconds * blocks; best_nf = nf; }
            if (fits && rep.seconds * blocks > 4.0 * best_total) break;                  // further fixing only multiplies the work
        }
        if (best_nf < 0) return false;
        g_plan_n_free = best_nf;
        if (n_free_out) *n_free_out = best_nf;
        best = INFINITY; best_b = INFINITY; best_cmds.clear();                          // re-score everything for this mask
        try {
            const Lowered L0 = lower(prog, plan_mask(prog.vars.size()), early_sum);
            best = lowered_cost_seconds(L0, n_amp, cm); best_b = lowered_cost_bytes(L0, n_amp, elem_bytes);
            base_time = best;
        } catch (const Error&) { base_time = INFINITY; }
    }
    {
        TreeNet tn = make_tn(g_plan_n_free);
        std::vector<std::vector<std::pair<int, int>>> seeds;
        std::vector<int> seed_roots;
        if (best_mf_root >= 0) {
            std::vector<std::pair<int, int>> sp;
            for (const PlanStep& st : best_mf_plan) sp.push_back({st.a, st.b});
            seeds.push_back(sp); seed_roots.push_back(best_mf_root);
        }
        std::vector<std::pair<int, int>> tplan;
        int troot = -1;
        TreeReport rep;
        optimize_tree(tn, cm, std::max(4, candidates), 24, seed ^ 0xD1B54A32D192ED03ull, seeds, seed_roots, tplan, troot, &rep);
        std::vector<PlanStep> plan;
        for (auto& st : tplan) plan.push_back(PlanStep{st.first, st.second});
        consider(emit(plan, troot));
    }
    if (best_cmds.empty()) {                               // nothing cheaper among the candidates: refine the given tree
        if (!std::isfinite(base)) return false;
        best_cmds = prog.cmds;
    }
    try {
        best_b = refine_by_rotations(best_cmds, best_b, n_amp, early_sum, elem_bytes, 48);
        best = std::min(best, program_seconds(best_cmds, n_amp, early_sum, cm));
    } catch (const Error&) {
    }
    if (!(best < base_time)) return false;
    prog.cmds = std::move(best_cmds);
    prog.analysed = false;
    analyse(prog);
    if (new_bytes) *new_bytes = best_b;
    if (seconds_out) *seconds_out = best;
    return true;
}

std::string program_text(const Program& p) {
    std::ostringstream o;
    o << "# version: 0.4.0\n";
    auto lab = [](const std::vector<int64_t>& v) {
        if (v.empty()) return std::string("0");
        std::string s;
        for (size_t i = 0; i < v.size(); ++i) s += (i ? "," : "") + std::to_string(v[i]);
        return s;
    };
    for (const Cmd& c : p.cmds) {
        switch (c.kind) {
        case CMD_LOAD: o << "load " << c.name << " " << c.label << " " << lab(c.dims) << "\n"; break;
        case CMD_OUTPUT: o << "output " << c.name << " " << c.idx << " " << c.dim << "\n"; break;
        case CMD_VIEW: o << "view " << c.name << " " << c.a << " " << c.label << " " << c.idx << " " << c.dim << "\n"; break;
        case CMD_NCON: o << "ncon " << c.name << " " << lab(c.cl) << " " << c.a << " " << lab(c.al) << " " << c.b << " " << lab(c.bl) << "\n"; break;
        case CMD_SAVE: o << "save " << c.name << " " << c.a << "\n"; break;
        }
    }
    return o.str();
}

}  // namespace qxb
