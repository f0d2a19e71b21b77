// qxb200 -- contraction-tree search on the recovered tensor network (planner-side lowering).
//
// The reference plans with an elimination order of the line graph (FlowCutter / min-fill,
// /root/reference/src/contraction_planning.jl:66-176) and turns it into pairwise steps
// (:386-448).  That is enough for CZ grids; fSim (Sycamore-like) networks and the batched
// execution model of this executor (bitstrings and slice variables are extra hyper-indices
// carried by every node above the leaves that own them) need a search over TREES scored by
// what the GPU actually pays.  This file holds
//   * a set-based cost model of a tree: per node  max(bytes / HBM bandwidth, flops / FMA rate),
//     bytes = es * (|A| + |B| + |C|), flops = 8 * 2^(bits of all indices of A and B); nodes that
//     depend on neither a bitstring nor a slice variable are folded at compile time (free);
//   * randomised greedy agglomeration (score = |C| - alpha (|A| + |B|), Boltzmann noise);
//   * subtree reconfiguration: a connected region of <= L frontier tensors is re-associated
//     optimally (dynamic programme over subsets) and swapped in when cheaper.
// The result is a list of pairwise steps; qxb_replan.cpp emits the ncon statements and scores the
// finished program with the executor's exact lowering before accepting it.
#include <algorithm>
#include <array>
#include <cmath>
#include <functional>
#include <cstdint>
#include <cstdlib>
#include <map>
#include <queue>
#include <random>
#include <set>
#include <vector>

#include "qxb_treeopt.h"

namespace qxb {

namespace {

using IdCnt = std::vector<std::pair<int, int>>;     // sorted (class, owners inside the subtree)

struct TNode {
    int l = -1, r = -1, parent = -1;
    IdCnt ids;                 // open classes with the number of leaves below that carry them
    bool amp = false;          // depends on the bitstrings
    bool var = false;          // depends on a slice variable
    double bits = 0;           // log2 of the stored elements
    double cost = 0;           // modelled seconds of the contraction producing this node (0 for leaves)
};

struct Tree {
    const TreeNet* net = nullptr;
    const TreeCostModel* cm = nullptr;
    std::vector<TNode> n;      // [0, nl) leaves, then internal nodes
    int root = -1;

    double bits_of(const IdCnt& ids) const {
        double b = 0;
        for (auto& ic : ids) b += net->wbits[ic.first];
        return b;
    }
    // merge two children: classes whose every owner is inside the union are summed (closed)
    void merge(const TNode& a, const TNode& b, TNode& c, double* union_bits) const {
        c.ids.clear();
        size_t i = 0, j = 0;
        double ub = 0;
        auto put = [&](int cls, int cnt) {
            ub += net->wbits[cls];
            if (cnt < net->total[cls] || cls == net->amp) c.ids.push_back({cls, cnt});
        };
        while (i < a.ids.size() || j < b.ids.size()) {
            if (j == b.ids.size() || (i < a.ids.size() && a.ids[i].first < b.ids[j].first)) { put(a.ids[i].first, a.ids[i].second); ++i; }
            else if (i == a.ids.size() || b.ids[j].first < a.ids[i].first) { put(b.ids[j].first, b.ids[j].second); ++j; }
            else { put(a.ids[i].first, a.ids[i].second + b.ids[j].second); ++i; ++j; }
        }
        c.amp = a.amp || b.amp;
        c.var = a.var || b.var;
        c.bits = bits_of(c.ids);
        if (union_bits) *union_bits = ub;
    }
    double node_cost(const TNode& a, const TNode& b, const TNode& c, double union_bits) const {
        // nodes without bitstring / slice-variable dependence are folded once at compile time; they are charged
        // const_weight of their cost so the search cannot hide unbounded work (or memory) in that phase
        const double w = (!c.amp && !c.var) ? cm->const_weight : 1.0;
        // shared_reread: an operand without the bitstring axis is streamed through L1/L2 once per bitstring row by
        // the kernels (ncu: R260 of the 7x7 program, L1 at 96 %), so it costs like a per-row operand
        const double ab = net->amp >= 0 ? net->wbits[net->amp] : 0.0;
        const double ea = (cm->shared_reread && c.amp && !a.amp) ? a.bits + ab : a.bits;
        const double eb = (cm->shared_reread && c.amp && !b.amp) ? b.bits + ab : b.bits;
        return w * cm->time(ea, eb, c.bits, union_bits, a.amp ? ab : 0.0, b.amp ? ab : 0.0);
    }
    void recompute(int v) {
        TNode& c = n[v];
        double ub = 0;
        merge(n[c.l], n[c.r], c, &ub);
        c.cost = node_cost(n[c.l], n[c.r], c, ub);
    }
    double total() const {
        double t = 0;
        for (const TNode& x : n) t += x.cost;
        return t;
    }
    double max_bits() const {
        double m = 0;
        for (const TNode& x : n) m = std::max(m, x.bits);
        return m;
    }
    void init_leaves() {
        const int nl = (int)net->leaf_ids.size();
        n.assign(nl, TNode{});
        for (int t = 0; t < nl; ++t) {
            std::vector<int> ids(net->leaf_ids[t]);
            std::sort(ids.begin(), ids.end());
            for (int c : ids) n[t].ids.push_back({c, 1});
            n[t].amp = net->amp >= 0 && std::binary_search(ids.begin(), ids.end(), net->amp);
            n[t].var = net->leaf_var[t] != 0;
            n[t].bits = bits_of(n[t].ids);
        }
    }
    void recompute_all() {                         // children before parents (slots are reused by reconfiguration)
        const int nl = (int)net->leaf_ids.size();
        for (int t = 0; t < nl; ++t) { n[t].bits = bits_of(n[t].ids); n[t].var = net->leaf_var[t] != 0; }
        std::vector<std::pair<int, int>> st{{root, 0}};
        while (!st.empty()) {
            auto [v, state] = st.back();
            st.pop_back();
            if (n[v].l < 0) continue;
            if (state == 0) { st.push_back({v, 1}); st.push_back({n[v].r, 0}); st.push_back({n[v].l, 0}); }
            else recompute(v);
        }
    }
    int add(int a, int b) {
        TNode c;
        c.l = a; c.r = b;
        n.push_back(c);
        const int v = (int)n.size() - 1;
        n[a].parent = v; n[b].parent = v;
        recompute(v);
        return v;
    }
};

// ------------------------------------------------------------------ greedy agglomeration
void greedy_build(Tree& T, double alpha, double temperature, std::mt19937_64& rng) {
    const TreeNet& net = *T.net;
    T.init_leaves();
    const int nl = (int)net.leaf_ids.size();
    std::vector<std::set<int>> owners(net.ncls);
    std::vector<char> alive(nl, 1);
    for (int t = 0; t < nl; ++t)
        for (auto& ic : T.n[t].ids) if (ic.first != net.amp) owners[ic.first].insert(t);
    struct Cand { double score; int a, b; bool operator<(const Cand& o) const { return score > o.score; } };
    std::priority_queue<Cand> pq;
    std::uniform_real_distribution<double> U(1e-12, 1.0);
    auto gumbel = [&]() { return temperature > 0 ? -temperature * std::log(-std::log(U(rng))) : 0.0; };
    auto push = [&](int a, int b) {
        TNode c;
        double ub;
        T.merge(T.n[a], T.n[b], c, &ub);
        // cotengra-style score in log space is unstable for mixed sizes; use sizes directly but compress with log2
        const double sc = std::exp2(c.bits) - alpha * (std::exp2(T.n[a].bits) + std::exp2(T.n[b].bits));
        const double s = (sc >= 0 ? std::log2(1.0 + sc) : -std::log2(1.0 - sc)) - gumbel();
        pq.push(Cand{s, a, b});
    };
    auto neighbours = [&](int a) {
        std::set<int> nb;
        for (auto& ic : T.n[a].ids) {
            if (ic.first == net.amp) continue;
            for (int o : owners[ic.first]) if (o != a) nb.insert(o);
        }
        return nb;
    };
    for (int a = 0; a < nl; ++a)
        for (int b : neighbours(a)) if (a < b) push(a, b);
    int remaining = nl;
    while (remaining > 1) {
        int a = -1, b = -1;
        while (!pq.empty()) {
            Cand c = pq.top(); pq.pop();
            if (alive[c.a] && alive[c.b]) { a = c.a; b = c.b; break; }
        }
        if (a < 0) {
            // disconnected (or connected through the bitstring axis only): join the two smallest
            std::vector<std::pair<double, int>> rest;
            for (size_t t = 0; t < alive.size(); ++t) if (alive[t]) rest.push_back({T.n[t].bits, (int)t});
            std::sort(rest.begin(), rest.end());
            a = rest[0].second; b = rest[1].second;
        }
        for (auto& ic : T.n[a].ids) if (ic.first != net.amp) owners[ic.first].erase(a);
        for (auto& ic : T.n[b].ids) if (ic.first != net.amp) owners[ic.first].erase(b);
        alive[a] = alive[b] = 0;
        const int c = T.add(a, b);
        alive.push_back(1);
        for (auto& ic : T.n[c].ids) if (ic.first != net.amp) owners[ic.first].insert(c);
        --remaining;
        for (int o : neighbours(c)) push(c, o);
    }
    for (size_t t = 0; t < alive.size(); ++t) if (alive[t]) T.root = (int)t;
}

// ------------------------------------------------------------------ recursive bisection
// Partition-based construction: split the leaves into two balanced halves with a small cut (sum of the bits of the
// index classes that have owners on both sides; Fiduccia-Mattheyses passes from a breadth-first seed partition),
// recurse, and contract the two halves last.  Lattice-like networks (Sycamore) get much narrower trees this way
// than from greedy agglomeration; small groups are finished greedily.
struct Bisector {
    Tree& T;
    std::mt19937_64& rng;
    double imbalance;                           // each side keeps at least (0.5 - imbalance) of the vertices
    const TreeNet& net;
    Bisector(Tree& t, std::mt19937_64& r, double imb) : T(t), rng(r), imbalance(imb), net(*t.net) {}

    int small_group(std::vector<int> nodes) {   // smallest-result-first agglomeration of a handful of tensors
        while (nodes.size() > 1) {
            double best = -1; size_t bi = 0, bj = 1;
            for (size_t i = 0; i < nodes.size(); ++i)
                for (size_t j = i + 1; j < nodes.size(); ++j) {
                    TNode c; double ub;
                    T.merge(T.n[nodes[i]], T.n[nodes[j]], c, &ub);
                    const double sc = std::exp2(c.bits) - std::exp2(T.n[nodes[i]].bits) - std::exp2(T.n[nodes[j]].bits);
                    if (best < 0 || sc < best) { best = sc; bi = i; bj = j; if (best < -1e300) break; }
                }
            const int c = T.add(nodes[bi], nodes[bj]);
            nodes.erase(nodes.begin() + bj); nodes.erase(nodes.begin() + bi);
            nodes.push_back(c);
        }
        return nodes[0];
    }

    int build(const std::vector<int>& leaves) {
        const int n = (int)leaves.size();
        if (n <= 6) return small_group(leaves);
        // local hypergraph: classes with >= 2 pins inside this group
        std::map<int, std::vector<int>> pins;                       // class -> local vertex ids
        for (int i = 0; i < n; ++i)
            for (int c : net.leaf_ids[leaves[i]]) if (c != net.amp && net.wbits[c] > 0) pins[c].push_back(i);
        std::vector<std::vector<int>> vcls(n);                       // per vertex: indices into `edges`
        std::vector<std::vector<int>> edges; std::vector<double> ew;
        for (auto& kv : pins) {
            if (kv.second.size() < 2) continue;
            for (int v : kv.second) vcls[v].push_back((int)edges.size());
            edges.push_back(kv.second); ew.push_back(net.wbits[kv.first]);
        }
        // seed partition: breadth-first growth from a random vertex until half of the vertices are taken
        std::vector<char> side(n, 1);
        {
            std::vector<char> seen(n, 0);
            std::vector<int> queue{(int)(rng() % n)};
            seen[queue[0]] = 1;
            size_t head = 0; int taken = 0;
            while (taken < n / 2) {
                if (head == queue.size()) {                          // disconnected: jump to an unseen vertex
                    for (int v = 0; v < n; ++v) if (!seen[v]) { queue.push_back(v); seen[v] = 1; break; }
                }
                const int v = queue[head++];
                side[v] = 0; ++taken;
                std::vector<int> nb;
                for (int e : vcls[v]) for (int u : edges[e]) if (!seen[u]) { seen[u] = 1; nb.push_back(u); }
                std::shuffle(nb.begin(), nb.end(), rng);
                queue.insert(queue.end(), nb.begin(), nb.end());
            }
        }
        const int min_side = std::max(1, (int)std::floor((0.5 - imbalance) * n));
        std::vector<std::array<int, 2>> cnt(edges.size());
        auto recount = [&]() {
            for (size_t e = 0; e < edges.size(); ++e) { cnt[e] = {0, 0}; for (int v : edges[e]) cnt[e][side[v]]++; }
        };
        auto gain = [&](int v) {
            const int s = side[v], t = 1 - s;
            double g = 0;
            for (int e : vcls[v]) {
                if (cnt[e][s] == 1 && cnt[e][t] > 0) g += ew[e];        // edge leaves the cut
                else if (cnt[e][t] == 0 && cnt[e][s] > 1) g -= ew[e];   // edge enters the cut
            }
            return g;
        };
        for (int pass = 0; pass < 6; ++pass) {                       // Fiduccia-Mattheyses passes
            recount();
            int n0 = 0;
            for (int v = 0; v < n; ++v) n0 += side[v] == 0;
            std::vector<char> locked(n, 0);
            std::vector<int> moved;
            double run = 0, best_run = 0; int best_len = 0;
            for (int step = 0; step < n; ++step) {
                int pick = -1; double pg = -1e300;
                for (int v = 0; v < n; ++v) {
                    if (locked[v]) continue;
                    const int from = side[v];
                    if ((from == 0 ? n0 : n - n0) - 1 < min_side) continue;
                    const double g = gain(v);
                    if (g > pg) { pg = g; pick = v; }
                }
                if (pick < 0) break;
                const int s = side[pick];
                for (int e : vcls[pick]) { cnt[e][s]--; cnt[e][1 - s]++; }
                side[pick] = (char)(1 - s);
                n0 += s == 0 ? -1 : 1;
                locked[pick] = 1; moved.push_back(pick);
                run += pg;
                if (run > best_run + 1e-9) { best_run = run; best_len = (int)moved.size(); }
                if ((int)moved.size() - best_len > 64) break;         // long unproductive tail
            }
            for (int i = (int)moved.size() - 1; i >= best_len; --i) side[moved[i]] = (char)(1 - side[moved[i]]);   // roll back
            if (best_len == 0) break;
        }
        std::vector<int> a, b;
        for (int v = 0; v < n; ++v) (side[v] == 0 ? a : b).push_back(leaves[v]);
        if (a.empty() || b.empty()) return small_group(leaves);
        const int ra = build(a), rb = build(b);
        return T.add(ra, rb);
    }
};

void bisect_build(Tree& T, double imbalance, std::mt19937_64& rng) {
    T.init_leaves();
    std::vector<int> all((int)T.net->leaf_ids.size());
    for (size_t i = 0; i < all.size(); ++i) all[i] = (int)i;
    Bisector bs(T, rng, imbalance);
    T.root = bs.build(all);
}

// ------------------------------------------------------------------ subtree reconfiguration
struct Bits {                      // up to 256 local classes
    uint64_t w[4] = {0, 0, 0, 0};
    void set(int i) { w[i >> 6] |= 1ull << (i & 63); }
};

bool reconfigure_at(Tree& T, int top, int L, std::mt19937_64* rng) {
    const TreeNet& net = *T.net;
    if (T.n[top].l < 0) return false;
    // grow the region: frontier starts as {children of top}; expand the largest internal frontier node
    std::vector<int> frontier{T.n[top].l, T.n[top].r};
    std::vector<int> inner{top};
    while ((int)frontier.size() < L) {
        int pick = -1; double best = -1;
        std::vector<int> internal;
        for (size_t i = 0; i < frontier.size(); ++i) {
            const TNode& f = T.n[frontier[i]];
            if (f.l < 0) continue;
            internal.push_back((int)i);
            if (f.bits > best) { best = f.bits; pick = (int)i; }
        }
        if (pick < 0) break;
        // stochastic sweeps grow the region through a random internal frontier node half of the time:
        // differently shaped regions reach re-associations the largest-first regions cannot
        if (rng && ((*rng)() & 1)) pick = internal[(size_t)((*rng)() % internal.size())];
        const int v = frontier[pick];
        inner.push_back(v);
        frontier[pick] = T.n[v].l;
        frontier.push_back(T.n[v].r);
    }
    const int F = (int)frontier.size();
    if (F < 3) return false;
    double old_cost = 0;
    for (int v : inner) old_cost += T.n[v].cost;
    // local classes
    std::map<int, int> loc;
    std::vector<int> cls;
    for (int f : frontier) for (auto& ic : T.n[f].ids) if (!loc.count(ic.first)) { loc[ic.first] = (int)cls.size(); cls.push_back(ic.first); }
    const int NC = (int)cls.size();
    if (NC > 256) return false;
    std::vector<uint32_t> own(NC, 0);
    std::vector<char> outside(NC, 0);
    {
        std::vector<int> inside(NC, 0);
        for (int i = 0; i < F; ++i)
            for (auto& ic : T.n[frontier[i]].ids) { own[loc[ic.first]] |= 1u << i; inside[loc[ic.first]] += ic.second; }
        for (int c = 0; c < NC; ++c) outside[c] = (cls[c] == net.amp) || inside[c] < net.total[cls[c]];
    }
    // weight planes (integer bits per class, amp handled separately as a flag)
    int maxw = 0;
    std::vector<int> wi(NC, 0);
    int amp_loc = -1;
    for (int c = 0; c < NC; ++c) {
        if (cls[c] == net.amp) { amp_loc = c; continue; }
        wi[c] = (int)std::lround(net.wbits[cls[c]]);
        maxw = std::max(maxw, wi[c]);
    }
    int planes = 0;
    while ((1 << planes) <= maxw) ++planes;
    std::vector<Bits> P(std::max(planes, 1));
    for (int c = 0; c < NC; ++c) for (int b = 0; b < planes; ++b) if ((wi[c] >> b) & 1) P[b].set(c);
    const double amp_bits = net.amp >= 0 ? net.wbits[net.amp] : 0.0;
    const uint32_t FULL = (1u << F) - 1u;
    std::vector<Bits> open(FULL + 1);
    std::vector<char> s_amp(FULL + 1, 0), s_var(FULL + 1, 0);
    std::vector<double> s_bits(FULL + 1, 0.0);
    auto wsum = [&](const Bits& x) {
        double s = 0;
        for (int b = 0; b < planes; ++b) {
            int pc = 0;
            for (int k = 0; k < 4; ++k) pc += __builtin_popcountll(x.w[k] & P[b].w[k]);
            s += (double)(pc << b);
        }
        return s;
    };
    for (uint32_t S = 1; S <= FULL; ++S) {
        Bits o;
        for (int c = 0; c < NC; ++c) {
            const uint32_t in = own[c] & S;
            if (in && ((own[c] & ~S) || outside[c])) o.set(c);
        }
        open[S] = o;
        bool a = false, v = false;
        for (int i = 0; i < F; ++i) if ((S >> i) & 1) { a |= T.n[frontier[i]].amp; v |= T.n[frontier[i]].var; }
        s_amp[S] = a; s_var[S] = v;
        const bool has_amp_open = amp_loc >= 0 && ((o.w[amp_loc >> 6] >> (amp_loc & 63)) & 1);
        s_bits[S] = wsum(o) + (has_amp_open ? amp_bits : 0.0);
    }
    // single frontier tensors keep their true stored size (identical, but be exact)
    for (int i = 0; i < F; ++i) s_bits[1u << i] = T.n[frontier[i]].bits;
    std::vector<double> best(FULL + 1, 0.0);
    std::vector<uint32_t> split(FULL + 1, 0);
    // subsets in increasing popcount order: iterate all S ascending works because proper subsets are smaller numbers
    for (uint32_t S = 1; S <= FULL; ++S) {
        if ((S & (S - 1)) == 0) { best[S] = 0; continue; }
        double bc = -1; uint32_t bs = 0;
        const uint32_t low = S & (~S + 1);                        // canonical: S1 contains the lowest member
        for (uint32_t S1 = (S - 1) & S; S1; S1 = (S1 - 1) & S) {
            if (!(S1 & low)) continue;
            const uint32_t S2 = S ^ S1;
            double c = best[S1] + best[S2];
            if (bc >= 0 && c >= bc) continue;
            {
                Bits u;
                for (int k = 0; k < 4; ++k) u.w[k] = open[S1].w[k] | open[S2].w[k];
                const bool ua = amp_loc >= 0 && ((u.w[amp_loc >> 6] >> (amp_loc & 63)) & 1);
                const double ub = wsum(u) + (ua ? amp_bits : 0.0);
                const double e1 = (T.cm->shared_reread && s_amp[S] && !s_amp[S1]) ? s_bits[S1] + amp_bits : s_bits[S1];
                const double e2 = (T.cm->shared_reread && s_amp[S] && !s_amp[S2]) ? s_bits[S2] + amp_bits : s_bits[S2];
                c += ((s_amp[S] || s_var[S]) ? 1.0 : T.cm->const_weight) * T.cm->time(e1, e2, s_bits[S], ub, s_amp[S1] ? amp_bits : 0.0, s_amp[S2] ? amp_bits : 0.0);
            }
            if (bc < 0 || c < bc) { bc = c; bs = S1; }
        }
        best[S] = bc; split[S] = bs;
    }
    if (!(best[FULL] < old_cost * (1.0 - 1e-9))) return false;
    // rebuild the region reusing the inner node slots (top stays the region's root)
    std::vector<int> slots(inner.begin() + 1, inner.end());
    const int parent_of_top = T.n[top].parent;
    std::vector<int> order;      // nodes to recompute, children first
    std::function<int(uint32_t, bool)> build = [&](uint32_t S, bool is_top) -> int {
        if ((S & (S - 1)) == 0) { int i = 0; while (!((S >> i) & 1)) ++i; return frontier[i]; }
        const int a = build(split[S], false), b = build(S ^ split[S], false);
        int v;
        if (is_top) v = top; else { v = slots.back(); slots.pop_back(); }
        T.n[v].l = a; T.n[v].r = b;
        T.n[a].parent = v; T.n[b].parent = v;
        order.push_back(v);
        return v;
    };
    build(FULL, true);
    T.n[top].parent = parent_of_top;
    for (int v : order) T.recompute(v);
    return true;
}

}  // namespace

double TreeCostModel::compute_seconds(double macs_bits, double c_bits, double m_bits, double n_bits, double k_bits) const {
    const double macs = std::exp2(macs_bits);
    if (l1_bandwidth <= 0 || (m_bits >= gemm_min_mn_bits && n_bits >= gemm_min_mn_bits && k_bits >= gemm_min_k_bits))
        return 8.0 * macs / flop_rate;
    // register tile as build_templates picks it: up to 2 M-only and 2 N-only bits, thread_bits of C stay thread bits
    int tm = (int)std::min(2.0, std::floor(m_bits + 1e-9)), tn = (int)std::min(2.0, std::floor(n_bits + 1e-9));
    const int room = std::max(0, (int)std::floor(c_bits + 1e-9) - thread_bits);
    while (tm + tn > room) { if (tn >= tm && tn > 0) --tn; else --tm; }
    const double loads_per_mac = (double)((1 << tm) + (1 << tn)) / (double)(1 << (tm + tn));
    return macs * loads_per_mac * elem_bytes / l1_bandwidth;
}

double TreeCostModel::time(double a_bits, double b_bits, double c_bits, double union_bits, double a_amp_bits, double b_amp_bits) const {
    const double bytes = elem_bytes * (std::exp2(a_bits) + std::exp2(b_bits) + std::exp2(c_bits));
    // a = batch + M + K, b = batch + N + K, c = batch + M + N, union = batch + M + N + K; the bitstring axis is a batch
    // axis when both operands carry it and otherwise rides with the operand that does -- it is never a tile axis, so
    // it is taken out of the per-row M / N / C bit counts
    const double c_amp = std::max(a_amp_bits, b_amp_bits);
    const double m_bits = (union_bits - b_bits) - (a_amp_bits > 0 && b_amp_bits == 0 ? a_amp_bits : 0.0);
    const double n_bits = (union_bits - a_bits) - (b_amp_bits > 0 && a_amp_bits == 0 ? b_amp_bits : 0.0);
    const double compute = compute_seconds(union_bits, c_bits - c_amp, m_bits, n_bits, union_bits - c_bits);
    return std::max(bytes / bandwidth, compute) + launch_s;
}

static void tree_to_plan(const Tree& T, std::vector<std::pair<int, int>>& plan, int& root) {
    const int nl = (int)T.net->leaf_ids.size();
    std::vector<int> newid(T.n.size(), -1);
    for (int t = 0; t < nl; ++t) newid[t] = t;
    plan.clear();
    // iterative post-order
    std::vector<std::pair<int, int>> st{{T.root, 0}};
    while (!st.empty()) {
        auto [v, state] = st.back();
        st.pop_back();
        if (T.n[v].l < 0) continue;
        if (state == 0) {
            st.push_back({v, 1});
            st.push_back({T.n[v].r, 0});
            st.push_back({T.n[v].l, 0});
        } else {
            newid[v] = nl + (int)plan.size();
            plan.push_back({newid[T.n[v].l], newid[T.n[v].r]});
        }
    }
    root = newid[T.root];
}

static void tree_from_plan(Tree& T, const std::vector<std::pair<int, int>>& plan, int root) {
    T.init_leaves();
    for (auto& s : plan) T.add(s.first, s.second);
    T.root = root;
}

static double refine(Tree& T, int rounds, int L, std::mt19937_64& rng) {
    const int nl = (int)T.net->leaf_ids.size();
    double cost = T.total();
    for (int r = 0; r < rounds; ++r) {
        // sweep over internal nodes from the most expensive down, plus random picks
        std::vector<std::pair<double, int>> byc;
        for (int v = nl; v < (int)T.n.size(); ++v) byc.push_back({-T.n[v].cost, v});
        std::sort(byc.begin(), byc.end());
        bool any = false;
        const size_t lim = std::min<size_t>(byc.size(), 96);
        for (size_t i = 0; i < lim; ++i)
            if (reconfigure_at(T, byc[i].second, L, nullptr)) any = true;
        for (int k = 0; k < 64 && byc.size() > lim; ++k) {
            const int v = byc[lim + (size_t)(rng() % (byc.size() - lim))].second;
            if (reconfigure_at(T, v, L, nullptr)) any = true;
        }
        const double c = T.total();
        if (!any || !(c < cost * (1.0 - 1e-6))) { cost = std::min(cost, c); break; }
        cost = c;
    }
    return cost;
}

// stochastic phase: random tops, randomly grown regions; every accepted move is a strict improvement
static double refine_random(Tree& T, int sweeps, int L, std::mt19937_64& rng) {
    const int nl = (int)T.net->leaf_ids.size();
    std::vector<int> internal;
    for (int v = nl; v < (int)T.n.size(); ++v) internal.push_back(v);
    for (int s = 0; s < sweeps; ++s) {
        std::shuffle(internal.begin(), internal.end(), rng);
        bool any = false;
        for (int v : internal) {
            if (T.n[v].cost <= 0 && (rng() & 3)) continue;           // mostly skip free (constant) nodes
            if (reconfigure_at(T, v, L, &rng)) any = true;
        }
        if (!any && s >= 2) break;
    }
    return T.total();
}

double optimize_tree(const TreeNet& net, const TreeCostModel& cm, int restarts, int refine_rounds, uint64_t seed,
                     const std::vector<std::vector<std::pair<int, int>>>& seeds_plans, const std::vector<int>& seeds_roots,
                     std::vector<std::pair<int, int>>& plan, int& root, TreeReport* report) {
    std::mt19937_64 rng(seed);
    const int L = getenv("QXB_TREEOPT_L") ? atoi(getenv("QXB_TREEOPT_L")) : 10;
    double best = -1;
    Tree bestT;
    auto consider = [&](Tree& T) {
        const double c = T.total();
        if (best < 0 || c < best) { best = c; bestT = T; }
    };
    // candidate pool: (cost, tree) -- keep the few best for the expensive refinement
    std::vector<std::pair<double, Tree>> pool;
    auto add_pool = [&](Tree&& T) {
        const double c = T.total();
        pool.push_back({c, std::move(T)});
        std::sort(pool.begin(), pool.end(), [](const auto& x, const auto& y) { return x.first < y.first; });
        if (pool.size() > 8) pool.pop_back();
    };
    for (size_t s = 0; s < seeds_plans.size(); ++s) {
        Tree T; T.net = &net; T.cm = &cm;
        tree_from_plan(T, seeds_plans[s], seeds_roots[s]);
        add_pool(std::move(T));
    }
    std::uniform_real_distribution<double> U(0.0, 1.0);
    for (int it = 0; it < restarts; ++it) {
        Tree T; T.net = &net; T.cm = &cm;
        const double alpha = it == 0 ? 1.0 : std::exp2(-3.0 + 4.0 * U(rng));           // 1/8 .. 2
        const double temp = it == 0 ? 0.0 : std::exp2(-6.0 + 6.0 * U(rng));            // 1/64 .. 1 (log2-size units)
        greedy_build(T, alpha, temp, rng);
        refine(T, 1, 8, rng);                                                          // one cheap sweep ranks candidates better
        add_pool(std::move(T));
    }
    if (cm.bisection_restarts > 0) {
        for (int it = 0; it < cm.bisection_restarts; ++it) {
            Tree T; T.net = &net; T.cm = &cm;
            bisect_build(T, 0.02 + 0.3 * U(rng), rng);
            refine(T, 1, 8, rng);
            add_pool(std::move(T));
        }
    }
    for (auto& pr : pool) {
        refine(pr.second, refine_rounds, L, rng);
        pr.first = pr.second.total();
    }
    std::sort(pool.begin(), pool.end(), [](const auto& x, const auto& y) { return x.first < y.first; });
    const int sweeps = getenv("QXB_TREEOPT_SWEEPS") ? atoi(getenv("QXB_TREEOPT_SWEEPS")) : 6;
    for (size_t i = 0; i < pool.size(); ++i) {
        if (i < 3 && sweeps > 0) {                                   // the best three also get the stochastic phase
            refine_random(pool[i].second, sweeps, L, rng);
            refine(pool[i].second, 4, L, rng);
        }
        consider(pool[i].second);
    }
    tree_to_plan(bestT, plan, root);
    if (report) {
        report->seconds = best;
        report->max_bits = bestT.max_bits();
        double fl = 0, by = 0;
        const int nl = (int)net.leaf_ids.size();
        for (int v = nl; v < (int)bestT.n.size(); ++v) {
            const TNode& c = bestT.n[v];
            if (!c.amp && !c.var) continue;
            TNode tmp; double ub;
            bestT.merge(bestT.n[c.l], bestT.n[c.r], tmp, &ub);
            fl += 8.0 * std::exp2(ub);
            by += cm.elem_bytes * (std::exp2(bestT.n[c.l].bits) + std::exp2(bestT.n[c.r].bits) + std::exp2(c.bits));
        }
        report->flops = fl; report->bytes = by;
    }
    return best;
}

std::vector<int> slice_tree(TreeNet net, const TreeCostModel& cm, std::vector<std::pair<int, int>>& plan, int& root,
                            double max_node_bits, const std::vector<char>& sliceable, int max_slices, uint64_t seed,
                            TreeReport* report) {
    std::mt19937_64 rng(seed);
    Tree T; T.net = &net; T.cm = &cm;
    tree_from_plan(T, plan, root);
    std::vector<int> chosen;
    const int nl = (int)net.leaf_ids.size();
    while ((int)chosen.size() < max_slices) {
        int big = -1;
        for (int v = 0; v < (int)T.n.size(); ++v) if (big < 0 || T.n[v].bits > T.n[big].bits) big = v;
        if (T.n[big].bits <= max_node_bits) break;
        // candidates: the classes of the largest node (and of its operands)
        std::set<int> cands;
        auto collect = [&](int v) { for (auto& ic : T.n[v].ids) if (ic.first != net.amp && net.wbits[ic.first] > 0 && sliceable[ic.first]) cands.insert(ic.first); };
        collect(big);
        if (T.n[big].l >= 0) { collect(T.n[big].l); collect(T.n[big].r); }
        if (cands.empty()) break;
        int pick = -1; double best = -1;
        for (int c : cands) {
            const double w = net.wbits[c];
            net.wbits[c] = 0;
            T.recompute_all();
            // work of all slices of this index; prefer indices that also shrink the largest node
            const double tot = T.total() * std::exp2(w) * (T.max_bits() < T.n[big].bits ? 1.0 : 4.0);
            net.wbits[c] = w;
            if (pick < 0 || tot < best) { pick = c; best = tot; }
        }
        net.wbits[pick] = 0;
        for (int t = 0; t < nl; ++t)
            if (std::find(net.leaf_ids[t].begin(), net.leaf_ids[t].end(), pick) != net.leaf_ids[t].end()) net.leaf_var[t] = 1;
        T.recompute_all();
        chosen.push_back(pick);
        refine(T, 2, 10, rng);
    }
    refine(T, 8, 10, rng);
    tree_to_plan(T, plan, root);
    if (report) { report->seconds = T.total(); report->max_bits = T.max_bits(); report->flops = 0; report->bytes = 0; }
    return chosen;
}

}  // namespace qxb
