// qxb200 -- host-side builder of row programs (qxb_rowprog.h): turns one phase of a lowered program into
// dependency levels, warp-sized work units, a size-aligned shared-memory arena plan and per-op descriptors.
// Pure host code (no CUDA calls): tests replay the descriptors on the CPU (tests/rowprog_emulator.py).
#include "qxb_rowplan.h"

#include <algorithm>
#include <cmath>
#include <cstring>
#include <map>

#include "../../include/qxb200.h"

namespace qxb {

namespace {

// merge (src bit, dst bit) pairs, sorted by src, into (src, dst, len) runs
int merge_runs(const std::vector<std::pair<int, int>>& bits, RSeg* out, int cap) {
    int n = 0;
    for (auto& b : bits) {
        if (n > 0 && out[n - 1].src + out[n - 1].len == b.first && out[n - 1].dst + out[n - 1].len == b.second) {
            out[n - 1].len = (unsigned char)(out[n - 1].len + 1);
        } else {
            if (n == cap) return -1;
            out[n++] = RSeg{(unsigned char)b.first, (unsigned char)b.second, 1, 0};
        }
    }
    return n;
}

// registers of a tile variant; must equal TileRegs<> in qxb_rowprog.cu
int tile_regs(int dtype, int ma, int nb, int kc) {
    const int rp = dtype == QXB_C32 ? 2 : 4;
    return rp * (((1 << ma) + (1 << nb)) * (1 << kc) + (1 << (ma + nb)));
}

struct Interval { int off, size, until; };     // live until level `until` inclusive

// lowest offset (multiple of `align`) where `size` elements overlap no live interval: first fit
int arena_alloc(std::vector<Interval>& live, int size, int until, int align = 1) {
    std::vector<std::pair<int, int>> iv;
    for (const Interval& x : live) iv.push_back({x.off, x.off + x.size});
    std::sort(iv.begin(), iv.end());
    int off = 0;
    for (auto& x : iv) {
        if (x.first - off >= size) break;
        off = std::max(off, (x.second + align - 1) / align * align);
    }
    live.push_back(Interval{off, size, until});
    return off;
}

}  // namespace

// The descriptor of one contraction for the unit interpreter: register tile, K chunk, tables, thread-tile maps.
// Tensor locations (oA / oB / oC, g*) are left to the caller.
// Layouts of the op's tensors inside the row arena: pa / pb / pc = address-bit permutations (old bit -> new bit) of A, B
// and C (nullptr: as lowered; C as lowered = identity).  free_operand (1 = A, 2 = B): that operand's layout is still to
// be chosen -- the bank model counts it conflict-free; group_out receives the C bits chosen as the low lane bits.
struct RowLayouts {
    const int* pa = nullptr;
    const int* pb = nullptr;
    const int* pc = nullptr;
    int free_operand = 0;
    std::vector<int>* group_out = nullptr;
};

bool describe_row_op(const LOp& op, int dtype, const RowPlanOptions& o, RowOp& d, int& n_units, double& unit_cost,
                     std::string& why, const RowLayouts* lay = nullptr) {
    auto bad = [&](const std::string& w) { why = w; return false; };
    memset(&d, 0, sizeof(d));
    d.lsA = d.lsB = d.lsC = kRowShared;
    const int nC = op.nC, nK = op.nK;
    if (nC > 16 || nK > 16) return bad("op too large for a row program");
    std::vector<int> mapA(nC, -1), mapB(nC, -1);
    for (auto& s : op.segA) for (int b = 0; b < s.len; ++b) mapA[s.src + b] = s.dst + b;
    for (auto& s : op.segB) for (int b = 0; b < s.len; ++b) mapB[s.src + b] = s.dst + b;
    std::vector<int> kposA(nK, -1), kposB(nK, -1);
    for (auto& s : op.segKA) for (int b = 0; b < s.len; ++b) kposA[s.src + b] = s.dst + b;
    for (auto& s : op.segKB) for (int b = 0; b < s.len; ++b) kposB[s.src + b] = s.dst + b;
    for (int b = 0; b < nC; ++b) if (mapA[b] > 15 || mapB[b] > 15) return bad("operand wider than 2^16 elements");
    for (int b = 0; b < nK; ++b) if (kposA[b] > 15 || kposB[b] > 15) return bad("operand wider than 2^16 elements");
    std::vector<int> mapC(nC);
    for (int b = 0; b < nC; ++b) mapC[b] = (lay && lay->pc) ? lay->pc[b] : b;
    if (lay && lay->pa) { for (int& x : mapA) if (x >= 0) x = lay->pa[x]; for (int& x : kposA) if (x >= 0) x = lay->pa[x]; }
    if (lay && lay->pb) { for (int& x : mapB) if (x >= 0) x = lay->pb[x]; for (int& x : kposB) if (x >= 0) x = lay->pb[x]; }
    const int free_operand = lay ? lay->free_operand : 0;
    // ComplexF64 with >= 3 M-only, >= 3 N-only and >= 2 K bits: 8 x 8 output tiles on the FP64 tensor pipe (kRowKindDmma)
    if (o.dmma && dtype == QXB_C64 && nK >= 2) {
        std::vector<int> ms, ns;
        for (int b = 0; b < nC; ++b) {
            if (mapA[b] >= 0 && mapB[b] < 0) ms.push_back(b);
            else if (mapB[b] >= 0 && mapA[b] < 0) ns.push_back(b);
        }
        if (ms.size() >= 3 && ns.size() >= 3) {
            // fragment index bits: the M-only (N-only) bits that sit lowest in A (B) -- nearby addresses within a fragment load
            std::sort(ms.begin(), ms.end(), [&](int x, int y) { return mapA[x] < mapA[y]; });
            std::sort(ns.begin(), ns.end(), [&](int x, int y) { return mapB[x] < mapB[y]; });
            const int mb[3] = {ms[0], ms[1], ms[2]}, nbq[3] = {ns[0], ns[1], ns[2]};
            RowOpHot& h = d.hot;
            h.kind = kRowKindDmma; h.nK = (uint8_t)nK; h.kc = 2; h.nb = 0; h.ks = 0;
            for (int lane = 0; lane < 32; ++lane) {
                const int g = lane >> 2, t = lane & 3;
                int a = 0, b = 0, c = 0;
                for (int j = 0; j < 3; ++j) if ((g >> j) & 1) { a |= 1 << mapA[mb[j]]; c |= 1 << mapC[mb[j]]; b |= 1 << mapB[nbq[j]]; }
                for (int j = 0; j < 2; ++j) if ((t >> j) & 1) { a |= 1 << kposA[j]; b |= 1 << kposB[j]; c |= 1 << mapC[nbq[j + 1]]; }
                d.frA[lane] = (uint16_t)a; d.frB[lane] = (uint16_t)b; d.frC[lane] = (uint16_t)c;
            }
            d.c_n1 = (uint16_t)(1 << mapC[nbq[0]]);
            for (int k = 0; k < (1 << std::min(nK, 4)); ++k) {
                int a = 0, b = 0;
                for (int t = 0; t < std::min(nK, 4); ++t) if ((k >> t) & 1) {
                    if (kposA[t] >= 0) a |= 1 << kposA[t];
                    if (kposB[t] >= 0) b |= 1 << kposB[t];
                }
                h.ktA[k] = (uint16_t)a; h.ktB[k] = (uint16_t)b;
            }
            // warp-tile index bits: every C bit that is not a fragment bit
            std::vector<bool> frag(nC, false);
            for (int j = 0; j < 3; ++j) { frag[mb[j]] = true; frag[nbq[j]] = true; }
            std::vector<std::pair<int, int>> ta, tb, tc, ka, kb;
            int t = 0;
            for (int b = 0; b < nC; ++b) {
                if (frag[b]) continue;
                tc.push_back({t, mapC[b]});
                if (mapA[b] >= 0) ta.push_back({t, mapA[b]});
                if (mapB[b] >= 0) tb.push_back({t, mapB[b]});
                ++t;
            }
            for (int b = 4; b < nK; ++b) {
                if (kposA[b] >= 0) ka.push_back({b - 4, kposA[b]});
                if (kposB[b] >= 0) kb.push_back({b - 4, kposB[b]});
            }
            const int nsA = merge_runs(ta, d.tA, kRowMaxSeg), nsB = merge_runs(tb, d.tB, kRowMaxSeg), nsC = merge_runs(tc, d.tC, kRowMaxSeg);
            const int nkA = merge_runs(ka, d.kA, kRowMaxKSeg), nkB = merge_runs(kb, d.kB, kRowMaxKSeg);
            if (nsA >= 0 && nsB >= 0 && nsC >= 0 && nkA >= 0 && nkB >= 0) {
                d.nsA = (uint8_t)nsA; d.nsB = (uint8_t)nsB; d.nsC = (uint8_t)nsC; d.nkA = (uint8_t)nkA; d.nkB = (uint8_t)nkB;
                const int tile_bits = nC - 6;                 // 2^tile_bits warp-tiles
                int per = std::max(0, std::min(2, tile_bits - 3));   // tiles per unit: 1, 2 or 4, at least 8 units when possible
                h.ma = (uint8_t)per;
                h.ntt = (uint8_t)tile_bits;
                n_units = 1 << (tile_bits - per);
                unit_cost = std::ldexp(1.0, per + nK) * 0.25 + 8.0;
                return true;
            }
            memset(&d.hot, 0, sizeof(d.hot));                 // too many segments: fall through to the SIMT tile
        }
    }
    // register tile: highest M-only / N-only bits, alternating sides, while >= 2^min_tt_bits thread-tiles remain
    std::vector<int> mcand, ncand, mbits, nbits;
    for (int b = nC - 1; b >= 0; --b) {
        if (mapA[b] >= 0 && mapB[b] < 0) mcand.push_back(b);
        else if (mapB[b] >= 0 && mapA[b] < 0) ncand.push_back(b);
    }
    while ((int)(mbits.size() + nbits.size()) < o.max_tile_bits && nC - (int)(mbits.size() + nbits.size()) - 1 >= o.min_tt_bits) {
        const bool can_m = mbits.size() < 2 && mbits.size() < mcand.size();
        const bool can_n = nbits.size() < 2 && nbits.size() < ncand.size();
        if (!can_m && !can_n) break;
        if (can_m && (!can_n || mbits.size() <= nbits.size())) mbits.push_back(mcand[mbits.size()]);
        else nbits.push_back(ncand[nbits.size()]);
    }
    int ma = (int)mbits.size(), nb = (int)nbits.size();
    int ntt = nC - ma - nb;
    RowOpHot& h = d.hot;
    if (ntt < 5) {                                       // fewer than 32 thread-tiles: lanes split K instead
        mbits.clear(); nbits.clear(); ma = nb = 0; ntt = nC;
    }
    static const bool bank_env = [] { const char* e = getenv("QXB_ROW_BANK_OPT"); return !e || atoi(e) != 0; }();
    const bool bank_opt = bank_env && o.bank_opt;
    // shared-memory wavefronts of one access of the lanes of a wavefront group (8 lanes for 16-byte elements, 16 for 8-byte
    // ones) whose index bits are the C bits `bits`: distinct addresses falling on the same 128-byte residue serialise
    const int q = dtype == QXB_C64 ? 3 : 4;
    auto waves = [&](const int* bits, const std::vector<int>& map) {
        int cnt[16] = {0}, seen[16], ns = 0;
        for (int l = 0; l < (1 << q); ++l) {
            int a = 0;
            for (int i = 0; i < q; ++i) if (((l >> i) & 1) && map[bits[i]] >= 0) a |= 1 << map[bits[i]];
            bool dup = false;
            for (int i = 0; i < ns; ++i) dup |= seen[i] == a;
            if (!dup) { seen[ns++] = a; ++cnt[a & ((1 << q) - 1)]; }
        }
        return *std::max_element(cnt, cnt + (1 << q));
    };
    // best group bits among the non-tile bits `ord` for a (2^tma x 2^tnb) register tile: (cost, bits)
    auto best_group = [&](const std::vector<int>& ord, int tma, int tnb, std::vector<int>& best) {
        const double nk = std::ldexp(1.0, nK), tm = std::ldexp(1.0, tma), tn = std::ldexp(1.0, tnb);
        const int n = (int)ord.size();
        double best_cost = -1;
        int idx[4], pick[4];
        for (int i = 0; i < q; ++i) idx[i] = i;
        while (true) {
            for (int i = 0; i < q; ++i) pick[i] = ord[idx[i]];
            const double wa = free_operand == 1 ? 1 : waves(pick, mapA), wb = free_operand == 2 ? 1 : waves(pick, mapB);
            const double cost = nk * (tm * wa + tn * wb) + tm * tn * waves(pick, mapC);
            if (best_cost < 0 || cost < best_cost) { best.assign(pick, pick + q); best_cost = cost; }
            int i = q - 1;
            while (i >= 0 && idx[i] == n - q + i) --i;
            if (i < 0) break;
            ++idx[i];
            for (int j = i + 1; j < q; ++j) idx[j] = idx[j - 1] + 1;
        }
        return best_cost;
    };
    if (bank_opt && o.bank_search_tiles && ntt >= 5 && ntt > q && ma + nb > 0) {
        // which M-only / N-only bits form the register tile: every choice of ma of the M-only and nb of the N-only bits
        std::vector<int> bm, bn;
        double bc = -1;
        std::vector<int> im(ma), in(nb);
        for (int i = 0; i < ma; ++i) im[i] = i;
        auto next_comb = [](std::vector<int>& idx, int n) {
            const int k = (int)idx.size();
            int i = k - 1;
            while (i >= 0 && idx[i] == n - k + i) --i;
            if (i < 0) return false;
            ++idx[i];
            for (int j = i + 1; j < k; ++j) idx[j] = idx[j - 1] + 1;
            return true;
        };
        do {
            for (int i = 0; i < nb; ++i) in[i] = i;
            do {
                std::vector<bool> tile(nC, false);
                for (int i : im) tile[mcand[i]] = true;
                for (int i : in) tile[ncand[i]] = true;
                std::vector<int> ord, grp;
                for (int b = 0; b < nC; ++b) if (!tile[b]) ord.push_back(b);
                const double c = best_group(ord, ma, nb, grp);
                if (bc < 0 || c < bc) {
                    bc = c; bm.clear(); bn.clear();
                    for (int i : im) bm.push_back(mcand[i]);
                    for (int i : in) bn.push_back(ncand[i]);
                }
            } while (next_comb(in, (int)ncand.size()));
        } while (next_comb(im, (int)mcand.size()));
        mbits = bm; nbits = bn;
    }
    std::sort(mbits.begin(), mbits.end()); std::sort(nbits.begin(), nbits.end());
    int kc = 0;
    if (ntt >= 5) {
        kc = std::min(nK, 2);
        while (kc > 0 && tile_regs(dtype, ma, nb, kc) > std::min(100, o.tile_reg_budget)) --kc;
        if (tile_regs(dtype, ma, nb, kc) > 100) return bad("no tile variant");
        h.kind = (uint8_t)row_tile_kind(ma, nb, kc, 0);
        h.ks = 0;
        n_units = 1 << (ntt - 5);
    } else {
        h.kind = kRowKindKred;
        h.ks = (uint8_t)std::min(5 - ntt, nK);
        n_units = 1;
    }
    h.nK = (uint8_t)nK; h.kc = (uint8_t)kc; h.ma = (uint8_t)ma; h.nb = (uint8_t)nb; h.ntt = (uint8_t)ntt;
    unit_cost = std::ldexp(1.0, ma + nb + nK - (int)h.ks) + 8.0;
    std::vector<bool> is_tile(nC, false);
    for (int b : mbits) is_tile[b] = true;
    for (int b : nbits) is_tile[b] = true;
    for (int jm = 0; jm < (1 << ma); ++jm) {
        int a = 0, c = 0;
        for (int t = 0; t < ma; ++t) if ((jm >> t) & 1) { a |= 1 << mapA[mbits[t]]; c |= 1 << mapC[mbits[t]]; }
        h.aT[jm] = (uint16_t)a;
        for (int jn = 0; jn < (1 << nb); ++jn) {
            int b = 0, c2 = c;
            for (int t = 0; t < nb; ++t) if ((jn >> t) & 1) { b |= 1 << mapB[nbits[t]]; c2 |= 1 << mapC[nbits[t]]; }
            h.bT[jn] = (uint16_t)b;
            h.cT[jm * (1 << nb) + jn] = (uint16_t)c2;
        }
    }
    for (int k = 0; k < (1 << std::min(nK, 4)); ++k) {
        int a = 0, b = 0;
        for (int t = 0; t < std::min(nK, 4); ++t) if ((k >> t) & 1) {
            if (kposA[t] >= 0) a |= 1 << kposA[t];
            if (kposB[t] >= 0) b |= 1 << kposB[t];
        }
        h.ktA[k] = (uint16_t)a; h.ktB[k] = (uint16_t)b;
    }
    // thread-tile index bits = the C bits outside the register tile; the low 5 of them are the lanes of a unit.
    // Which C bits become the LOW lane bits decides the shared-memory bank conflicts: the lanes of one wavefront group
    // (8 lanes for 16-byte elements, 16 for 8-byte ones) should hit distinct 128-byte residues -- or the same address --
    // in A, in B and in C.  With the lowest C bits as lanes (conflict-free for C only) the fused chain of the headline
    // plan spent 8960 wavefronts per row where 6016 suffice (scripts/analyze_chain_banks.py); pick the group bits that
    // minimise the op's wavefronts, ties to the lowest bits.
    std::vector<int> order;
    for (int b = 0; b < nC; ++b) if (!is_tile[b]) order.push_back(b);
    auto arrange = [&](bool optimise) {
        std::vector<int> ord = order;
        if (optimise && ntt >= 5 && (int)ord.size() > q) {
            std::vector<int> best, rest;
            best_group(ord, ma, nb, best);
            for (int b : ord) if (std::find(best.begin(), best.end(), b) == best.end()) rest.push_back(b);
            ord = best; ord.insert(ord.end(), rest.begin(), rest.end());
        }
        std::vector<std::pair<int, int>> ta, tb, tc;
        for (int t = 0; t < (int)ord.size(); ++t) {
            const int b = ord[t];
            tc.push_back({t, mapC[b]});
            if (mapA[b] >= 0) ta.push_back({t, mapA[b]});
            if (mapB[b] >= 0) tb.push_back({t, mapB[b]});
        }
        if (lay && lay->group_out) lay->group_out->assign(ord.begin(), ord.begin() + std::min<size_t>(ord.size(), (size_t)q));
        const int a = merge_runs(ta, d.tA, kRowMaxSeg), b2 = merge_runs(tb, d.tB, kRowMaxSeg), c = merge_runs(tc, d.tC, kRowMaxSeg);
        if (a < 0 || b2 < 0 || c < 0) return false;
        d.nsA = (uint8_t)a; d.nsB = (uint8_t)b2; d.nsC = (uint8_t)c;
        return true;
    };
    if (!(bank_opt && arrange(true)) && !arrange(false)) return bad("too many address segments");   // permuted bits split runs
    std::vector<std::pair<int, int>> ka, kb;
    for (int b = 4; b < nK; ++b) {
        if (kposA[b] >= 0) ka.push_back({b - 4, kposA[b]});
        if (kposB[b] >= 0) kb.push_back({b - 4, kposB[b]});
    }
    const int nkA = merge_runs(ka, d.kA, kRowMaxKSeg), nkB = merge_runs(kb, d.kB, kRowMaxKSeg);
    if (nkA < 0 || nkB < 0) return bad("too many address segments");
    const int nsA = d.nsA, nsB = d.nsB, nsC = d.nsC;
    d.nsA = (uint8_t)nsA; d.nsB = (uint8_t)nsB; d.nsC = (uint8_t)nsC; d.nkA = (uint8_t)nkA; d.nkB = (uint8_t)nkB;
    return true;
}

RowProgramHost build_row_program(const Lowered& L, Phase phase, int dtype, const RowPlanOptions& o,
                                 const std::vector<int>* subset) {
    RowProgramHost rp;
    rp.phase = phase;
    rp.elem_bytes = dtype == QXB_C32 ? 8 : 16;
    auto fail = [&](const std::string& why) { rp.ok = false; rp.why = why; return rp; };
    // ---- ops of the phase, producers, levels
    std::vector<int> sel;                                   // indices into L.ops
    if (subset) {
        if (phase != PH_CHUNK) return fail("a fused chain lives in the chunk phase");
        sel = *subset;
        for (int i : sel) if (i < 0 || i >= (int)L.ops.size() || L.ops[i].phase != PH_CHUNK) return fail("bad chain op");
    } else {
        for (size_t i = 0; i < L.ops.size(); ++i) if (L.ops[i].phase == phase) sel.push_back((int)i);
    }
    if (sel.empty()) return fail("no op in this phase");
    if (sel.size() > 60000) return fail("too many ops");
    std::map<int, int> local_of_lop;                        // L.ops index -> local index
    for (size_t j = 0; j < sel.size(); ++j) local_of_lop[sel[j]] = (int)j;
    const int n = (int)sel.size();
    std::map<int, int> producer;                            // LTensor -> local op
    for (int j = 0; j < n; ++j) producer[L.ops[sel[j]].c] = j;
    std::vector<std::vector<int>> deps(n), users(n);
    for (int j = 0; j < n; ++j) {
        const LOp& op = L.ops[sel[j]];
        std::vector<int> d;
        if (phase == PH_CHUNK) {
            // the arena is re-planned here: only true data dependencies count
            for (int t : {op.a, op.b}) { auto it = producer.find(t); if (it != producer.end()) d.push_back(it->second); }
        } else {
            // tensors stay where plan_memory put them: keep its write-after-read edges too
            for (int dd : op.deps) { auto it = local_of_lop.find(dd); if (it != local_of_lop.end()) d.push_back(it->second); }
        }
        std::sort(d.begin(), d.end()); d.erase(std::unique(d.begin(), d.end()), d.end());
        for (int x : d) { if (x >= j) return fail("ops not in topological order"); users[x].push_back(j); }
        deps[j] = d;
    }
    std::vector<int> level(n, 0);
    int n_levels = 0;
    for (int j = 0; j < n; ++j) {
        for (int d : deps[j]) level[j] = std::max(level[j], level[d] + 1);
        n_levels = std::max(n_levels, level[j] + 1);
    }
    if (o.alap) {
        // as late as possible: small early nodes run beside the big ones and their results live shorter
        for (int j = n - 1; j >= 0; --j) {
            if (users[j].empty()) { level[j] = std::max(level[j], 0); continue; }
            int lv = n_levels;
            for (int u : users[j]) lv = std::min(lv, level[u] - 1);
            level[j] = std::max(level[j], lv);
        }
    }
    // operands that live outside the chunk phase (leaves, const- and block-phase results): stage the small ones
    struct Staged { int tensor, copy_level, until, off; };
    std::vector<Staged> staged;
    if (phase == PH_CHUNK && o.stage_shared) {
        std::map<int, std::pair<int, int>> use;             // tensor -> (first, last) level of use
        for (int j = 0; j < n; ++j)
            for (int t : {L.ops[sel[j]].a, L.ops[sel[j]].b}) {
                const LTensor& T = L.tensors[t];
                if (producer.count(t)) continue;
                if (subset) {
                    if (T.span_bits > o.stage_max_bits) return fail("chain input larger than the staging limit");
                } else if (T.is_output_leaf || T.span_bits > o.stage_max_bits) continue;
                auto it = use.find(t);
                if (it == use.end()) use[t] = {level[j], level[j]};
                else { it->second.first = std::min(it->second.first, level[j]); it->second.second = std::max(it->second.second, level[j]); }
            }
        if (!use.empty()) {
            for (int j = 0; j < n; ++j) ++level[j];          // level 0 = the copies the first ops need
            ++n_levels;
            for (auto& kv : use) staged.push_back(Staged{kv.first, kv.second.first, kv.second.second + 1, -1});
        }
    }
    if (n_levels > kRowMaxLevels) return fail("more than " + std::to_string(kRowMaxLevels) + " levels");

    // ---- arena plan (chunk phase): first fit over level-granular live ranges (addresses are base + offsets, so a
    //      tensor may sit anywhere); operands shared by all rows are STAGED: copied into the arena by cp.async one
    //      level before their first use (a global load at use time is an L2 round trip on the critical path)
    std::map<int, int> arena_off;                           // LTensor -> element offset
    if (phase == PH_CHUNK) {
        auto last_level = [&](int t) {
            int lv = -1;
            for (int j = 0; j < n; ++j) if (L.ops[sel[j]].a == t || L.ops[sel[j]].b == t) lv = std::max(lv, level[j]);
            return lv;
        };
        std::vector<Interval> live;
        for (int t : subset ? std::vector<int>() : L.output_leaves) {
            const LTensor& T = L.tensors[t];
            if (T.span_bits > 16) return fail("output leaf too large");
            const int lu = last_level(t);
            if (lu < 0) continue;                               // unused leaf
            arena_off[t] = arena_alloc(live, 1 << T.span_bits, lu);
            rp.leaves.push_back(RowLeaf{arena_off[t], T.span_bits, (int)T.out_idx});
        }
        int peak = 0;
        for (const Interval& iv : live) peak = std::max(peak, iv.off + iv.size);
        for (int lv = 0; lv < n_levels; ++lv) {
            live.erase(std::remove_if(live.begin(), live.end(), [&](const Interval& iv) { return iv.until < lv; }), live.end());
            for (Staged& st : staged)
                if (st.copy_level == lv) {
                    st.off = arena_alloc(live, 1 << L.tensors[st.tensor].span_bits, st.until);
                    arena_off[st.tensor] = st.off;
                    peak = std::max(peak, st.off + (1 << L.tensors[st.tensor].span_bits));
                }
            std::vector<int> here;
            for (int j = 0; j < n; ++j) if (level[j] == lv) here.push_back(j);
            std::sort(here.begin(), here.end(), [&](int x, int y) { return L.ops[sel[x]].nC > L.ops[sel[y]].nC; });
            for (int j : here) {
                const LOp& op = L.ops[sel[j]];
                if (op.nC > 16) return fail("intermediate larger than 2^16 elements");
                const int t = op.c;
                int until = last_level(t);
                if (subset) {
                    // a chain result read outside the chain (or the root) goes straight to global memory
                    bool outside = t == L.root;
                    for (size_t q = 0; q < L.ops.size(); ++q)
                        if ((L.ops[q].a == t || L.ops[q].b == t) && !local_of_lop.count((int)q)) outside = true;
                    if (outside && until >= 0) return fail("a chain tensor is read both inside and outside the chain");
                    if (outside || until < 0) continue;           // not in the arena: in_arena_c = 0 -> global
                }
                if (t == L.root || until < 0) until = n_levels;       // the root (and anything unread) lives to the end
                arena_off[t] = arena_alloc(live, 1 << op.nC, until);
                peak = std::max(peak, arena_off[t] + (1 << op.nC));
            }
        }
        rp.arena_elems = peak;
        if (!subset) {
            if (!arena_off.count(L.root)) return fail("the root is not produced in the chunk phase");
            rp.root_off = arena_off[L.root];
            rp.root_span = L.tensors[L.root].span_bits;
        }
        const size_t es = dtype == QXB_C32 ? 8 : 16;
        if ((size_t)peak * es > (size_t)o.max_arena_bytes) return fail("row arena of " + std::to_string((size_t)peak * es) + " bytes exceeds the budget");
    }

    // ---- per-op descriptors
    rp.ops.resize(n); rp.ref_a.resize(n); rp.ref_b.resize(n); rp.ref_c.resize(n); rp.lop.resize(n);
    rp.in_arena_a.assign(n, 0); rp.in_arena_b.assign(n, 0); rp.in_arena_c.assign(n, 0);
    std::vector<double> unit_cost(n, 0);
    std::vector<int> n_units(n, 1);
    // Fused chain: an intermediate lives only in the arena, so its LAYOUT is free.  Backwards over the chain: an op's result
    // layout is what its consumer chose; the op then picks its lane-group bits with its own intermediate operand counted
    // conflict-free, and that operand's layout puts the group's bits at address bits 0, 1, 2 (distinct 16-byte bank groups).
    std::map<int, std::vector<int>> perm;                   // LTensor -> old address bit -> new address bit
    static const bool layout_opt = [] { const char* e = getenv("QXB_ROW_LAYOUT_OPT"); return !e || atoi(e) != 0; }();
    if (subset && o.bank_search_tiles && o.bank_opt && layout_opt) {
        std::map<int, int> n_readers;
        for (const LOp& op : L.ops) { ++n_readers[op.a]; ++n_readers[op.b]; }
        for (int j = n - 1; j >= 0; --j) {
            const LOp& op = L.ops[sel[j]];
            int freeop = 0, X = -1;
            for (int w = 1; w <= 2; ++w) {
                const int t = w == 1 ? op.a : op.b;
                if (producer.count(t) && arena_off.count(t) && n_readers[t] == 1 && t != L.root && op.a != op.b) { freeop = w; X = t; }
            }
            if (!freeop) continue;
            std::vector<int> grp;
            RowLayouts lay;
            auto pc = perm.find(op.c);
            lay.pc = pc != perm.end() ? pc->second.data() : nullptr;
            lay.free_operand = freeop; lay.group_out = &grp;
            RowOp tmp; int nu = 0; double uc = 0; std::string w2;
            if (!describe_row_op(op, dtype, o, tmp, nu, uc, w2, &lay) || grp.empty()) continue;
            const int span = L.tensors[X].span_bits;
            std::vector<int> mapX(op.nC, -1);
            for (auto& sg : (freeop == 1 ? op.segA : op.segB)) for (int b = 0; b < sg.len; ++b) mapX[sg.src + b] = sg.dst + b;
            std::vector<int> pi(16, -1);
            int cnt = 0;
            for (int gb : grp) if (mapX[gb] >= 0 && mapX[gb] < span && pi[mapX[gb]] < 0) pi[mapX[gb]] = cnt++;
            for (int b = 0; b < 16; ++b) if (pi[b] < 0) pi[b] = (b < span) ? cnt++ : b;
            perm[X] = pi;
        }
    }
    for (int j = 0; j < n; ++j) {
        const LOp& op = L.ops[sel[j]];
        RowOp& d = rp.ops[j];
        memset(&d, 0, sizeof(d));
        rp.lop[j] = sel[j]; rp.ref_a[j] = op.a; rp.ref_b[j] = op.b; rp.ref_c[j] = op.c;
        RowLayouts lay;
        { auto it = perm.find(op.a); if (it != perm.end()) lay.pa = it->second.data(); }
        { auto it = perm.find(op.b); if (it != perm.end()) lay.pb = it->second.data(); }
        { auto it = perm.find(op.c); if (it != perm.end()) lay.pc = it->second.data(); }
        if (!describe_row_op(op, dtype, o, d, n_units[j], unit_cost[j], rp.why, &lay)) { rp.ok = false; return rp; }
        RowOpHot& h = d.hot;
        // where the tensors live
        auto place = [&](int tensor, int& off, char& in_arena) {
            auto it = arena_off.find(tensor);
            if (it != arena_off.end()) { off = it->second; in_arena = 1; }
            else { off = 0; in_arena = 0; }
        };
        place(op.a, d.oA, rp.in_arena_a[j]); place(op.b, d.oB, rp.in_arena_b[j]); place(op.c, d.oC, rp.in_arena_c[j]);
        if (subset) {                                        // tensors left in global memory: per-row ones carry their stride
            if (!rp.in_arena_a[j]) return fail("chain operand not staged");
            if (!rp.in_arena_b[j]) return fail("chain operand not staged");
            if (!rp.in_arena_c[j] && L.tensors[op.c].amp) d.lsC = (uint8_t)L.tensors[op.c].span_bits;
        }
        h.gen = (uint8_t)(!(rp.in_arena_a[j] && rp.in_arena_b[j] && rp.in_arena_c[j]));
        const LTensor &TA = L.tensors[op.a], &TB = L.tensors[op.b];
        rp.flops_per_row += 8.0 * op.macs_per_amp;
        rp.elems_per_row_amp += (TA.amp ? op.elems_a : 0) + (TB.amp ? op.elems_b : 0) + op.elems_c;
        rp.elems_shared += (TA.amp ? 0 : op.elems_a) + (TB.amp ? 0 : op.elems_b);
    }

    // ---- copy pseudo-ops of the staged operands: 256 elements per unit
    for (const Staged& st : staged) {
        RowOp d;
        memset(&d, 0, sizeof(d));
        const int span = L.tensors[st.tensor].span_bits;
        d.hot.kind = kRowKindCopy;
        d.lsA = d.lsB = d.lsC = kRowShared;
        if (subset && L.tensors[st.tensor].amp) d.lsA = (uint8_t)span;     // a per-row tensor another kernel produced
        d.hot.ntt = (uint8_t)std::min(span, 8);              // log2 elements per unit
        d.hot.gen = 1;
        d.oC = st.off;
        rp.ops.push_back(d);
        rp.lop.push_back(-1); rp.ref_a.push_back(st.tensor); rp.ref_b.push_back(-1); rp.ref_c.push_back(-1);
        rp.in_arena_a.push_back(0); rp.in_arena_b.push_back(1); rp.in_arena_c.push_back(1);
        level.push_back(st.copy_level);
        unit_cost.push_back(4.0);
        n_units.push_back(1 << std::max(0, span - 8));
    }
    const int n_all = (int)rp.ops.size();

    // ---- units per level: costly units first, so the round-robin over warps balances
    rp.level_start.assign(n_levels + 1, 0);
    for (int lv = 0; lv < n_levels; ++lv) {
        std::vector<int> here;
        for (int j = 0; j < n_all; ++j) if (level[j] == lv) here.push_back(j);
        std::stable_sort(here.begin(), here.end(), [&](int x, int y) { return unit_cost[x] > unit_cost[y]; });
        for (int j : here)
            for (int c = 0; c < n_units[j]; ++c) rp.units.push_back(RowUnit{(uint16_t)j, (uint16_t)c});
        rp.level_start[lv + 1] = (int)rp.units.size();
    }
    if (rp.units.size() > 2048 && phase == PH_CHUNK) return fail("more than 2048 units");
    if (rp.units.size() > 40000) return fail("too many units");     // descriptor indices are 16 bits, 0xFFFF = none
    rp.n_levels = n_levels;
    rp.ok = true;
    return rp;
}


std::vector<int> select_chain(const Lowered& L, int dtype, const RowPlanOptions& o) {
    const int n = (int)L.ops.size();
    // the only consumer of every op's result (-1: none / several / the root)
    std::vector<int> next(n, -1), uses(L.tensors.size(), 0);
    for (int i = 0; i < n; ++i) { ++uses[L.ops[i].a]; ++uses[L.ops[i].b]; }
    std::vector<int> producer(L.tensors.size(), -1);
    for (int i = 0; i < n; ++i) producer[L.ops[i].c] = i;
    for (int i = 0; i < n; ++i) {
        const LOp& op = L.ops[i];
        if (op.phase != PH_CHUNK) continue;
        for (int t : {op.a, op.b}) {
            const int p = producer[t];
            if (p >= 0 && L.ops[p].phase == PH_CHUNK && uses[t] == 1 && t != L.root) next[p] = i;
        }
    }
    auto small = [&](int i) {
        const LOp& op = L.ops[i];
        return op.phase == PH_CHUNK && L.tensors[op.c].amp && op.nC <= o.stage_max_bits && op.nK <= 8 &&
               L.tensors[op.a].span_bits <= o.stage_max_bits && L.tensors[op.b].span_bits <= o.stage_max_bits;
    };
    std::vector<char> has_pred(n, 0);
    for (int i = 0; i < n; ++i) if (next[i] >= 0 && small(i) && small(next[i])) has_pred[next[i]] = 1;
    std::vector<int> best;
    double best_flops = 0;
    for (int i = 0; i < n; ++i) {
        if (!small(i) || has_pred[i]) continue;
        std::vector<int> path;
        double fl = 0;
        for (int j = i; j >= 0 && small(j); j = next[j]) { path.push_back(j); fl += L.ops[j].macs_per_amp; }
        if (path.size() >= 2 && fl > best_flops) { best = path; best_flops = fl; }
    }
    if (best.empty()) return best;
    // tiny nodes at either end cost a fused row a full dependency level each (~1300 cycles of latency with two rows in
    // flight per SM) but next to nothing as kernels of their own: keep the chain to the nodes that carry the work
    while (best.size() >= 2 && L.ops[best.front()].macs_per_amp < o.chain_min_macs) best.erase(best.begin());
    while (best.size() >= 2 && L.ops[best.back()].macs_per_amp < o.chain_min_macs) best.pop_back();
    if (best.size() < 2) { best.clear(); return best; }
    // trim the cheap ends until the chain's program fits two CTAs per SM (or give up at two ops)
    auto fits = [&](const std::vector<int>& c) {
        RowPlanOptions oo = o;
        RowProgramHost rp = build_row_program(L, PH_CHUNK, dtype, oo, &c);
        return rp.ok;
    };
    while (best.size() >= 2 && !fits(best)) {
        if (L.ops[best.front()].macs_per_amp <= L.ops[best.back()].macs_per_amp) best.erase(best.begin());
        else best.pop_back();
    }
    if (best.size() < 2) { best.clear(); return best; }
    // Side branches: the producer of a chain op's OTHER operand (small, read by nobody else) joins the program.  It sits
    // one level before its consumer, next to the previous chain op -- whose level keeps only half of the CTA's warps
    // busy (4 units of 32 thread-tiles for 2^11 outputs) -- so its work runs in idle issue slots and its result never
    // goes through HBM (as kernels of their own these nodes were 4.4 ms of a 10.3 ms step at 4-6 TB/s).  Not for the first
    // chain op (its producer would need a level of its own in front).  Greedy by work, while the arena still fits.
    // MEASURED (profiles/r2_summary.md): not a win -- the fused launch grows from 6.52 to 8.10 ms (depth 1: 5 side nodes worth
    // ~1 ms as kernels) / 8.81 ms (depth 2), the step from 10.26 to 10.92 / 10.83 ms: the extra staging copies and units
    // lengthen every level.  Off by default; qxb_options.chain_side (or QXB_CHAIN_SIDE=<depth>) enables it (GPU parity tests pass with it on).
    static const int side_env = [] { const char* e = getenv("QXB_CHAIN_SIDE"); return e ? atoi(e) : 0; }();
    const int side_depth = o.chain_side > 0 ? o.chain_side : side_env;
    if (side_depth > 0) {
        std::vector<char> in(n, 0);
        for (int i : best) in[i] = 1;
        std::vector<int> level_of(n, -1);
        for (size_t j = 0; j < best.size(); ++j) level_of[best[j]] = (int)j;
        std::vector<int> set = best;
        for (int round = 0; round < side_depth; ++round) {
            std::vector<std::pair<double, int>> cand;
            for (int i : set) {
                if (level_of[i] <= 0) continue;
                for (int t : {L.ops[i].a, L.ops[i].b}) {
                    const int p = producer[t];
                    if (p < 0 || in[p] || !small(p) || uses[t] != 1 || t == L.root || L.ops[p].nC > 10) continue;
                    cand.push_back({-L.ops[p].macs_per_amp, p});
                    level_of[p] = level_of[i] - 1;
                }
            }
            std::sort(cand.begin(), cand.end());
            bool any = false;
            for (auto& c : cand) {
                std::vector<int> trial = set;
                trial.push_back(c.second);
                std::sort(trial.begin(), trial.end());
                if (fits(trial)) { set = trial; in[c.second] = 1; any = true; }
            }
            if (!any) break;
        }
        best = set;
    }
    return best;
}

std::vector<int> make_contiguous(Lowered& L, const std::vector<int>& chain) {
    const int n = (int)L.ops.size();
    if (chain.size() < 2) return chain;
    std::vector<char> in_chain(n, 0);
    for (int i : chain) in_chain[i] = 1;
    const int last = chain.back();
    std::vector<int> order;                      // new position -> old index
    for (int i = 0; i < n; ++i) {
        if (in_chain[i] && i != last) continue;
        if (i == last) { for (int c : chain) order.push_back(c); continue; }
        order.push_back(i);
    }
    std::vector<int> new_of(n, -1);
    for (int p = 0; p < n; ++p) new_of[order[p]] = p;
    std::vector<LOp> ops(n);
    for (int p = 0; p < n; ++p) {
        ops[p] = L.ops[order[p]];
        for (int& d : ops[p].deps) d = new_of[d];
    }
    L.ops.swap(ops);
    for (LTensor& T : L.tensors) { T.first_use = -1; if (T.last_use != -1 || true) T.last_use = -1; }
    for (int p = 0; p < n; ++p)
        for (int t : {L.ops[p].a, L.ops[p].b}) {
            LTensor& T = L.tensors[t];
            if (T.first_use < 0) T.first_use = p;
            T.last_use = p;
        }
    std::vector<int> out;
    for (int c : chain) out.push_back(new_of[c]);
    L.fused_first = out.front(); L.fused_last = out.back();      // plan_memory: no operand released inside the launch
    return out;
}

static int eval_segs(const RSeg* s, int n, unsigned x) {
    int r = 0;
    for (int i = 0; i < n; ++i) r |= (int)(((x >> s[i].src) & ((1u << s[i].len) - 1u)) << s[i].dst);
    return r;
}

RowDeviceTables build_row_tables(const RowProgramHost& rp, const std::vector<RowOp>& ops) {
    RowDeviceTables t;
    t.level_start.assign(rp.n_levels + 1, 0);
    for (int lv = 0; lv < rp.n_levels; ++lv) {
        for (int u = rp.level_start[lv]; u < rp.level_start[lv + 1]; ++u) {
            const RowUnit un = rp.units[u];
            const RowOp& op = ops[un.op];
            RowUnitDesc d;
            memset(&d, 0, sizeof(d));
            d.hot = op.hot;
            d.gA = op.gA; d.gB = op.gB; d.gC = op.gC;
            memcpy(d.kA, op.kA, sizeof(d.kA)); memcpy(d.kB, op.kB, sizeof(d.kB));
            d.nkA = op.nkA; d.nkB = op.nkB;
            d.lsA = op.lsA; d.lsB = op.lsB; d.lsC = op.lsC;
            const int ntt = op.hot.ntt;
            if (op.hot.kind == kRowKindCopy) {
                // gA = first source element of this unit (set by the caller per op; advanced here per chunk), lC[0..1] =
                // destination element offset (32 bits), 2^ntt elements
                const long long first = (long long)un.chunk << 8;
                d.gA = op.gA + (unsigned long long)first * (unsigned long long)rp.elem_bytes;
                const unsigned dst = (unsigned)(op.oC + first);
                d.lC[0] = (uint16_t)(dst & 0xFFFF); d.lC[1] = (uint16_t)(dst >> 16);
                t.slots.push_back((uint16_t)t.descs.size());
                t.descs.push_back(d);
                t.desc_op.push_back(un.op);
                continue;
            }
            if (op.hot.kind == kRowKindDmma) {
                const int nt = 1 << op.hot.ma;
                for (int i = 0; i < nt; ++i) {
                    const unsigned wt = (unsigned)(un.chunk * nt + i);
                    d.hot.aT[i] = (uint16_t)eval_segs(op.tA, op.nsA, wt);
                    d.hot.bT[i] = (uint16_t)eval_segs(op.tB, op.nsB, wt);
                    d.hot.cT[i] = (uint16_t)eval_segs(op.tC, op.nsC, wt);
                }
                d.hot.cT[4] = op.c_n1;
                for (int lane = 0; lane < 32; ++lane) {
                    d.lA[lane] = (uint16_t)(op.oA + op.frA[lane]);
                    d.lB[lane] = (uint16_t)(op.oB + op.frB[lane]);
                    d.lC[lane] = (uint16_t)(op.oC + op.frC[lane]);
                }
                t.slots.push_back((uint16_t)t.descs.size());
                t.descs.push_back(d);
                t.desc_op.push_back(un.op);
                continue;
            }
            for (int lane = 0; lane < 32; ++lane) {
                int tt; bool active;
                if (op.hot.kind == kRowKindKred) { tt = lane & ((1 << ntt) - 1); active = (lane >> ntt) < (1 << op.hot.ks); }
                else { tt = un.chunk * 32 + lane; active = tt < (1 << ntt); }
                if (!active) { d.lA[lane] = d.lB[lane] = 0; d.lC[lane] = kRowNull; continue; }
                d.lA[lane] = (uint16_t)(op.oA + eval_segs(op.tA, op.nsA, (unsigned)tt));
                d.lB[lane] = (uint16_t)(op.oB + eval_segs(op.tB, op.nsB, (unsigned)tt));
                d.lC[lane] = (uint16_t)(op.oC + eval_segs(op.tC, op.nsC, (unsigned)tt));
            }
            t.slots.push_back((uint16_t)t.descs.size());
            t.descs.push_back(d);
            t.desc_op.push_back(un.op);
        }
        while (t.slots.size() % kRowWarps) t.slots.push_back(kRowNull);      // every warp has a slot in every round
        if ((int)t.slots.size() == t.level_start[lv]) for (int w = 0; w < kRowWarps; ++w) t.slots.push_back(kRowNull);
        t.level_start[lv + 1] = (int)t.slots.size();
    }
    return t;
}


std::vector<RowUnitDesc> build_ring_descs(const LOp& op, int dtype, const RowPlanOptions& o, std::string& why) {
    std::vector<RowUnitDesc> out;
    RowOp d;
    int n_units = 0;
    double cost = 0;
    if (!describe_row_op(op, dtype, o, d, n_units, cost, why)) return out;
    if (d.hot.kind == kRowKindKred) { why = "fewer than 32 thread-tiles"; return out; }
    RowProgramHost rp;
    rp.n_levels = 1;
    rp.level_start = {0, n_units};
    for (int c = 0; c < n_units; ++c) rp.units.push_back(RowUnit{0, (uint16_t)c});
    rp.elem_bytes = dtype == QXB_C32 ? 8 : 16;
    RowDeviceTables t = build_row_tables(rp, std::vector<RowOp>{d});
    return t.descs;
}

}  // namespace qxb
