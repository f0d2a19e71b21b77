// qxb200 -- executor and C ABI (include/qxb200.h).
//
// Replaces the executors behind the reference's hot path: QXTns.contract_tn!
// (call site /root/reference/src/simulation.jl:89) and the "Simulation" section of
// QXContexts.execute (call site /root/reference/bin/qxrun.jl:83-87):
//     for bitstring: for slice: run every ncon; acc += scalar
// Here the two loops are batch axes of each node's single launch (qxb_lower.cpp).
#include <cuda_runtime.h>
#include <cuda_profiler_api.h>

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <map>
#include <memory>
#include <cstdlib>
#include <set>
#include <sstream>
#include <tuple>

#include "../../include/qxb200.h"
#include "qxb_ir.h"
#include "qxb_kernels.cuh"
#include "qxb_rowplan.h"

using namespace qxb;

namespace {

thread_local std::string g_err;
// One host thread drives the library.  g_device is the CURRENT device; every device that was ever selected keeps its
// own stream and SM count (qxb_multi drives several devices of one node from the same thread).
int g_device = -1;
cudaStream_t g_own_stream = nullptr;
cudaStream_t g_ext_stream = nullptr;
bool g_use_ext = false;
int g_num_sms = 148;
struct DevState { cudaStream_t own = nullptr; int sms = 148; };
std::map<int, DevState> g_devs;

inline cudaStream_t stream() { return g_use_ext ? g_ext_stream : g_own_stream; }

#define CUDA_OK(expr)                                                                              \
    do {                                                                                           \
        cudaError_t e_ = (expr);                                                                   \
        if (e_ != cudaSuccess)                                                                     \
            throw Error(QXB_ERR_CUDA, std::string(#expr) + ": " + cudaGetErrorString(e_));         \
    } while (0)

// make `device` current: cudaSetDevice + this device's stream and SM count (created on first use)
void use_device(int device) {
    if (device == g_device) return;
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess || n == 0)
        throw Error(QXB_ERR_CUDA, "no CUDA device available (libqxb200 has no CPU fallback)");
    if (device < 0 || device >= n) throw Error(QXB_ERR_ARG, "device index out of range");
    CUDA_OK(cudaSetDevice(device));
    auto it = g_devs.find(device);
    if (it == g_devs.end()) {
        DevState st;
        CUDA_OK(cudaStreamCreateWithFlags(&st.own, cudaStreamNonBlocking));
        cudaDeviceProp prop;
        CUDA_OK(cudaGetDeviceProperties(&prop, device));
        st.sms = prop.multiProcessorCount;
        it = g_devs.emplace(device, st).first;
    }
    g_device = device;
    g_own_stream = it->second.own;
    g_num_sms = it->second.sms;
}

void ensure_init() {
    if (g_device >= 0) return;
    use_device(0);
}

// per-device, per-kernel one-time configuration (function attributes belong to the device's context)
bool first_use(const void* func) {
    static std::set<std::pair<int, const void*>> seen;
    return seen.insert({g_device, func}).second;
}

struct DevBuf {
    void* p = nullptr;
    size_t bytes = 0;
    bool reserve(size_t n) {      // true when the buffer moved
        if (n <= bytes) return false;
        if (p) { cudaFree(p); p = nullptr; bytes = 0; }
        cudaError_t e = cudaMalloc(&p, n);
        if (e != cudaSuccess) {
            p = nullptr;
            throw Error(QXB_ERR_MEM, "cudaMalloc of " + std::to_string(n) + " bytes failed: " + cudaGetErrorString(e));
        }
        bytes = n;
        return true;
    }
    void release() { if (p) cudaFree(p); p = nullptr; bytes = 0; }
};

struct OpProfile { double ms = 0; double flops = 0; double bytes = 0; long long launches = 0; const char* kernel = ""; };

struct Variant {
    Lowered L;
    int key = 0;                         // dense id (profile bookkeeping)
    DevBuf const_arena;
    DevBuf outleaf_desc;
    std::vector<OpParams> tmpl;          // per op, pointers unset
    std::vector<GemmParams> gtmpl;       // per op: tiled-GEMM parameters (valid when gemm_tmb > 0)
    std::vector<int> gemm_tmb, gemm_tnb;
    std::vector<GemmParams> gtmpl5;      // per op: parameters of the tcgen05 kernel (valid when gemm5[i])
    std::vector<char> gemm5;
    std::vector<OpProfile> prof;
    // row programs (qxb_rowprog.h): the chunk phase as one persistent kernel with the row's intermediates in shared
    // memory, the block phase as a single-CTA program; device copies of the descriptors per set of fixed values
    bool rows_block = false, rows_chunk = false;         // which phases have a row program
    RowProgramHost rp_block, rp_chunk;
    struct RowDev {
        DevBuf descs_block, slots_block, descs_chunk, slots_chunk, leaves;
        std::vector<int> level_start_block, level_start_chunk;      // in slots
        const void *block_base = nullptr, *const_base = nullptr;     // arena bases the pointers were resolved against
    };
    std::map<std::vector<int64_t>, RowDev> rowdev;
    OpProfile prof_rows, prof_block_rows;                // profile mode: the two fused launches
    DevBuf row_timing;                                   // profile mode: per-level clock cycles of CTA 0 (chunk program)
    // fused chain (qxb_rowplan.h select_chain): the dominant contractions as ONE row-program launch
    std::vector<int> chain;                              // op indices (contiguous after make_contiguous); empty: none
    RowProgramHost rp_chain;
    struct ChainDev { DevBuf descs, slots; std::vector<int> level_start; };
    std::vector<std::unique_ptr<ChainDev>> chain_dev;    // one per built step (pointers depend on the batch)
    OpProfile prof_chain;
    struct RingDescs { DevBuf buf; int n_units = 0; bool tried = false; };
    std::map<int, RingDescs> ring;                       // per op: unit descriptors of the TMA ring kernel
};

struct EventPair { cudaEvent_t a, b; int variant, op; };

// A whole qxb_amplitudes step captured as a CUDA graph, keyed by everything the
// launches depend on.  Replay removes ~550 launch gaps per step.
struct StepKey {
    const void* bits; void* out; int64_t n_amp, s0, s1; uint64_t mask; std::vector<int64_t> vals; cudaStream_t st;
    bool operator<(const StepKey& o) const {       // the fixed VALUES themselves, not a hash of them: no collisions
        return std::tie(bits, out, n_amp, s0, s1, mask, vals, st) <
               std::tie(o.bits, o.out, o.n_amp, o.s0, o.s1, o.mask, o.vals, o.st);
    }
};
struct Node;
// hoisted: the step is ONE block whose block phase (slice-only contractions: independent of the bitstrings) runs as a single
// row-program launch -- it is kept out of the captured graph and launched only when the block arena does not already hold
// the results for (variant, fixed values): a job that sends batch after batch of bitstrings over the same slices pays for
// the block phase once (47 us per call on the headline plan; the whole call is 95 us at 10 bitstrings).
struct StepGraph {
    cudaGraphExec_t exec = nullptr; qxb_stats stats{};
    bool hoisted = false; bool writes_block = false;
    std::shared_ptr<Node> block_launch; int block_variant = -1; std::vector<int64_t> block_vals;
};

}  // namespace

struct qxb_graph {
    int dtype = QXB_C64;
    Program prog;
    std::map<std::string, HostData> data;
    bool compiled = false;
    int device = -1;                 // the device the graph was compiled on (its buffers, streams and CUDA graphs live there)
    qxb_options opts{};
    std::map<std::string, DevBuf> leafbuf;
    std::map<uint64_t, std::unique_ptr<Variant>> variants;
    DevBuf block_arena, chunk_arena, acc, d_bits, d_out;
    // what the block arena holds (hoisted block phases, see StepGraph): variant key + fixed values; -1 = nothing known
    int block_owner_variant = -1; std::vector<int64_t> block_owner_vals; cudaStream_t block_owner_stream = nullptr;
    qxb_stats stats{};
    std::vector<EventPair> events;
    size_t events_used = 0;
    std::map<StepKey, StepGraph> step_graphs;
    std::map<uint64_t, int64_t> ws_cache;       // free-variable mask -> workspace bytes per bitstring row
    void drop_step_graphs() {
        for (auto& kv : step_graphs) if (kv.second.exec) cudaGraphExecDestroy(kv.second.exec);
        step_graphs.clear();
        for (auto& kv : variants) {                      // the chain tables those graphs pointed at
            for (auto& cd : kv.second->chain_dev) { cd->descs.release(); cd->slots.release(); }
            kv.second->chain_dev.clear();
        }
    }
    size_t es() const { return dtype == QXB_C32 ? 8 : 16; }
    ~qxb_graph() {
        for (auto& kv : leafbuf) kv.second.release();
        for (auto& kv : variants) {
            kv.second->const_arena.release(); kv.second->outleaf_desc.release(); kv.second->row_timing.release();
            for (auto& r : kv.second->ring) r.second.buf.release();
            for (auto& cd : kv.second->chain_dev) { cd->descs.release(); cd->slots.release(); }
            for (auto& rd : kv.second->rowdev) {
                rd.second.descs_block.release(); rd.second.slots_block.release(); rd.second.descs_chunk.release();
                rd.second.slots_chunk.release(); rd.second.leaves.release();
            }
        }
        block_arena.release(); chunk_arena.release(); acc.release(); d_bits.release(); d_out.release();
        for (auto& e : events) { cudaEventDestroy(e.a); cudaEventDestroy(e.b); }
        drop_step_graphs();
    }
};

namespace {

void ensure_analysed(qxb_graph* g) {
    if (!g->prog.analysed) analyse(g->prog);
}

// values per bitstring in the result: 1 for a closed network, prod of the saved tensor's extents for an open one
int64_t root_elems(qxb_graph* g) {
    ensure_analysed(g);
    int64_t n = 1;
    for (const Mode& m : g->prog.defs[g->prog.root].modes) n *= m.ext;
    return n;
}

// ------------------------------------------------------------------ leaves
void upload_leaves(qxb_graph* g) {
    for (const TensorDef& d : g->prog.defs) {
        if (d.kind != T_LOAD) continue;
        auto it = g->data.find(d.data_label);
        if (it == g->data.end())
            throw Error(QXB_ERR_STATE, "load " + d.name + ": no data set for label '" + d.data_label + "'");
        const HostData& h = it->second;
        std::vector<int64_t> dims;
        for (const Mode& m : d.modes) dims.push_back(m.full_ext);
        if (h.dims != dims)
            throw Error(QXB_ERR_ARG, "load " + d.name + ": dims do not match the data of '" + d.data_label + "'");
        if (g->leafbuf.count(d.data_label)) continue;
        int span = 0;
        std::vector<int> pos(dims.size());
        for (size_t m = 0; m < dims.size(); ++m) { pos[m] = span; span += d.modes[m].nbits; }
        const size_t n = size_t(1) << span;
        std::vector<double> buf64;
        std::vector<float> buf32;
        if (g->dtype == QXB_C64) buf64.assign(2 * n, 0.0); else buf32.assign(2 * n, 0.f);
        std::vector<int64_t> idx(dims.size(), 0);
        for (size_t lin = 0; lin < h.v.size(); ++lin) {
            size_t addr = 0;
            for (size_t m = 0; m < dims.size(); ++m) addr |= size_t(idx[m]) << pos[m];
            if (g->dtype == QXB_C64) { buf64[2 * addr] = h.v[lin].real(); buf64[2 * addr + 1] = h.v[lin].imag(); }
            else { buf32[2 * addr] = (float)h.v[lin].real(); buf32[2 * addr + 1] = (float)h.v[lin].imag(); }
            for (size_t m = 0; m < dims.size(); ++m) { if (++idx[m] < dims[m]) break; idx[m] = 0; }
        }
        DevBuf& db = g->leafbuf[d.data_label];
        db.reserve(std::max<size_t>(n, 2) * g->es());
        CUDA_OK(cudaMemcpy(db.p, g->dtype == QXB_C64 ? (void*)buf64.data() : (void*)buf32.data(), n * g->es(),
                           cudaMemcpyHostToDevice));
    }
}

// ------------------------------------------------------------- op launching
struct RunCtx {
    qxb_graph* g;
    Variant* v;
    int variant_key;
    const int64_t* fixed_vals;   // [k] values of the slice variables (only fixed ones are read)
    long long n;                 // bitstrings in this batch
};

char* tensor_ptr(const RunCtx& c, const LTensor& T) {
    qxb_graph* g = c.g;
    const size_t es = g->es();
    long long off = 0;
    for (auto& f : T.fixed) off += (long long)c.fixed_vals[f.first] << f.second;
    if (T.is_leaf && !T.is_output_leaf) {
        return (char*)g->leafbuf.at(T.data_label).p + off * es;
    }
    if (T.phase == PH_CHUNK) return (char*)g->chunk_arena.p + (T.offset * c.n + off) * es;
    if (T.phase == PH_BLOCK) return (char*)g->block_arena.p + (T.offset + off) * es;
    return (char*)c.v->const_arena.p + (T.offset + off) * es;
}

// Split the bits of C into thread bits / register-tile bits / hi bits and compose
// the address maps accordingly (see contract_kernel).
// a knob: the option if set, else the round-1 environment variable (experiments), else the default
// every entry 0, 1, 2 ('+') or 3 ('-')?  OR-reduction: vectorises, unlike an early-exit loop
bool bits_valid(const uint8_t* bits, size_t n) {
    unsigned char seen = 0;
    for (size_t i = 0; i < n; ++i) seen |= bits[i];
    return seen <= 3;
}

int knob(int opt, const char* env, int dflt) {
    if (opt > 0) return opt;
    const char* e = getenv(env);
    return e ? atoi(e) : dflt;
}

void build_templates(Variant& v, int dtype, const qxb_options& opts) {
    // register budget of the K chunk: 160 / 128 and min_lob 6 are what the measured choice of bench.py picked on two
    // separate B200s (BENCH_r01.json config.autotune: l1bw14/lob6/kc160; 128 / 96 was 4.6 % faster than 96 / 64 before)
    const int kc_regs_multi = knob(opts.kc_regs_multi, "QXB_KC_REGS_MULTI", 160);
    const int kc_regs_one = knob(opts.kc_regs_one, "QXB_KC_REGS_ONE", 128);
    const int min_lob = std::min(8, std::max(5, knob(opts.min_lob, "QXB_MIN_LOB", 6)));
    v.tmpl.resize(v.L.ops.size());
    v.gtmpl.resize(v.L.ops.size());
    v.gemm_tmb.assign(v.L.ops.size(), 0); v.gemm_tnb.assign(v.L.ops.size(), 0);
    v.gtmpl5.resize(v.L.ops.size()); v.gemm5.assign(v.L.ops.size(), 0);
    v.prof.assign(v.L.ops.size(), OpProfile{});
    for (size_t i = 0; i < v.L.ops.size(); ++i) {
        const LOp& op = v.L.ops[i];
        OpParams& p = v.tmpl[i];
        memset(&p, 0, sizeof(p));
        const int nC = op.nC;
        std::vector<int> mapA(nC, -1), mapB(nC, -1);
        for (auto& s : op.segA) for (int b = 0; b < s.len; ++b) mapA[s.src + b] = s.dst + b;
        for (auto& s : op.segB) for (int b = 0; b < s.len; ++b) mapB[s.src + b] = s.dst + b;
        p.nC = nC; p.nK = op.nK;
        // register tile: lowest M-only / N-only bits at C positions >= 5, so the 32 lanes of a warp
        // still cover C bits 0..4 (one contiguous 32-element store per warp and tile element)
        // QXB_MIN_LOB (5..8, default 8) = thread bits that must remain: below 8 the 256 threads of a CTA span
        // 2^(8 - lob) tiles (bitstring rows), which lets nodes with <= 2^8 elements per bitstring have a register
        // tile at all (experiment for the L1-bound small nodes of the tree-searched plans, profiles/r1p_summary.md)
        std::vector<int> mbits, nbits;
        if (nC > min_lob) {
            for (int b = 5; b < nC; ++b) {
                if (mapA[b] >= 0 && mapB[b] < 0 && mbits.size() < 2) mbits.push_back(b);
                else if (mapB[b] >= 0 && mapA[b] < 0 && nbits.size() < 2) nbits.push_back(b);
            }
            while ((int)(mbits.size() + nbits.size()) > nC - min_lob) {      // keep min_lob thread bits
                if (nbits.size() >= mbits.size() && !nbits.empty()) nbits.pop_back(); else mbits.pop_back();
            }
        }
        // K chunk.  Registers ~ RP * ((2^ma + 2^nb) * 2^kc  [staged operands]
        //                              + 2^(ma+nb)             [accumulators, only when K needs > 1 chunk]).
        // The tile comes first (it is what cuts L1/L2 traffic: (2^ma + 2^nb) / 2^(ma+nb) loads per output
        // and k); with a full 4x4 tile and K > 2^kc the schedule is the GEMM rank-1 update (kc = 0).
        const int rp = dtype == QXB_C32 ? 2 : 4;
        const int tmn = (int)((1u << mbits.size()) + (1u << nbits.size()));
        const int tile = 1 << (mbits.size() + nbits.size());
        int kc = std::min(op.nK, 3);
        while (kc > 0) {
            const bool multi = kc < op.nK;
            const int regs = rp * ((tmn << kc) + (multi ? tile : 0));
            if (regs <= (multi ? kc_regs_multi : kc_regs_one)) break;
            --kc;
        }
        p.kc = kc;
        p.ma = (int)mbits.size(); p.nb = (int)nbits.size();
        std::vector<bool> is_tile(nC, false);
        for (int b : mbits) is_tile[b] = true;
        for (int b : nbits) is_tile[b] = true;
        for (int jm = 0; jm < (1 << p.ma); ++jm) {
            long long a = 0, c = 0;
            for (int t = 0; t < p.ma; ++t) if ((jm >> t) & 1) { a |= 1ll << mapA[mbits[t]]; c |= 1ll << mbits[t]; }
            p.aT[jm] = a;
            for (int jn = 0; jn < (1 << p.nb); ++jn) {
                long long b = 0, c2 = c;
                for (int t = 0; t < p.nb; ++t) if ((jn >> t) & 1) { b |= 1ll << mapB[nbits[t]]; c2 |= 1ll << nbits[t]; }
                p.bT[jn] = b;
                p.cT[jm * (1 << p.nb) + jn] = c2;
            }
        }
        // thread bits = the lowest 8 (or fewer) non-tile positions, hi bits = the rest
        p.lob = std::min(nC - p.ma - p.nb, 8);
        std::vector<std::pair<int, std::pair<int, int>>> alo, blo, clo, ahi, bhi, chi;
        int t = 0, h = 0;
        for (int b = 0; b < nC; ++b) {
            if (is_tile[b]) continue;
            if (t < p.lob) {
                clo.push_back({t, {b, 1}});
                if (mapA[b] >= 0) alo.push_back({t, {mapA[b], 1}});
                if (mapB[b] >= 0) blo.push_back({t, {mapB[b], 1}});
                ++t;
            } else {
                chi.push_back({h, {b, 1}});
                if (mapA[b] >= 0) ahi.push_back({h, {mapA[b], 1}});
                if (mapB[b] >= 0) bhi.push_back({h, {mapB[b], 1}});
                ++h;
            }
        }
        p.hb = h;
        auto merge = [](const std::vector<std::pair<int, std::pair<int, int>>>& parts, DSeg* dst, int cap,
                        const std::string& name) {
            int n = 0;
            for (auto& pr : parts) {
                const int src = pr.first, d = pr.second.first, len = pr.second.second;
                if (n > 0 && dst[n - 1].src + dst[n - 1].len == src && dst[n - 1].dst + dst[n - 1].len == d) {
                    dst[n - 1].len = (unsigned char)(dst[n - 1].len + len);
                } else {
                    if (n == cap) throw Error(QXB_ERR_UNSUPP, "ncon " + name + ": too many address segments");
                    dst[n++] = DSeg{(unsigned char)src, (unsigned char)d, (unsigned char)len, 0};
                }
            }
            return n;
        };
        p.nsAlo = merge(alo, p.sAlo, 8, op.name); p.nsBlo = merge(blo, p.sBlo, 8, op.name);
        p.nsClo = merge(clo, p.sClo, 8, op.name);
        p.nsAhi = merge(ahi, p.sAhi, kMaxSeg, op.name); p.nsBhi = merge(bhi, p.sBhi, kMaxSeg, op.name);
        p.nsChi = merge(chi, p.sChi, kMaxSeg, op.name);
        if (op.segKA.size() > kMaxKSeg || op.segKB.size() > kMaxKSeg)
            throw Error(QXB_ERR_UNSUPP, "ncon " + op.name + ": too many address segments");
        p.nkA = (int)op.segKA.size(); p.nkB = (int)op.segKB.size();
        for (size_t j = 0; j < op.segKA.size(); ++j) p.kA[j] = DSeg{op.segKA[j].src, op.segKA[j].dst, op.segKA[j].len, 0};
        for (size_t j = 0; j < op.segKB.size(); ++j) p.kB[j] = DSeg{op.segKB[j].src, op.segKB[j].dst, op.segKB[j].len, 0};
        for (int k = 0; k < (1 << std::min(op.nK, 4)); ++k) {
            long long a = 0, b = 0;
            for (auto& s : op.segKA) a |= (long long)((k >> s.src) & ((1 << s.len) - 1)) << s.dst;
            for (auto& s : op.segKB) b |= (long long)((k >> s.src) & ((1 << s.len) - 1)) << s.dst;
            p.ktabA[k] = a; p.ktabB[k] = b;
        }
        // ---- GEMM-shaped node?  (K >= 2^kcb, >= 5 M-only and >= 5 N-only bits, spans <= 2^31: the kernels add
        //      distinct power-of-two offsets below 2^31 in 32-bit registers)
        const int kcb = gemm_kcb(dtype);
        std::vector<int> mall, nall;
        for (int b = 0; b < nC; ++b) {
            if (mapA[b] >= 0 && mapB[b] < 0) mall.push_back(b);
            else if (mapB[b] >= 0 && mapA[b] < 0) nall.push_back(b);
        }
        const LTensor &TA = v.L.tensors[op.a], &TB = v.L.tensors[op.b];
        // tile bits: the M-only (N-only) bits that sit lowest in A (B): ascending operand addresses per load
        std::sort(mall.begin(), mall.end(), [&](int x, int y) { return mapA[x] < mapA[y]; });
        std::sort(nall.begin(), nall.end(), [&](int x, int y) { return mapB[x] < mapB[y]; });
        // one GemmParams for a (tmb, tnb, kcb) block tile; smem_bit(is_b, bit) = layout of the kernel that will read it
        auto make_gemm = [&](GemmParams& q, int tmb, int tnb, int kcb_, auto smem_bit) {
            memset(&q, 0, sizeof(q));
            std::vector<int> mt(mall.begin(), mall.begin() + tmb), nt(nall.begin(), nall.begin() + tnb);
            std::vector<long long> kposA(op.nK, 0), kposB(op.nK, 0);
            for (auto& s : op.segKA) for (int b = 0; b < s.len; ++b) kposA[s.src + b] = 1ll << (s.dst + b);
            for (auto& s : op.segKB) for (int b = 0; b < s.len; ++b) kposB[s.src + b] = 1ll << (s.dst + b);
            std::vector<std::pair<long long, int>> la, lb;        // (global offset, tile-index bit)
            for (int t = 0; t < tmb; ++t) la.push_back({1ll << mapA[mt[t]], t});
            for (int c = 0; c < kcb_; ++c) la.push_back({kposA[c], tmb + c});
            for (int t = 0; t < tnb; ++t) lb.push_back({1ll << mapB[nt[t]], t});
            for (int c = 0; c < kcb_; ++c) lb.push_back({kposB[c], tnb + c});
            std::sort(la.begin(), la.end()); std::sort(lb.begin(), lb.end());
            for (size_t j = 0; j < la.size(); ++j) {
                q.aLoadOff[j] = la[j].first; q.aLoadSm[j] = 1 << la[j].second;
                q.aLoadSmT[j] = smem_bit(false, la[j].second);
            }
            for (size_t j = 0; j < lb.size(); ++j) {
                q.bLoadOff[j] = lb[j].first; q.bLoadSm[j] = 1 << lb[j].second;
                q.bLoadSmT[j] = smem_bit(true, lb[j].second);
            }
            for (int t = 0; t < tmb; ++t) q.cM[t] = 1ll << mt[t];
            for (int t = 0; t < tnb; ++t) q.cN[t] = 1ll << nt[t];
            std::vector<bool> tile_bit(nC, false);
            for (int b : mt) tile_bit[b] = true;
            for (int b : nt) tile_bit[b] = true;
            std::vector<std::pair<int, std::pair<int, int>>> gah, gbh, gch;
            int gh = 0;
            for (int b = 0; b < nC; ++b) {
                if (tile_bit[b]) continue;
                gch.push_back({gh, {b, 1}});
                if (mapA[b] >= 0) gah.push_back({gh, {mapA[b], 1}});
                if (mapB[b] >= 0) gbh.push_back({gh, {mapB[b], 1}});
                ++gh;
            }
            q.hb = gh; q.nK = op.nK;
            q.nsAhi = merge(gah, q.sAhi, kMaxSeg, op.name); q.nsBhi = merge(gbh, q.sBhi, kMaxSeg, op.name);
            q.nsChi = merge(gch, q.sChi, kMaxSeg, op.name);
            q.nkA = p.nkA; q.nkB = p.nkB;
            memcpy(q.kA, p.kA, sizeof(q.kA)); memcpy(q.kB, p.kB, sizeof(q.kB));
        };
        if (op.nK >= kcb && mall.size() >= 5 && nall.size() >= 5 && nC <= 31 && TA.span_bits <= 31 && TB.span_bits <= 31) {
            const int tmb = (int)std::min<size_t>(6, mall.size()), tnb = (int)std::min<size_t>(6, nall.size());
            make_gemm(v.gtmpl[i], tmb, tnb, kcb, [&](bool is_b, int bit) { return gemm_mma_smem_bit(dtype, is_b, is_b ? tnb : tmb, bit); });
            if (gemm_func(dtype, tmb, tnb)) { v.gemm_tmb[i] = tmb; v.gemm_tnb[i] = tnb; }
        }
        // tcgen05 / TMEM kernel (ComplexF32): 2^7 x 2^6 x 2^4 block tile
        // (K >= 2^6: with fewer than four 16-k chunks per tile the two-stage pipeline never fills -- measured on the K = 2^4
        //  nodes of the Sycamore depth-7 plan: 10.8 TFLOP/s against 14.3 for the mma.sync kernel, profiles/r2_summary.md)
        if (dtype == QXB_C32 && op.nK >= knob(0, "QXB_TC5_MIN_K", 6) && mall.size() >= 7 && nall.size() >= 6 && nC <= 31 &&
            TA.span_bits <= 31 && TB.span_bits <= 31) {
            make_gemm(v.gtmpl5[i], 7, 6, 4, [&](bool is_b, int bit) { return gemm_tc5_smem_bit(is_b ? 6 : 7, bit); });
            v.gemm5[i] = 1;
        }
    }
}

// ------------------------------------------------------------------ step nodes
// A step is a DAG of kernel nodes.  Edges: producer -> consumer, last reader ->
// re-user of an arena region (from plan_memory), and the serial spine
// memset -> [block: (chunk: output leaves -> ... -> root reduce)*]* -> finalize.
// The same list is either launched in order on the stream (profiling / no-graph
// mode) or instantiated as a CUDA graph with exactly these edges, so independent
// branches of the contraction tree run concurrently.
struct Node {
    const void* func = nullptr;          // nullptr: memset node
    dim3 grid{1, 1, 1}, block{1, 1, 1};
    std::vector<char> argmem;            // packed argument values
    std::vector<size_t> argoff;          // offset of each argument in argmem
    std::vector<int> deps;
    void* ms_ptr = nullptr; size_t ms_bytes = 0;
    void* pre_zero_ptr = nullptr; size_t pre_zero_bytes = 0;   // kernel nodes: region to clear first (split-K partial sums)
    int variant = -1, op = -1;           // contraction nodes: where to book profile time
    const char* kname = "";              // which kernel family runs the node (profile dump)
    double flops = 0, bytes = 0;
    size_t smem = 0;
    template <typename T> void arg(const T& v) {
        size_t off = (argmem.size() + alignof(T) - 1) / alignof(T) * alignof(T);
        if (alignof(T) < 16 && sizeof(T) >= 16) off = (argmem.size() + 15) / 16 * 16;
        argmem.resize(off + sizeof(T));
        memcpy(argmem.data() + off, &v, sizeof(T));
        argoff.push_back(off);
    }
};

// 0 (auto) resolves to the kernel measured faster on B200 for the dtype; QXB_GEMM_MODE overrides (experiments)
int gemm_mode(const qxb_graph* g) {
    static const int env = [] { const char* e = getenv("QXB_GEMM_MODE"); return e ? atoi(e) : 0; }();
    int m = env ? env : g->opts.gemm_mode;
    if (m == 0) m = 2;      // measured (profiles/r1p_gemm.md): DMMA 27 vs DFMA 18 TFLOP/s; 3xTF32 49 vs FFMA 38 TFLOP/s
    return m;
}

Node contract_node(const RunCtx& c, int i) {
    qxb_graph* g = c.g;
    Lowered& L = c.v->L;
    const LOp& op = L.ops[i];
    const LTensor &A = L.tensors[op.a], &B = L.tensors[op.b], &C = L.tensors[op.c];
    OpParams p = c.v->tmpl[i];
    p.A = tensor_ptr(c, A); p.B = tensor_ptr(c, B); p.C = tensor_ptr(c, C);
    p.sUA = A.amp ? (1ll << A.span_bits) : 0;
    p.sUB = B.amp ? (1ll << B.span_bits) : 0;
    p.sUC = C.amp ? (1ll << C.span_bits) : 0;
    p.U = C.amp ? (int)c.n : 1;
    p.tiles = (long long)p.U << p.hb;
    const int sub_bits = 8 - p.lob;
    const long long blocks = (p.tiles + (1ll << sub_bits) - 1) >> sub_bits;
    const long long cap = (long long)g_num_sms * 8;
    Node n;
    int tma_stages = 0;                    // > 0: the node runs contract_tma_kernel (second kernel argument)
    const double outputs = (double)p.U * std::ldexp(1.0, p.nC);
    // "big x small" streaming node (qxb_kred.cu, bigsmall_kernel): one operand huge, the other tiny with a few N-only
    // bits -- every thread owns one position of the big operand and 2^nlo outputs (all of them when the small operand
    // has <= 5 N bits; further N bits go to the CTA index, next to the thread bits so that the re-reads hit in L2)
    if (!g->opts.no_gemm && g->opts.streaming != 1 && knob(0, "QXB_BIGSMALL", 1) != 0 && op.n_batch == 0 && op.nK <= 5) {
        const bool a_big = op.elems_a >= op.elems_b;
        const LTensor &TB = a_big ? A : B, &TS = a_big ? B : A;
        const auto& segBig = a_big ? op.segA : op.segB;
        const auto& segSm = a_big ? op.segB : op.segA;
        const auto& kBig = a_big ? op.segKA : op.segKB;
        const auto& kSm = a_big ? op.segKB : op.segKA;
        const int nN = a_big ? op.n_n : op.n_m;
        const double big_elems = a_big ? op.elems_a : op.elems_b;
        const bool packed = g->dtype == QXB_C32 && (g->opts.streaming == 3 || knob(0, "QXB_BIGSMALL_FFMA2", 0) != 0);
        const int max_lo = (g->dtype == QXB_C32 && !packed) ? 5 : 4;
        const int nlo = nN <= max_lo ? nN : 4, nhi = nN - nlo;
        const void* bf = nN >= 1 ? bigsmall_func(g->dtype, nlo, knob(0, "QXB_BIGSMALL_KT", 1) != 0 ? op.nK : 0, packed) : nullptr;
        const size_t small_bytes = ((size_t)1 << (op.nK + nN)) * g->es();
        if (bf && big_elems >= std::ldexp(1.0, knob(0, "QXB_BIGSMALL_MIN_BITS", 20)) && small_bytes <= 64 * 1024 && nhi <= 8 &&
            TS.span_bits <= 24 && TB.span_bits <= 40 && op.nK + nN >= 3 && p.nC - nN >= 8) {
            std::vector<int> mapBig(p.nC, -1), mapSm(p.nC, -1);
            for (auto& sg : segBig) for (int b = 0; b < sg.len; ++b) mapBig[sg.src + b] = sg.dst + b;
            for (auto& sg : segSm) for (int b = 0; b < sg.len; ++b) mapSm[sg.src + b] = sg.dst + b;
            BigSmallParams q;
            memset(&q, 0, sizeof(q));
            std::vector<int> nbits, pbits;
            for (int b = 0; b < p.nC; ++b) (mapSm[b] >= 0 ? nbits : pbits).push_back(b);
            // position index layout: [8 position bits][nhi N bits][remaining position bits]
            std::vector<std::pair<int, int>> lay;          // (C bit, big-operand bit or -1)
            for (int j = 0; j < 8; ++j) lay.push_back({pbits[j], mapBig[pbits[j]]});
            for (int j = 0; j < nhi; ++j) lay.push_back({nbits[nlo + j], -1});
            for (size_t j = 8; j < pbits.size(); ++j) lay.push_back({pbits[j], mapBig[pbits[j]]});
            int nta = 0, ntc = 0;
            bool ok = (int)nbits.size() == nN;
            auto push = [&](DSeg* dst, int& cnt, int src, int d) {
                if (cnt > 0 && dst[cnt - 1].src + dst[cnt - 1].len == src && dst[cnt - 1].dst + dst[cnt - 1].len == d) { ++dst[cnt - 1].len; return; }
                if (cnt == 16) { ok = false; return; }
                dst[cnt++] = DSeg{(unsigned char)src, (unsigned char)d, 1, 0};
            };
            for (size_t j = 0; j < lay.size() && ok; ++j) {
                push(q.tC, ntc, (int)j, lay[j].first);
                if (lay[j].second >= 0) push(q.tA, nta, (int)j, lay[j].second);
            }
            if (ok) {
                q.ntA = nta; q.ntC = ntc; q.nK = op.nK; q.nNhi = nhi;
                q.n_pos = 1ll << lay.size();
                for (int k = 0; k < (1 << op.nK); ++k) {
                    long long a = 0, b = 0;
                    for (auto& sg : kBig) a |= (long long)((k >> sg.src) & ((1 << sg.len) - 1)) << sg.dst;
                    for (auto& sg : kSm) b |= (long long)((k >> sg.src) & ((1 << sg.len) - 1)) << sg.dst;
                    q.aK[k] = a; q.bK[k] = (int)b;
                }
                for (int j = 0; j < (1 << nlo); ++j) {
                    long long cc = 0, b = 0;
                    for (int t = 0; t < nlo; ++t) if ((j >> t) & 1) { cc |= 1ll << nbits[t]; b |= 1ll << mapSm[nbits[t]]; }
                    q.cN[j] = cc; q.bN[j] = (int)b;
                }
                for (int j = 0; j < (1 << nhi); ++j) {
                    long long b = 0;
                    for (int t = 0; t < nhi; ++t) if ((j >> t) & 1) b |= 1ll << mapSm[nbits[nlo + t]];
                    q.bH[j] = (int)b;
                }
                q.big = a_big ? p.A : p.B; q.small_ = a_big ? p.B : p.A; q.C = p.C;
                q.sUbig = a_big ? p.sUA : p.sUB; q.sUsmall = a_big ? p.sUB : p.sUA; q.sUC = p.sUC;
                q.U = p.U;
                // TMA variant: the 256 positions of a CTA are one contiguous run of the big operand at every k
                const size_t small_pad = (small_bytes + 15) & ~(size_t)15;
                const int stages = (int)std::min<size_t>(3, (110 * 1024 - small_pad - 64) / kBigSmallStageBytes);
                const void* tf = ((g->opts.streaming == 2 || knob(0, "QXB_BIGSMALL_TMA", 0) != 0) && q.U == 1 && stages >= 2 && q.tA[0].src == 0 && q.tA[0].dst == 0 &&
                                  q.tA[0].len >= 8 && (uintptr_t)q.big % 16 == 0) ? bigsmall_tma_func(g->dtype, nlo) : nullptr;
                n.block = dim3(kThreads);
                if (tf) {
                    n.func = tf; n.kname = "bigsmall_tma";
                    n.smem = (size_t)stages * kBigSmallStageBytes + small_pad + 8 * (size_t)stages;
                    if (first_use(n.func))
                        CUDA_OK(cudaFuncSetAttribute(n.func, cudaFuncAttributeMaxDynamicSharedMemorySize, 112 * 1024));
                    n.grid = dim3((unsigned)std::max<long long>(1, std::min<long long>(q.n_pos >> 8, (long long)g_num_sms * 2)));
                    n.arg(q); n.arg(stages);
                } else {
                    n.func = bf; n.kname = "bigsmall";
                    n.smem = small_bytes;
                    if (first_use(n.func))
                        CUDA_OK(cudaFuncSetAttribute(n.func, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024));
                    // measured: ~10 waves of CTAs (4736) beat one persistent wave sized by the occupancy (53.8 vs 58.3 ms per
                    // Sycamore depth-12 slice)
                    n.grid = dim3((unsigned)std::max<long long>(1, std::min<long long>(q.n_pos >> 8, cap * 4)));
                    n.arg(q);
                }
                n.variant = c.variant_key; n.op = i;
                const double u = (double)p.U;
                n.flops = 8.0 * op.macs_per_amp * u;
                n.bytes = (double)g->es() * (op.elems_a * (A.amp ? c.n : 1) + op.elems_b * (B.amp ? c.n : 1) + op.elems_c * u);
                return n;
            }
        }
    }
    // tcgen05 / TMEM 3xTF32 kernel (ComplexF32): auto prefers it wherever the shape has 2^7 M-only x 2^6 N-only bits
    // (81 TFLOP/s complex-equivalent stand-alone against 49 for the mma.sync kernel); gemm_mode 1 / 2 keep the others
    const int gm = gemm_mode(g);
    if (!g->opts.no_gemm && g->dtype == QXB_C32 && c.v->gemm5[i] && outputs >= 65536.0 && gm == 2) {
        GemmParams q = c.v->gtmpl5[i];
        q.A = p.A; q.B = p.B; q.C = p.C; q.sUA = p.sUA; q.sUB = p.sUB; q.sUC = p.sUC; q.U = p.U;
        q.tiles = (long long)p.U << q.hb;
        n.func = gemm_tc5_func(); n.kname = "gemm_tc5";
        n.block = dim3((unsigned)gemm_tc5_threads());
        n.smem = gemm_tc5_smem_bytes();
        if (first_use(n.func))
            CUDA_OK(cudaFuncSetAttribute(n.func, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)n.smem));
        n.grid = dim3((unsigned)std::max<long long>(1, std::min<long long>(q.tiles, (long long)g_num_sms)));
        n.arg(q);
        n.variant = c.variant_key; n.op = i;
        const double u = (double)p.U;
        n.flops = 8.0 * op.macs_per_amp * u;
        n.bytes = (double)g->es() * (op.elems_a * (A.amp ? c.n : 1) + op.elems_b * (B.amp ? c.n : 1) + op.elems_c * u);
        return n;
    }
    if (!g->opts.no_gemm && c.v->gemm_tmb[i] > 0 && outputs >= 65536.0) {
        // GEMM-shaped node: shared-memory-tiled FMA GEMM
        GemmParams q = c.v->gtmpl[i];
        q.A = p.A; q.B = p.B; q.C = p.C; q.sUA = p.sUA; q.sUB = p.sUB; q.sUC = p.sUC; q.U = p.U;
        q.tiles = (long long)p.U << q.hb;
        const int tmb = c.v->gemm_tmb[i], tnb = c.v->gemm_tnb[i];
        const void* tc = gm >= 2 ? gemm_mma_func(g->dtype, tmb, tnb) : nullptr;
        if (tc) {
            // tensor-core variant (DMMA / 3xTF32 mma.sync); persistent grid = SMs x resident CTAs
            n.func = tc; n.kname = "gemm_tc";
            n.block = dim3((unsigned)gemm_mma_threads(g->dtype));
            n.smem = gemm_mma_smem_bytes(g->dtype);
            static std::map<std::pair<int, const void*>, int> resident;
            auto it = resident.find({g_device, tc});
            if (it == resident.end()) {
                int nb = 0;
                if (n.smem > 48 * 1024)
                    CUDA_OK(cudaFuncSetAttribute(tc, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)n.smem));
                CUDA_OK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, tc, (int)n.block.x, n.smem));
                it = resident.emplace(std::make_pair(g_device, tc), std::max(nb, 1)).first;
            }
            n.grid = dim3((unsigned)std::max<long long>(1, std::min<long long>(q.tiles, (long long)g_num_sms * it->second)));
        } else {
            n.func = gemm_func(g->dtype, tmb, tnb); n.kname = "gemm_simt";
            n.grid = dim3((unsigned)std::max<long long>(1, std::min<long long>(q.tiles, (long long)g_num_sms * 2)));
            n.block = dim3(1u << (tmb + tnb - 4));
        }
        n.arg(q);
        n.variant = c.variant_key; n.op = i;
        const double u = (double)p.U;
        n.flops = 8.0 * op.macs_per_amp * u;
        n.bytes = (double)g->es() * (op.elems_a * (A.amp ? c.n : 1) + op.elems_b * (B.amp ? c.n : 1) + op.elems_c * u);
        return n;
    }
    if (p.nK >= 5 && p.nC <= 8 && outputs < 32768.0) {
        // reduction-shaped: too few outputs to fill the GPU with one thread each
        // the reduction kernels index an output by ALL its C bits: undo the register-tile split of the template
        // (since min_lob < 8 is the default, nodes of 2^7 - 2^8 outputs carry tile bits)
        if (p.ma + p.nb > 0 || p.hb > 0) {
            if (op.segA.size() > 8 || op.segB.size() > 8) throw Error(QXB_ERR_UNSUPP, "ncon " + op.name + ": too many address segments");
            p.ma = p.nb = 0; p.hb = 0; p.lob = p.nC;
            p.nsAlo = (int)op.segA.size(); p.nsBlo = (int)op.segB.size(); p.nsClo = 1;
            for (size_t j = 0; j < op.segA.size(); ++j) p.sAlo[j] = DSeg{op.segA[j].src, op.segA[j].dst, op.segA[j].len, 0};
            for (size_t j = 0; j < op.segB.size(); ++j) p.sBlo[j] = DSeg{op.segB[j].src, op.segB[j].dst, op.segB[j].len, 0};
            p.sClo[0] = DSeg{0, 0, (unsigned char)p.nC, 0};
            p.nsAhi = p.nsBhi = p.nsChi = 0;
            p.tiles = (long long)p.U;
        }
        // tiled split-K: operand tiles of a K chunk staged in shared memory once per CTA (instead of one pass over K per
        // output); needs few distinct operand rows and a single small batch of rows
        bool tiled = false;
        if (p.nK >= 12 && p.U <= 64 && !g->opts.no_gemm && knob(0, "QXB_KRED_TILE", 1) != 0) {
            KredTile kt;
            memset(&kt, 0, sizeof(kt));
            // The K chunk a CTA stages = the low kKredTileKBits bits of the k index.  The two operands usually store the
            // K bits in different orders (one of them can be scrambled: every 8-byte element of a chunk in its own 32-byte
            // sector, 4x over-fetch).  Re-number k so that the chunk bits are the K bits sitting lowest in EITHER operand
            // (sorted by the lower of the two address positions): both operands then read whole sectors.
            if (knob(0, "QXB_KRED_KORDER", 1) != 0 && p.nK <= kMaxKSeg) {
                std::vector<int> posA(p.nK, 0), posB(p.nK, 0), order(p.nK);
                for (auto& sg : op.segKA) for (int b = 0; b < sg.len; ++b) posA[sg.src + b] = sg.dst + b;
                for (auto& sg : op.segKB) for (int b = 0; b < sg.len; ++b) posB[sg.src + b] = sg.dst + b;
                for (int b = 0; b < p.nK; ++b) order[b] = b;
                // round robin: the K bit sitting lowest in A, then the one lowest in B, ... -- each operand gets a contiguous
                // run from its address bit 0 up (its row bits in between are read by the same CTA): 256 - 512 B pieces.
                // (Sorting by the lower of the two positions gave 64 B pieces on the scrambled operand: 1.4 TB/s.)
                std::vector<int> byA = order, byB = order, picked;
                std::stable_sort(byA.begin(), byA.end(), [&](int x, int y) { return posA[x] < posA[y]; });
                std::stable_sort(byB.begin(), byB.end(), [&](int x, int y) { return posB[x] < posB[y]; });
                std::vector<bool> taken(p.nK, false);
                for (size_t ia = 0, ib = 0; (int)picked.size() < kKredTileKBits;) {
                    auto& lst = (picked.size() % 2 == 0) ? byA : byB;
                    size_t& it = (picked.size() % 2 == 0) ? ia : ib;
                    while (taken[lst[it]]) ++it;
                    taken[lst[it]] = true; picked.push_back(lst[it]);
                }
                std::vector<int> unpicked;
                for (int b = 0; b < p.nK; ++b) if (!taken[b]) unpicked.push_back(b);
                std::vector<int> perm = picked, rest = unpicked;
                std::sort(perm.begin(), perm.end(), [&](int x, int y) { return posA[x] < posA[y]; });   // staging threads walk A ascending
                // ... and B ascending: the threads staging B take the chunk's k in B's bit order (kperm_b)
                std::vector<int> cb(kKredTileKBits);
                for (int j = 0; j < kKredTileKBits; ++j) cb[j] = j;
                std::sort(cb.begin(), cb.end(), [&](int x, int y) { return posB[perm[x]] < posB[perm[y]]; });
                for (int j = 0; j < kKredTileKBits; ++j) kt.kperm_b[j] = (unsigned char)cb[j];
                kt.kperm_set = 1;
                std::sort(rest.begin(), rest.end());
                perm.insert(perm.end(), rest.begin(), rest.end());
                auto rebuild = [&](const std::vector<int>& pos, DSeg* dst) {
                    int cnt = 0;
                    for (int j = 0; j < p.nK; ++j) {
                        const int d = pos[perm[j]];
                        if (cnt > 0 && dst[cnt - 1].src + dst[cnt - 1].len == j && dst[cnt - 1].dst + dst[cnt - 1].len == d) ++dst[cnt - 1].len;
                        else dst[cnt++] = DSeg{(unsigned char)j, (unsigned char)d, 1, 0};
                    }
                    return cnt;
                };
                p.nkA = rebuild(posA, p.kA);
                p.nkB = rebuild(posB, p.kB);
            }
            std::map<long long, int> ra, rb;
            bool ok = true;
            for (int cc = 0; cc < (1 << p.nC) && ok; ++cc) {
                long long oa = 0, ob = 0;
                for (int sidx = 0; sidx < p.nsAlo; ++sidx) oa |= (long long)((cc >> p.sAlo[sidx].src) & ((1 << p.sAlo[sidx].len) - 1)) << p.sAlo[sidx].dst;
                for (int sidx = 0; sidx < p.nsBlo; ++sidx) ob |= (long long)((cc >> p.sBlo[sidx].src) & ((1 << p.sBlo[sidx].len) - 1)) << p.sBlo[sidx].dst;
                auto ia = ra.find(oa);
                if (ia == ra.end()) { if ((int)ra.size() == kKredMaxRows) { ok = false; break; } kt.off_a[ra.size()] = oa; ia = ra.emplace(oa, (int)ra.size()).first; }
                auto ib = rb.find(ob);
                if (ib == rb.end()) { if ((int)rb.size() == kKredMaxRows) { ok = false; break; } kt.off_b[rb.size()] = ob; ib = rb.emplace(ob, (int)rb.size()).first; }
                kt.row_a[cc] = (unsigned char)ia->second; kt.row_b[cc] = (unsigned char)ib->second;
            }
            // the low kKredTileKBits bits of k must be ordinary k bits of both operands (they always are: K bits are in A and B)
            if (ok) {
                kt.n_rows_a = (int)ra.size(); kt.n_rows_b = (int)rb.size();
                size_t smem = (size_t)(kt.n_rows_a + kt.n_rows_b) * (kKredTileK + 1) * g->es();
                // outputs = the full grid rows(A) x rows(B): 2 x 2 outputs per thread, two-stage cp.async ring
                bool grid_ok = knob(0, "QXB_KRED_GRID", 1) != 0 && kt.n_rows_a % 2 == 0 && kt.n_rows_b % 2 == 0 &&
                               kt.n_rows_a * kt.n_rows_b == (1 << p.nC) && 2 * smem <= 160 * 1024;
                if (grid_ok) {
                    std::vector<int> seen(1 << p.nC, 0);
                    for (int cc = 0; cc < (1 << p.nC); ++cc) {
                        const int slot = kt.row_a[cc] * kt.n_rows_b + kt.row_b[cc];
                        if (seen[slot]++) { grid_ok = false; break; }
                        kt.grid_c[slot] = (unsigned char)cc;
                    }
                }
                if (grid_ok) smem *= 2;
                if (smem <= (grid_ok ? 160 : 96) * 1024) {
                    const int gt = (kt.n_rows_a % 4 == 0 && kt.n_rows_b % 4 == 0) ? 4 : 2;
                    n.func = grid_ok ? kreduce_grid_func(g->dtype, gt) : kreduce_tile_func(g->dtype);
                    n.kname = grid_ok ? "kreduce_grid" : "kreduce_tile";
                    n.smem = smem;
                    if (first_use(n.func))
                        CUDA_OK(cudaFuncSetAttribute(n.func, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024));
                    const long long nchunks = 1ll << (p.nK - kKredTileKBits);
                    // persistent grid = SMs x resident CTAs (a partial second wave would run at half occupancy: the chunks
                    // are dealt out statically)
                    static std::map<std::tuple<int, const void*, size_t>, int> resident;
                    auto key = std::make_tuple(g_device, n.func, smem);
                    auto it = resident.find(key);
                    if (it == resident.end()) {
                        int nb = 0;
                        CUDA_OK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, n.func, kThreads, smem));
                        it = resident.emplace(key, std::max(nb, 1)).first;
                    }
                    n.grid = dim3((unsigned)std::max<long long>(1, std::min<long long>(nchunks, (long long)g_num_sms * it->second)));
                    n.block = dim3(kThreads);
                    n.pre_zero_ptr = p.C;
                    n.pre_zero_bytes = (size_t)((C.amp ? c.n : 1) << C.span_bits) * g->es();
                    n.arg(p); n.arg(kt);
                    n.variant = c.variant_key; n.op = i;
                    const double uu = (double)p.U;
                    n.flops = 8.0 * op.macs_per_amp * uu;
                    n.bytes = (double)g->es() * (op.elems_a * (A.amp ? c.n : 1) + op.elems_b * (B.amp ? c.n : 1) + op.elems_c * uu);
                    return n;
                }
            }
        }
        (void)tiled;
        if (p.nK >= 16 && outputs * 2 <= (double)cap && !g->opts.no_gemm) {
            // huge K, a handful of outputs (root of a GEMM-shaped tree): split K over several blocks per output
            int sb = 0;
            while (outputs * std::ldexp(1.0, sb) < (double)cap && p.nK - sb > 13) ++sb;
            p.kc = sb;
            n.func = kreduce_split_func(g->dtype); n.kname = "kreduce_split";
            n.grid = dim3((unsigned)std::max<long long>(1, std::min<long long>((long long)std::ldexp(outputs, sb), cap)));
            n.pre_zero_ptr = p.C;                                // partial sums are combined with atomicAdd
            n.pre_zero_bytes = (size_t)((C.amp ? c.n : 1) << C.span_bits) * g->es();
        } else if (p.nK >= 12 && outputs <= 8192.0) {
            n.func = kreduce_block_func(g->dtype); n.kname = "kreduce_block";          // very long K, few outputs: a block per output
            n.grid = dim3((unsigned)std::max<long long>(1, std::min<long long>((long long)outputs, cap)));
        } else {
            n.func = kreduce_func(g->dtype); n.kname = "kreduce";
            const long long warps = (long long)outputs;
            n.grid = dim3((unsigned)std::max<long long>(1, std::min<long long>((warps * 32 + kThreads - 1) / kThreads, cap)));
        }
    } else {
        // broadcast-type node: both operands small per bitstring row, C much larger -> stage them in
        // shared memory once per row (otherwise every output re-reads them through L1/L2)
        const size_t stage = ((size_t(1) << A.span_bits) + (size_t(1) << B.span_bits)) * g->es();
        const void* sf = nullptr;
        // experiments: QXB_SMEM_RATIO (minimum |C| / (|A| + |B|) for staging), QXB_SMEM_MINHB (minimum hi bits)
        const double smem_ratio = [] { const char* e = getenv("QXB_SMEM_RATIO"); return e ? atof(e) : 1.5; }();
        const int smem_minhb = [] { const char* e = getenv("QXB_SMEM_MINHB"); return e ? atoi(e) : 1; }();
        // QXB_SMEM_SHARED=1: also stage nodes with one operand shared by all rows (it is loaded once per CTA and may
        // take up to 200 KB: one CTA per SM); the register tile may then be one-sided
        const int smem_shared = [] { const char* e = getenv("QXB_SMEM_SHARED"); return e ? atoi(e) : 0; }();
        const bool one_shared = smem_shared && (A.amp != B.amp) && p.U >= 4 * g_num_sms;
        const size_t cap_bytes = one_shared ? 200 * 1024 : 96 * 1024;
        const double amp_elems = (A.amp ? op.elems_a : 0) + (B.amp ? op.elems_b : 0);
        if (!g->opts.no_smem_stage && p.lob == 8 && (one_shared ? p.ma + p.nb >= 1 : (p.ma >= 1 && p.nb >= 1)) &&
            stage <= cap_bytes && p.hb >= (one_shared ? 0 : smem_minhb) &&
            A.lay.size() && B.lay.size() && op.elems_c >= (one_shared ? 0.25 * amp_elems : smem_ratio * (op.elems_a + op.elems_b)) &&
            op.elems_a == std::ldexp(1.0, A.span_bits) && op.elems_b == std::ldexp(1.0, B.span_bits) && p.U >= g_num_sms)
            sf = contract_smem_func(g->dtype, p.kc, p.ma, p.nb, p.kc == p.nK);
        // EXPERIMENT QXB_SMEM_TMA=1 (never run on hardware yet, see contract_tma_kernel): operand rows through 1-D TMA
        // bulk copies into a ring of stages; QXB_SMEM_TMA_RATIO = minimum |C| / (|A| + |B|)
        const int smem_tma = g->opts.smem_tma == 2 ? 0 : knob(g->opts.smem_tma, "QXB_SMEM_TMA", 1);   // default on since r2 (measured 13.1 vs 14.2 ms)
        const void* tf = nullptr;
        if (smem_tma && !g->opts.no_smem_stage && p.lob == 8 && p.ma + p.nb >= 1 && A.lay.size() && B.lay.size() &&
            (A.amp || B.amp) && p.U >= g_num_sms &&
            op.elems_a == std::ldexp(1.0, A.span_bits) && op.elems_b == std::ldexp(1.0, B.span_bits)) {
            const double tma_ratio = [] { const char* e = getenv("QXB_SMEM_TMA_RATIO"); return e ? atof(e) : 0.3; }();
            const size_t bytes_a = (size_t(1) << A.span_bits) * g->es(), bytes_b = (size_t(1) << B.span_bits) * g->es();
            const bool aligned = bytes_a % 16 == 0 && bytes_b % 16 == 0 && (uintptr_t)p.A % 16 == 0 && (uintptr_t)p.B % 16 == 0;
            const int fit = (int)std::min<size_t>(4, (200 * 1024 - 64) / stage);
            if (aligned && fit >= 2 && op.elems_c >= tma_ratio * (op.elems_a + op.elems_b))
                tf = contract_tma_func(g->dtype, p.kc, p.ma, p.nb, p.kc == p.nK);
            if (tf) tma_stages = fit;
        }
        // TMA ring kernel (qxb_rowprog.h): operand rows in by bulk copies, result row out by a bulk store, compute
        // from shared memory into shared memory with a free thread <-> element mapping
        if (knob(0, "QXB_RING", 1) != 0 && g->opts.ring != 1 && C.amp && (A.amp || B.amp) && p.U >= g_num_sms && p.nC >= 8 &&
            A.lay.size() && B.lay.size() && op.elems_a == std::ldexp(1.0, A.span_bits) && op.elems_b == std::ldexp(1.0, B.span_bits) &&
            A.span_bits <= 15 && B.span_bits <= 15 && p.nC <= 15 &&
            // measured (scripts/probe_ring.py, profiles/r2_summary.md): the ring wins on the big rows (68-72 KB per row:
            // 80-84 % of the HBM peak against 51-62 % for contract_kernel) and loses below ~48 KB per row, where its
            // per-row hand-shake (~3000 cycles with one CTA per SM) is longer than the row's HBM time
            ((double)(1ll << A.span_bits) + (double)(1ll << B.span_bits) + std::ldexp(1.0, p.nC)) * (double)g->es() >=
                (double)knob(0, "QXB_RING_MIN_ROW_BYTES", 48 * 1024)) {
            const size_t es = g->es();
            const int nA = 1 << A.span_bits, nB = 1 << B.span_bits, nCe = 1 << p.nC;
            const bool shA = !A.amp, shB = !B.amp;
            const bool aligned = (nA * es) % 16 == 0 && (nB * es) % 16 == 0 && (nCe * es) % 16 == 0 &&
                                 (uintptr_t)p.A % 16 == 0 && (uintptr_t)p.B % 16 == 0 && (uintptr_t)p.C % 16 == 0;
            Variant::RingDescs& rd = c.v->ring[i];
            if (aligned && !rd.tried) {
                rd.tried = true;
                RowPlanOptions ro;
                ro.bank_search_tiles = true;
                ro.bank_opt = g->opts.row_bank_opt != 1;
                ro.min_tt_bits = knob(0, "QXB_RING_MIN_TT", 8);
                ro.tile_reg_budget = knob(0, "QXB_RING_TILE_REGS", 100);
                ro.dmma = (g->opts.row_dmma == 2 || (g->opts.row_dmma == 0 && knob(0, "QXB_ROW_DMMA", 0) != 0));
                std::string why;
                std::vector<RowUnitDesc> descs = build_ring_descs(op, g->dtype, ro, why);
                if (!descs.empty()) {
                    rd.buf.reserve(descs.size() * sizeof(RowUnitDesc));
                    CUDA_OK(cudaMemcpy(rd.buf.p, descs.data(), descs.size() * sizeof(RowUnitDesc), cudaMemcpyHostToDevice));
                    rd.n_units = (int)descs.size();
                }
            }
            int stages = 0;
            if (aligned && rd.n_units > 0) {
                const size_t budget = 220 * 1024;
                for (int st = kRingMaxStages; st >= 2; --st)
                    if (ring_smem_bytes(rd.n_units, nA, nB, nCe, shA, shB, st, es) <= budget) { stages = st; break; }
            }
            if (stages >= 2) {
                RingLaunch R;
                memset(&R, 0, sizeof(R));
                R.descs = (const RowUnitDesc*)rd.buf.p; R.n_units = rd.n_units;
                R.A = p.A; R.B = p.B; R.C = p.C; R.sUA = p.sUA; R.sUB = p.sUB; R.sUC = p.sUC; R.U = p.U;
                R.nA = nA; R.nB = nB; R.nC = nCe; R.stages = stages;
                n.func = ring_func(g->dtype); n.kname = "ring";
                n.smem = ring_smem_bytes(rd.n_units, nA, nB, nCe, shA, shB, stages, es);
                n.grid = dim3((unsigned)std::max<long long>(1, std::min<long long>(p.U, (long long)g_num_sms)));
                n.block = dim3(kRowThreads);
                if (first_use(n.func))
                    CUDA_OK(cudaFuncSetAttribute(n.func, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
                n.arg(R);
                n.variant = c.variant_key; n.op = i;
                const double u = (double)p.U;
                n.flops = 8.0 * op.macs_per_amp * u;
                n.bytes = (double)g->es() * (op.elems_a * (A.amp ? c.n : 1) + op.elems_b * (B.amp ? c.n : 1) + op.elems_c * u);
                return n;
            }
        }
        if (tf) {
            p.aBits = A.span_bits; p.bBits = B.span_bits;
            n.func = tf; n.kname = "tma";
            n.smem = stage * (size_t)tma_stages + 64;
            n.grid = dim3((unsigned)std::max<long long>(1, std::min<long long>(p.U, (long long)g_num_sms)));
            if (first_use(tf))
                CUDA_OK(cudaFuncSetAttribute(tf, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
        } else if (sf) {
            p.aBits = A.span_bits; p.bBits = B.span_bits;
            n.func = sf; n.kname = "smem";
            n.smem = stage;
            n.grid = dim3((unsigned)std::max<long long>(1, std::min<long long>(p.U, (long long)g_num_sms * 2)));
            if (first_use(sf))
                CUDA_OK(cudaFuncSetAttribute(sf, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
        } else {
            // QXB_MINB=3 (experiment): the register allocation bounded for three resident CTAs per SM where it does not spill
            const int minb = [] { const char* e = getenv("QXB_MINB"); return e ? atoi(e) : 2; }();
            n.func = contract_func(g->dtype, p.kc, p.ma, p.nb, p.kc == p.nK, minb); n.kname = "contract";
            n.grid = dim3((unsigned)std::max<long long>(1, std::min(blocks, cap)));
        }
    }
    n.block = dim3(kThreads);
    n.arg(p);      // p is final here (aBits/bBits set above)
    if (tma_stages) n.arg(tma_stages);
    n.variant = c.variant_key; n.op = i;
    const double u = (double)p.U;
    n.flops = 8.0 * op.macs_per_amp * u;
    n.bytes = (double)g->es() * (op.elems_a * (A.amp ? c.n : 1) + op.elems_b * (B.amp ? c.n : 1) + op.elems_c * u);
    return n;
}

void launch_node(const Node& n, cudaStream_t st) {
    if (!n.func) { CUDA_OK(cudaMemsetAsync(n.ms_ptr, 0, n.ms_bytes, st)); return; }
    if (n.pre_zero_bytes) CUDA_OK(cudaMemsetAsync(n.pre_zero_ptr, 0, n.pre_zero_bytes, st));
    void* args[12];
    for (size_t a = 0; a < n.argoff.size(); ++a) args[a] = (void*)(n.argmem.data() + n.argoff[a]);
    CUDA_OK(cudaLaunchKernel(n.func, n.grid, n.block, args, n.smem, st));
}

// constant folding at compile time: serial launches of the const-phase nodes
void run_const_phase(const RunCtx& c) {
    Lowered& L = c.v->L;
    for (size_t i = 0; i < L.ops.size(); ++i)
        if (L.ops[i].phase == PH_CONST) launch_node(contract_node(c, (int)i), stream());
    CUDA_OK(cudaGetLastError());
}


// ------------------------------------------------------------------ row programs
// Host part: levels, units, arena plan, descriptors (qxb_rowplan.cpp).  Used when the saved tensor is a scalar,
// the root is produced in the chunk phase and the live set of one bitstring row fits shared memory.
void build_row_programs(qxb_graph* g, Variant& v) {
    v.rows_block = v.rows_chunk = false;
    if (knob(0, "QXB_ROWPROG", 1) == 0 || g->opts.row_programs == 1) return;
    RowPlanOptions o;
    o.bank_opt = g->opts.row_bank_opt != 1;
    o.min_tt_bits = knob(g->opts.row_min_tt_bits, "QXB_ROW_MIN_TT", 6);
    o.tile_reg_budget = knob(g->opts.row_tile_regs, "QXB_ROW_TILE_REGS", 100);
    o.alap = knob(0, "QXB_ROW_ALAP", 1) != 0;
    o.max_tile_bits = knob(0, "QXB_ROW_MAX_TILE", 4);
    o.stage_shared = knob(0, "QXB_ROW_STAGE", 1) != 0;
    o.dmma = (g->opts.row_dmma == 2 || (g->opts.row_dmma == 0 && knob(0, "QXB_ROW_DMMA", 0) != 0));
    o.max_arena_bytes = 227 * 1024 - (long long)row_fixed_smem_bytes(2048 + kRowWarps * kRowMaxLevels);   // descriptor buffers + slot table
    // block phase (slice-only nodes: hundreds of tiny contractions, pure launch latency as separate kernels)
    v.rp_block = build_row_program(v.L, PH_BLOCK, g->dtype, o);
    v.rows_block = v.rp_block.ok;
    // chunk phase: needs a scalar root produced in the chunk phase and a row that fits shared memory
    for (const auto& rm : v.L.root_modes) if (rm.nbits > 0) return;     // tensor-valued save: per-op path
    v.rp_chunk = build_row_program(v.L, PH_CHUNK, g->dtype, o);
    v.rows_chunk = v.rp_chunk.ok;
}

// Which phases of this call run as row programs.  The block program always (one 35 us launch instead of ~250
// kernels).  The chunk program wins while the step is launch-latency-bound (3 launches instead of ~350: 20 x faster
// at 10 bitstrings); with only two rows in flight per SM it is latency-bound itself, and above ~10^4 bitstrings
// the per-op kernels (HBM-bound, every SM streaming) are faster on a B200 (profiles/r2_summary.md).
bool chunk_as_rows(const qxb_graph* g, const Variant& v, int64_t n_amp) {
    if (!v.rows_chunk) return false;
    if (g->opts.row_programs == 2) return true;
    if (g->opts.row_programs == 3) return false;
    const bool has_block = std::any_of(v.L.ops.begin(), v.L.ops.end(), [](const LOp& op) { return op.phase == PH_BLOCK; });
    if (has_block && !v.rows_block) return false;
    return n_amp <= (int64_t)knob(g->opts.row_chunk_max_amps, "QXB_ROW_CHUNK_MAX_AMPS", 512);
}

// Device copies of the descriptors with the pointers of this block resolved (fixed slice values, arena bases).
Variant::RowDev& row_device_tables(const RunCtx& c) {
    Variant& v = *c.v;
    qxb_graph* g = c.g;
    const int k = (int)g->prog.vars.size();
    std::vector<int64_t> key(k, 0);
    for (int i = 0; i < k; ++i) if (!((v.L.free_mask >> i) & 1ull)) key[i] = c.fixed_vals[i];
    Variant::RowDev& rd = v.rowdev[key];
    if ((rd.descs_chunk.p || rd.descs_block.p) && rd.block_base == g->block_arena.p && rd.const_base == v.const_arena.p) return rd;
    rd.block_base = g->block_arena.p; rd.const_base = v.const_arena.p;
    auto fixed_off = [&](const LTensor& T) {
        long long off = 0;
        for (auto& f : T.fixed) off += (long long)c.fixed_vals[f.first] << f.second;
        return off;
    };
    auto resolve = [&](const RowProgramHost& rp, DevBuf& d_descs, DevBuf& d_slots, std::vector<int>& level_start) {
        std::vector<RowOp> ops = rp.ops;
        for (size_t j = 0; j < ops.size(); ++j) {
            RowOp& d = ops[j];
            if (rp.lop[j] < 0) {
                // copy pseudo-op of a staged operand: the whole tensor from its base (the consumers add the fixed offsets)
                const LTensor& A = v.L.tensors[rp.ref_a[j]];
                d.gA = (unsigned long long)(tensor_ptr(c, A) - fixed_off(A) * (long long)g->es());
                continue;
            }
            const LTensor &A = v.L.tensors[rp.ref_a[j]], &B = v.L.tensors[rp.ref_b[j]], &C = v.L.tensors[rp.ref_c[j]];
            if (rp.in_arena_a[j]) d.oA += (int)fixed_off(A); else { d.gA = (unsigned long long)tensor_ptr(c, A); d.oA = 0; }
            if (rp.in_arena_b[j]) d.oB += (int)fixed_off(B); else { d.gB = (unsigned long long)tensor_ptr(c, B); d.oB = 0; }
            if (!rp.in_arena_c[j]) { d.gC = (unsigned long long)tensor_ptr(c, C); d.oC = 0; }
        }
        RowDeviceTables t = build_row_tables(rp, ops);
        level_start = t.level_start;
        d_descs.reserve(std::max<size_t>(t.descs.size(), 1) * sizeof(RowUnitDesc));
        d_slots.reserve(std::max<size_t>(t.slots.size(), 1) * sizeof(uint16_t));
        CUDA_OK(cudaMemcpy(d_descs.p, t.descs.data(), t.descs.size() * sizeof(RowUnitDesc), cudaMemcpyHostToDevice));
        CUDA_OK(cudaMemcpy(d_slots.p, t.slots.data(), t.slots.size() * sizeof(uint16_t), cudaMemcpyHostToDevice));
    };
    if (v.rows_block) resolve(v.rp_block, rd.descs_block, rd.slots_block, rd.level_start_block);
    if (v.rows_chunk) resolve(v.rp_chunk, rd.descs_chunk, rd.slots_chunk, rd.level_start_chunk);
    rd.leaves.reserve(std::max<size_t>(v.rp_chunk.leaves.size(), 1) * sizeof(RowLeaf));
    if (v.rows_chunk && !v.rp_chunk.leaves.empty())
        CUDA_OK(cudaMemcpy(rd.leaves.p, v.rp_chunk.leaves.data(), v.rp_chunk.leaves.size() * sizeof(RowLeaf), cudaMemcpyHostToDevice));
    return rd;
}

Variant* get_variant(qxb_graph* g, uint64_t free_mask) {
    auto it = g->variants.find(free_mask);
    if (it != g->variants.end()) return it->second.get();
    std::unique_ptr<Variant> v(new Variant());
    v->L = lower(g->prog, free_mask, !g->opts.sum_at_root);
    // fused chain: pick it and make its ops contiguous BEFORE the memory plan (its intermediates never reach HBM; the
    // arena plan and the dependency edges must see the chain as one step)
    RowPlanOptions co;
    co.bank_search_tiles = knob(0, "QXB_CHAIN_BANK_TILES", 1) != 0;
    co.bank_opt = g->opts.row_bank_opt != 1;
    co.chain_side = g->opts.chain_side;
    const bool chain_on = knob(0, "QXB_CHAIN", 1) != 0 && g->opts.row_programs != 1 && g->opts.chain != 1;
    if (chain_on) {
        co.min_tt_bits = knob(0, "QXB_CHAIN_MIN_TT", 7);
        co.tile_reg_budget = knob(g->opts.row_tile_regs, "QXB_ROW_TILE_REGS", 100);
        co.chain_min_macs = (double)knob(0, "QXB_CHAIN_MIN_MACS", 2048);
        co.dmma = (g->opts.row_dmma == 2 || (g->opts.row_dmma == 0 && knob(0, "QXB_ROW_DMMA", 0) != 0));
        // two CTAs per SM: (228 KB / 2 - 1 KB reserved) minus descriptor buffers and slot table
        co.max_arena_bytes = knob(0, "QXB_CHAIN_ARENA_KB", 0) ? 1024ll * knob(0, "QXB_CHAIN_ARENA_KB", 0)
                                                              : (233472 / 2 - 1024) - (long long)row_fixed_smem_bytes(512);
        v->chain = select_chain(v->L, g->dtype, co);
        if (!v->chain.empty()) v->chain = make_contiguous(v->L, v->chain);
    }
    plan_memory(v->L);
    build_templates(*v, g->dtype, g->opts);
    build_row_programs(g, *v);
    if (!v->chain.empty()) {
        v->rp_chain = build_row_program(v->L, PH_CHUNK, g->dtype, co, &v->chain);
        if (!v->rp_chain.ok) v->chain.clear();
    }
    const size_t es = g->es();
    v->const_arena.reserve(std::max<int64_t>(v->L.const_elems, 2) * es);
    std::vector<OutLeafDesc> descs;
    for (int ti : v->L.output_leaves) {
        const LTensor& T = v->L.tensors[ti];
        descs.push_back(OutLeafDesc{T.offset, T.span_bits, (int)T.out_idx});
    }
    if (!descs.empty()) {
        v->outleaf_desc.reserve(descs.size() * sizeof(OutLeafDesc));
        CUDA_OK(cudaMemcpy(v->outleaf_desc.p, descs.data(), descs.size() * sizeof(OutLeafDesc), cudaMemcpyHostToDevice));
    }
    // constant folding: nodes that depend on no slice variable and no output bit run once
    std::vector<int64_t> zeros(g->prog.vars.size() + 1, 0);
    v->key = (int)g->variants.size();
    RunCtx c{g, v.get(), v->key, zeros.data(), 1};
    run_const_phase(c);
    CUDA_OK(cudaStreamSynchronize(stream()));
    Variant* raw = v.get();
    g->variants[free_mask] = std::move(v);
    return raw;
}

struct Block { uint64_t free_mask; std::vector<int64_t> vals; };

// aligned blocks of a contiguous range of linear slice ids (low variables free)
std::vector<Block> decompose(const Program& p, int64_t b, int64_t e) {
    const int k = (int)p.vars.size();
    std::vector<int64_t> place(k + 1, 1);
    for (int i = 0; i < k; ++i) place[i + 1] = place[i] * p.vars[i].dim;
    std::vector<Block> out;
    while (b < e) {
        int j = 0;
        while (j < k && b % place[j + 1] == 0 && b + place[j + 1] <= e) ++j;
        Block blk; blk.free_mask = low_mask(j); blk.vals.assign(k + 1, 0);
        slice_values(p, b, blk.vals.data());
        out.push_back(std::move(blk));
        b += place[j];
    }
    return out;
}

struct StepPlan {
    std::vector<Block> blocks;
    std::vector<Variant*> variants;
    int64_t chunk = 0;
};

int64_t hbm_budget(qxb_graph* g) {
    if (g->opts.hbm_budget_bytes > 0) return g->opts.hbm_budget_bytes;
    size_t free_b = 0, total_b = 0;
    CUDA_OK(cudaMemGetInfo(&free_b, &total_b));
    return (int64_t)(0.6 * (double)(free_b + g->block_arena.bytes + g->chunk_arena.bytes));
}

// Workspace (bytes) of one bitstring row + the block arena when the variables in `mask` are batched.
// Pure host computation (lowering + arena plan), cached per mask.
int64_t workspace_per_amp(qxb_graph* g, uint64_t mask) {
    auto it = g->ws_cache.find(mask);
    if (it != g->ws_cache.end()) return it->second;
    int64_t b = INT64_MAX;                      // a mask the lowering rejects (tensor too large) never fits
    try {
        Lowered L = lower(g->prog, mask, !g->opts.sum_at_root);
        plan_memory(L);
        const int64_t es = (int64_t)g->es();
        b = std::max<int64_t>(L.block_elems, 2) * es + std::max<int64_t>(L.chunk_elems_per_amp, 2) * es;
    } catch (const Error& e) {
        if (e.code != QXB_ERR_UNSUPP || mask == 0) throw;
    }
    g->ws_cache[mask] = b;
    return b;
}

// Batching every slice variable can need more HBM than there is (the sliced modes come back as
// batch bits).  Fix variables -- the one whose fixing shrinks the workspace most first -- and loop
// over their values until one bitstring row fits the budget.
void expand_blocks(qxb_graph* g, const Block& blk, int64_t budget, std::vector<Block>& out) {
    if (workspace_per_amp(g, blk.free_mask) <= budget || blk.free_mask == 0) { out.push_back(blk); return; }
    const int k = (int)g->prog.vars.size();
    int best = -1; int64_t best_ws = 0;
    for (int v = 0; v < k; ++v) {
        if (!((blk.free_mask >> v) & 1ull)) continue;
        const int64_t ws = workspace_per_amp(g, blk.free_mask & ~(1ull << v));
        if (best < 0 || ws < best_ws) { best = v; best_ws = ws; }
    }
    for (int64_t val = 0; val < g->prog.vars[best].dim; ++val) {
        Block sub = blk;
        sub.free_mask &= ~(1ull << best);
        sub.vals[best] = val;
        expand_blocks(g, sub, budget, out);
    }
}

// Everything that may allocate, synchronise or query the device happens here,
// before any node is built.
StepPlan prepare_step(qxb_graph* g, std::vector<Block> blocks, int64_t n_amp) {
    StepPlan sp;
    const size_t es = g->es();
    const int64_t budget = hbm_budget(g);
    for (const Block& b : blocks) expand_blocks(g, b, budget, sp.blocks);
    int64_t chunk = n_amp;
    if (g->opts.amp_batch > 0) chunk = std::min<int64_t>(chunk, g->opts.amp_batch);
    int64_t max_block = 2 * (int64_t)es, max_per_amp = 2 * (int64_t)es;
    for (const Block& blk : sp.blocks) {
        Variant* v = get_variant(g, blk.free_mask);
        sp.variants.push_back(v);
        const int64_t block_bytes = std::max<int64_t>(v->L.block_elems, 2) * es;
        // a row program keeps the chunk phase in shared memory: no HBM workspace per bitstring
        const bool rows = chunk_as_rows(g, *v, n_amp);
        const int64_t per_amp = rows ? 0 : std::max<int64_t>(v->L.chunk_elems_per_amp, 2) * es;
        if (rows) {
            max_block = std::max(max_block, block_bytes);
            g->stats.workspace_bytes = std::max<int64_t>(g->stats.workspace_bytes, block_bytes + (int64_t)v->const_arena.bytes);
            continue;
        }
        const int64_t fit = (budget - block_bytes) / per_amp;
        if (fit < 1)
            throw Error(QXB_ERR_MEM, "workspace for one bitstring (" + std::to_string(block_bytes + per_amp) +
                                         " bytes) exceeds the HBM budget (" + std::to_string(budget) + ")");
        chunk = std::min(chunk, fit);
        max_block = std::max(max_block, block_bytes);
        max_per_amp = std::max(max_per_amp, per_amp);
        g->stats.workspace_bytes = std::max<int64_t>(g->stats.workspace_bytes,
                                                      block_bytes + (int64_t)v->const_arena.bytes);
    }
    sp.chunk = chunk;
    bool moved = g->acc.reserve(sizeof(double) * 2 * n_amp * root_elems(g));
    moved |= g->block_arena.reserve(max_block);
    moved |= g->chunk_arena.reserve(max_per_amp * chunk);
    if (moved) { g->drop_step_graphs(); g->block_owner_variant = -1; }
    g->stats.workspace_bytes += max_per_amp * chunk;
    g->stats.amp_batch = chunk;
    g->stats.n_blocks = (int64_t)sp.blocks.size();
    return sp;
}


// The two fused launches of a block: op = -3 (block phase, one CTA over global memory), op = -2 (chunk phase, one CTA
// per bitstring row at a time, intermediates in shared memory).
constexpr int kOpRowChunk = -2, kOpRowBlock = -3;

Node row_node(const RunCtx& c, Variant::RowDev& rd, bool chunk, const uint8_t* d_bits, int64_t a0) {
    qxb_graph* g = c.g;
    Variant& v = *c.v;
    const RowProgramHost& rp = chunk ? v.rp_chunk : v.rp_block;
    RowLaunch P;
    memset(&P, 0, sizeof(P));
    P.descs = (const RowUnitDesc*)(chunk ? rd.descs_chunk.p : rd.descs_block.p);
    P.slots = (const uint16_t*)(chunk ? rd.slots_chunk.p : rd.slots_block.p);
    P.leaves = (const RowLeaf*)rd.leaves.p;
    P.bits = d_bits;
    P.acc = chunk ? (double*)g->acc.p : nullptr;
    P.amp0 = chunk ? a0 : 0;
    P.n_rows = chunk ? c.n : 1;
    P.scale = v.L.root_scale;
    P.n_levels = rp.n_levels;
    P.n_leaves = chunk ? (int)rp.leaves.size() : 0;
    P.n_outputs = g->prog.n_outputs;
    P.root_off = rp.root_off; P.root_span = rp.root_span;
    const std::vector<int>& ls = chunk ? rd.level_start_chunk : rd.level_start_block;
    for (int i = 0; i <= rp.n_levels; ++i) P.level_start[i] = ls[i];
    const size_t n_slots = (size_t)ls[rp.n_levels];
    Node n;
    n.func = rowprog_func(g->dtype);
    n.block = dim3(kRowThreads);
    P.slots_bytes = (int)row_slots_bytes(n_slots);
    P.timing = nullptr;
    if (chunk && g->opts.profile) {
        if (v.row_timing.reserve(sizeof(long long) * (kRowMaxLevels + 1)))
            CUDA_OK(cudaMemset(v.row_timing.p, 0, v.row_timing.bytes));
        P.timing = (long long*)v.row_timing.p;
    }
    n.smem = row_fixed_smem_bytes(n_slots) + (chunk ? (size_t)rp.arena_elems * g->es() : 0);
    if (first_use(n.func))
        CUDA_OK(cudaFuncSetAttribute(n.func, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    int per_sm = 1;
    if (chunk) {
        CUDA_OK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, n.func, kRowThreads, n.smem));
        per_sm = std::max(1, std::min(per_sm, knob(g->opts.row_ctas_per_sm, "QXB_ROW_CTAS", 2)));
    }
    n.grid = dim3(chunk ? (unsigned)std::max<long long>(1, std::min<long long>(c.n, (long long)g_num_sms * per_sm)) : 1u);
    n.arg(P);
    n.variant = c.variant_key; n.op = chunk ? kOpRowChunk : kOpRowBlock;
    const double rows = chunk ? (double)c.n : 1.0;
    n.flops = rp.flops_per_row * rows;
    n.bytes = (double)g->es() * (rp.elems_per_row_amp * rows + rp.elems_shared);   // same accounting as contract_node
    return n;
}

constexpr int kOpRowChain = -4;

// The fused chain of this variant for the rows of the current pass (c.n rows; chunk-arena pointers depend on it).
Node chain_node(const RunCtx& c) {
    qxb_graph* g = c.g;
    Variant& v = *c.v;
    const RowProgramHost& rp = v.rp_chain;
    std::vector<RowOp> ops = rp.ops;
    auto fixed_off = [&](const LTensor& T) {
        long long off = 0;
        for (auto& f : T.fixed) off += (long long)c.fixed_vals[f.first] << f.second;
        return off;
    };
    for (size_t j = 0; j < ops.size(); ++j) {
        RowOp& d = ops[j];
        if (rp.lop[j] < 0) {                              // staged input: whole tensor from its base (row 0 of this pass)
            const LTensor& A = v.L.tensors[rp.ref_a[j]];
            d.gA = (unsigned long long)(tensor_ptr(c, A) - fixed_off(A) * (long long)g->es());
            continue;
        }
        const LTensor &A = v.L.tensors[rp.ref_a[j]], &B = v.L.tensors[rp.ref_b[j]], &C = v.L.tensors[rp.ref_c[j]];
        d.oA += (int)fixed_off(A); d.oB += (int)fixed_off(B);          // every chain operand is staged in the arena
        if (!rp.in_arena_c[j]) { d.gC = (unsigned long long)tensor_ptr(c, C); d.oC = 0; }
    }
    RowDeviceTables t = build_row_tables(rp, ops);
    v.chain_dev.emplace_back(new Variant::ChainDev());
    Variant::ChainDev& cd = *v.chain_dev.back();
    cd.level_start = t.level_start;
    cd.descs.reserve(std::max<size_t>(t.descs.size(), 1) * sizeof(RowUnitDesc));
    cd.slots.reserve(std::max<size_t>(t.slots.size(), 1) * sizeof(uint16_t));
    CUDA_OK(cudaMemcpy(cd.descs.p, t.descs.data(), t.descs.size() * sizeof(RowUnitDesc), cudaMemcpyHostToDevice));
    CUDA_OK(cudaMemcpy(cd.slots.p, t.slots.data(), t.slots.size() * sizeof(uint16_t), cudaMemcpyHostToDevice));
    RowLaunch P;
    memset(&P, 0, sizeof(P));
    P.descs = (const RowUnitDesc*)cd.descs.p;
    P.slots = (const uint16_t*)cd.slots.p;
    P.amp0 = 0; P.n_rows = c.n; P.scale = 1.0;
    P.n_levels = rp.n_levels;
    for (int i = 0; i <= rp.n_levels; ++i) P.level_start[i] = cd.level_start[i];
    const size_t n_slots = (size_t)cd.level_start[rp.n_levels];
    P.slots_bytes = (int)row_slots_bytes(n_slots);
    if (g->opts.profile) {                                  // per-level cycles of CTA 0 (diagnostics, profile dump)
        if (v.row_timing.reserve(sizeof(long long) * (kRowMaxLevels + 1)))
            CUDA_OK(cudaMemset(v.row_timing.p, 0, v.row_timing.bytes));
        P.timing = (long long*)v.row_timing.p;
    }
    Node n;
    n.func = rowprog_func(g->dtype); n.kname = "chain";
    n.block = dim3(kRowThreads);
    n.smem = row_fixed_smem_bytes(n_slots) + (size_t)rp.arena_elems * g->es();
    if (first_use(n.func))
        CUDA_OK(cudaFuncSetAttribute(n.func, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    int per_sm = 1;
    CUDA_OK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, n.func, kRowThreads, n.smem));
    per_sm = std::max(1, std::min(per_sm, knob(g->opts.row_ctas_per_sm, "QXB_ROW_CTAS", 2)));
    n.grid = dim3((unsigned)std::max<long long>(1, std::min<long long>(c.n, (long long)g_num_sms * per_sm)));
    n.arg(P);
    n.variant = c.variant_key; n.op = kOpRowChain;
    n.flops = rp.flops_per_row * (double)c.n;
    n.bytes = (double)g->es() * (rp.elems_per_row_amp * (double)c.n + rp.elems_shared);    // algorithmic, as contract_node counts it
    return n;
}

std::vector<Node> build_step(qxb_graph* g, const StepPlan& sp, const uint8_t* d_bits, int64_t n_amp, void* d_out,
                             std::shared_ptr<Node>* hoist = nullptr, bool* writes_block = nullptr) {
    std::vector<Node> nodes;
    if (writes_block) *writes_block = false;
    {
        Node m; m.ms_ptr = g->acc.p; m.ms_bytes = sizeof(double) * 2 * n_amp * root_elems(g);
        nodes.push_back(std::move(m));
    }
    int sink = 0;                                     // last node of the serial spine
    for (size_t bi = 0; bi < sp.blocks.size(); ++bi) {
        const Block& blk = sp.blocks[bi];
        Variant* v = sp.variants[bi];
        Lowered& L = v->L;
        RunCtx c{g, v, v->key, blk.vals.data(), 1};
        const bool has_block = std::any_of(L.ops.begin(), L.ops.end(), [](const LOp& op) { return op.phase == PH_BLOCK; });
        const bool rows_chunk = chunk_as_rows(g, *v, n_amp);
        const bool rows_block = v->rows_block && has_block && g->opts.row_programs != 1;
        int block_node = -1;                          // the single-CTA block program, when the block phase runs as one
        if (writes_block && has_block) *writes_block = true;
        if (rows_block) {
            Variant::RowDev& rd = row_device_tables(c);
            Node b = row_node(c, rd, false, d_bits, 0);
            if (hoist && sp.blocks.size() == 1) {
                // kept out of the step: launched by run_blocks only when the block arena holds something else
                *hoist = std::make_shared<Node>(std::move(b));
                block_node = sink;                    // consumers of block-phase tensors start right after the memset
            } else {
                b.deps.push_back(sink);
                block_node = (int)nodes.size();
                nodes.push_back(std::move(b));
            }
        }
        if (rows_chunk) {
            // the whole block as two launches: block-phase program (one CTA), then one persistent kernel that takes
            // every bitstring row through the chunk phase in shared memory and adds the root into its accumulator
            Variant::RowDev& rd = row_device_tables(c);
            c.n = n_amp;
            Node r = row_node(c, rd, true, d_bits, 0);
            r.deps.push_back(block_node >= 0 ? block_node : sink);
            sink = (int)nodes.size();
            nodes.push_back(std::move(r));
            continue;
        }
        std::vector<int> node_of(L.ops.size(), -1);   // block-phase ops of this block
        for (size_t i = 0; i < L.ops.size(); ++i) {
            if (L.ops[i].phase != PH_BLOCK) continue;
            if (block_node >= 0) { node_of[i] = block_node; continue; }       // produced by the block program
            Node n = contract_node(c, (int)i);
            for (int d : L.ops[i].deps) if (node_of[d] >= 0) n.deps.push_back(node_of[d]);
            if (n.deps.empty()) n.deps.push_back(sink);
            node_of[i] = (int)nodes.size();
            nodes.push_back(std::move(n));
        }
        const LTensor& R = L.tensors[L.root];
        int root_op = -1;
        for (size_t i = 0; i < L.ops.size(); ++i) if (L.ops[i].c == L.root) root_op = (int)i;
        for (int64_t a0 = 0; a0 < n_amp; a0 += sp.chunk) {
            c.n = std::min(sp.chunk, n_amp - a0);
            int start = sink;
            if (!L.output_leaves.empty()) {
                Node n;
                n.func = outleaf_func(g->dtype);
                const long long nb = std::min<long long>(64, std::max<long long>(1, (c.n * 2 + 255) / 256));
                n.grid = dim3((unsigned)nb, (unsigned)L.output_leaves.size()); n.block = dim3(256);
                n.arg((void*)g->chunk_arena.p); n.arg((const OutLeafDesc*)v->outleaf_desc.p); n.arg((const unsigned char*)d_bits);
                n.arg((int)g->prog.n_outputs); n.arg((long long)a0); n.arg((long long)c.n);
                n.deps.push_back(sink);
                start = (int)nodes.size();
                nodes.push_back(std::move(n));
            }
            std::vector<int> cnode(L.ops.size(), -1);
            std::vector<char> in_chain(L.ops.size(), 0);
            const bool use_chain = !v->chain.empty() && c.n >= (long long)knob(0, "QXB_CHAIN_MIN_AMPS", 2 * g_num_sms);
            if (use_chain) for (int ci : v->chain) in_chain[ci] = 1;
            for (size_t i = 0; i < L.ops.size(); ++i) {
                if (L.ops[i].phase != PH_CHUNK) continue;
                if (in_chain[i]) {
                    if ((int)i != v->chain.front()) continue;
                    // the whole chain as one launch, at the position of its (contiguous) ops
                    Node n = chain_node(c);
                    for (int ci : v->chain) {
                        for (int d : L.ops[ci].deps) {
                            if (in_chain[d]) continue;
                            const int nd = L.ops[d].phase == PH_CHUNK ? cnode[d] : node_of[d];
                            if (nd >= 0) n.deps.push_back(nd);
                        }
                        if (L.tensors[L.ops[ci].a].is_output_leaf || L.tensors[L.ops[ci].b].is_output_leaf) n.deps.push_back(start);
                    }
                    if (n.deps.empty()) n.deps.push_back(start);
                    for (int ci : v->chain) cnode[ci] = (int)nodes.size();
                    nodes.push_back(std::move(n));
                    continue;
                }
                Node n = contract_node(c, (int)i);
                for (int d : L.ops[i].deps) {
                    const int nd = L.ops[d].phase == PH_CHUNK ? cnode[d] : node_of[d];
                    if (nd >= 0) n.deps.push_back(nd);
                }
                if (L.tensors[L.ops[i].a].is_output_leaf || L.tensors[L.ops[i].b].is_output_leaf || n.deps.empty())
                    n.deps.push_back(start);
                cnode[i] = (int)nodes.size();
                nodes.push_back(std::move(n));
            }
            Node r;
            bool open_root = false;
            for (const auto& rm : L.root_modes) open_root |= rm.nbits > 0;
            if (open_root) {
                // tensor-valued save (open network): gather the modes into Julia order, sum the open slice bits
                RootDesc d{};
                d.elems = L.root_elems; d.n_modes = (int)L.root_modes.size();
                for (int m = 0; m < d.n_modes; ++m) { d.ext[m] = (int)L.root_modes[m].ext; d.pos[m] = (unsigned char)L.root_modes[m].pos; }
                for (const LayEntry& e : R.lay) {
                    if (e.key >= 0) continue;
                    if (d.n_vseg == 16) throw Error(QXB_ERR_UNSUPP, "more than 16 batched slice variables open in a tensor-valued root");
                    d.vpos[d.n_vseg] = (unsigned char)e.pos; d.vbits[d.n_vseg] = (unsigned char)e.nbits; d.vtotal += e.nbits; ++d.n_vseg;
                }
                r.func = reduce_root_open_func(g->dtype);
                const long long rb = (c.n * d.elems + 255) / 256;
                r.grid = dim3((unsigned)std::max<long long>(1, std::min<long long>(rb, 148 * 8))); r.block = dim3(256);
                r.arg((const void*)tensor_ptr(c, R)); r.arg((long long)(R.amp ? (1ll << R.span_bits) : 0));
                r.arg((long long)c.n); r.arg((double)L.root_scale); r.arg((double*)g->acc.p); r.arg((long long)a0); r.arg(d);
            } else {
            r.func = reduce_root_func(g->dtype);
            long long rb = (c.n * 32 + 255) / 256;
            r.grid = dim3((unsigned)std::max<long long>(1, std::min<long long>(rb, 148 * 8))); r.block = dim3(256);
            r.arg((const void*)tensor_ptr(c, R)); r.arg((long long)(R.amp ? (1ll << R.span_bits) : 0)); r.arg((int)R.span_bits);
            r.arg((long long)c.n); r.arg((double)L.root_scale); r.arg((double*)g->acc.p); r.arg((long long)a0);
            }
            r.deps.push_back(sink);
            if (start != sink) r.deps.push_back(start);
            if (root_op >= 0) {
                const int nd = L.ops[root_op].phase == PH_CHUNK ? cnode[root_op] : node_of[root_op];
                if (nd >= 0) r.deps.push_back(nd);
            }
            // every node of this iteration must be done before the spine moves on
            for (size_t i = 0; i < L.ops.size(); ++i) if (cnode[i] >= 0 && L.tensors[L.ops[i].c].last_use == -1) r.deps.push_back(cnode[i]);
            sink = (int)nodes.size();
            nodes.push_back(std::move(r));
        }
    }
    Node f;
    f.func = finalize_func(g->dtype);
    const long long n_vals = (long long)n_amp * root_elems(g);
    f.grid = dim3((unsigned)std::max<long long>(1, std::min<long long>((n_vals + 255) / 256, 1024))); f.block = dim3(256);
    f.arg((const double*)g->acc.p); f.arg((void*)d_out); f.arg((long long)n_vals);
    f.deps.push_back(sink);
    nodes.push_back(std::move(f));
    return nodes;
}

void account(qxb_graph* g, const std::vector<Node>& nodes) {
    for (const Node& n : nodes) {
        if (n.func) g->stats.kernel_launches++;
        if (n.op >= 0 || n.op == kOpRowChunk || n.op == kOpRowBlock || n.op == kOpRowChain) {
            g->stats.contract_launches++; g->stats.flops += n.flops; g->stats.bytes += n.bytes;
        }
    }
}

// QXB_NCU_OPS=R474,R260,...: bracket exactly those ops with cudaProfilerStart/Stop (serial-launch mode), so that
// `ncu --profile-from-start off` captures the named contractions and nothing else
const std::set<std::string>& ncu_ops() {
    static const std::set<std::string> ops = [] {
        std::set<std::string> s;
        if (const char* e = getenv("QXB_NCU_OPS")) {
            std::stringstream ss(e);
            std::string tok;
            while (std::getline(ss, tok, ',')) if (!tok.empty()) s.insert(tok);
        }
        return s;
    }();
    return ops;
}

void launch_serial(qxb_graph* g, const std::vector<Node>& nodes) {
    cudaStream_t st = stream();
    for (const Node& n : nodes) {
        bool bracket = false;
        if (n.op >= 0 && !ncu_ops().empty())
            for (auto& kv : g->variants)
                if (kv.second->key == n.variant && ncu_ops().count(kv.second->L.ops[n.op].name)) bracket = true;
        if ((n.op == kOpRowChunk && ncu_ops().count("ROWPROG_CHUNK")) || (n.op == kOpRowBlock && ncu_ops().count("ROWPROG_BLOCK")) ||
            (n.op == kOpRowChain && ncu_ops().count("ROWPROG_CHAIN")))
            bracket = true;
        if (bracket) { CUDA_OK(cudaStreamSynchronize(st)); cudaProfilerStart(); }
        struct Stop { bool on; cudaStream_t s; ~Stop() { if (on) { cudaStreamSynchronize(s); cudaProfilerStop(); } } } stop{bracket, st};
        EventPair* ev = nullptr;
        const bool fused = n.op == kOpRowChunk || n.op == kOpRowBlock || n.op == kOpRowChain;
        if (g->opts.profile && (n.op >= 0 || fused)) {
            if (g->events_used == g->events.size()) {
                EventPair e{};
                CUDA_OK(cudaEventCreate(&e.a)); CUDA_OK(cudaEventCreate(&e.b));
                g->events.push_back(e);
            }
            ev = &g->events[g->events_used++];
            ev->variant = n.variant; ev->op = n.op;
            CUDA_OK(cudaEventRecord(ev->a, st));
        }
        launch_node(n, st);
        if (ev) {
            CUDA_OK(cudaEventRecord(ev->b, st));
            for (auto& kv : g->variants)
                if (kv.second->key == n.variant) {
                    OpProfile& pr = n.op == kOpRowChunk ? kv.second->prof_rows : n.op == kOpRowBlock ? kv.second->prof_block_rows
                                  : n.op == kOpRowChain ? kv.second->prof_chain : kv.second->prof[n.op];
                    pr.flops += n.flops; pr.bytes += n.bytes; pr.launches++; pr.kernel = n.kname;
                }
        }
    }
    CUDA_OK(cudaGetLastError());
}

cudaGraphExec_t instantiate(const std::vector<Node>& nodes) {
    cudaGraph_t graph = nullptr;
    CUDA_OK(cudaGraphCreate(&graph, 0));
    std::vector<cudaGraphNode_t> gn(nodes.size(), nullptr);
    try {
        for (size_t i = 0; i < nodes.size(); ++i) {
            const Node& n = nodes[i];
            std::vector<cudaGraphNode_t> deps;
            for (int d : n.deps) deps.push_back(gn[d]);
            std::sort(deps.begin(), deps.end());
            deps.erase(std::unique(deps.begin(), deps.end()), deps.end());
            if (!n.func) {
                cudaMemsetParams mp{};
                mp.dst = n.ms_ptr; mp.value = 0; mp.elementSize = 1; mp.width = n.ms_bytes; mp.height = 1; mp.pitch = n.ms_bytes;
                CUDA_OK(cudaGraphAddMemsetNode(&gn[i], graph, deps.data(), deps.size(), &mp));
            } else {
                if (n.pre_zero_bytes) {          // clear C first; the kernel then depends on the memset alone
                    cudaMemsetParams mp{};
                    mp.dst = n.pre_zero_ptr; mp.value = 0; mp.elementSize = 1; mp.width = n.pre_zero_bytes; mp.height = 1;
                    mp.pitch = n.pre_zero_bytes;
                    cudaGraphNode_t z = nullptr;
                    CUDA_OK(cudaGraphAddMemsetNode(&z, graph, deps.data(), deps.size(), &mp));
                    deps.assign(1, z);
                }
                void* args[12];
                for (size_t a = 0; a < n.argoff.size(); ++a) args[a] = (void*)(n.argmem.data() + n.argoff[a]);
                cudaKernelNodeParams kp{};
                kp.func = (void*)n.func; kp.gridDim = n.grid; kp.blockDim = n.block; kp.sharedMemBytes = (unsigned)n.smem;
                kp.kernelParams = args; kp.extra = nullptr;
                CUDA_OK(cudaGraphAddKernelNode(&gn[i], graph, deps.data(), deps.size(), &kp));
            }
        }
        cudaGraphExec_t exec = nullptr;
        CUDA_OK(cudaGraphInstantiate(&exec, graph, 0));
        cudaGraphDestroy(graph);
        return exec;
    } catch (...) {
        cudaGraphDestroy(graph);
        throw;
    }
}

void launch_step_graph(qxb_graph* g, const StepGraph& sg, cudaStream_t st) {
    if (sg.hoisted) {
        // (another stream gives no ordering against the launch that filled the arena: run it again)
        if (g->block_owner_variant != sg.block_variant || g->block_owner_vals != sg.block_vals || g->block_owner_stream != st) {
            launch_node(*sg.block_launch, st);        // stream order: before the graph that reads its results
            g->block_owner_variant = sg.block_variant;
            g->block_owner_vals = sg.block_vals;
            g->block_owner_stream = st;
        }
    } else if (sg.writes_block) {
        g->block_owner_variant = -1;                  // this step runs block phases of its own in the same arena
    }
    CUDA_OK(cudaGraphLaunch(sg.exec, st));
}

void run_blocks(qxb_graph* g, std::vector<Block> blocks, StepKey key, const uint8_t* d_bits, int64_t n_amp, void* d_out) {
    g->stats = qxb_stats{};
    g->events_used = 0;
    for (auto& kv : g->variants) {
        kv.second->prof.assign(kv.second->L.ops.size(), OpProfile{});
        kv.second->prof_rows = OpProfile{}; kv.second->prof_block_rows = OpProfile{}; kv.second->prof_chain = OpProfile{};
    }
    if (n_amp == 0) return;
    cudaStream_t st = stream();
    key.st = st;
    const bool use_graph = !g->opts.profile && !g->opts.no_cuda_graph;
    if (use_graph) {
        auto it = g->step_graphs.find(key);
        if (it != g->step_graphs.end()) {
            g->stats = it->second.stats;
            launch_step_graph(g, it->second, st);
            return;
        }
    }
    if (!use_graph) {
        // serial-launch modes rebuild the step on every call: the chain tables of the previous call are dead once it is done
        bool any = false;
        for (auto& kv : g->variants) any |= !kv.second->chain_dev.empty();
        if (any) {
            CUDA_OK(cudaStreamSynchronize(st));
            for (auto& kv : g->variants) {
                for (auto& cd : kv.second->chain_dev) { cd->descs.release(); cd->slots.release(); }
                kv.second->chain_dev.clear();
            }
        }
    }
    StepPlan sp = prepare_step(g, std::move(blocks), n_amp);
    static const bool hoist_on = [] { const char* e = getenv("QXB_HOIST_BLOCK"); return !e || atoi(e) != 0; }();
    const bool want_hoist = use_graph && hoist_on && sp.blocks.size() == 1;
    StepGraph sg;
    std::vector<Node> nodes = build_step(g, sp, d_bits, n_amp, d_out, want_hoist ? &sg.block_launch : nullptr, &sg.writes_block);
    if (sg.block_launch) {
        if (nodes.size() > 100000) throw Error(QXB_ERR_UNSUPP, "step too large");
        sg.hoisted = true;
        sg.block_variant = sp.variants[0]->key;
        sg.block_vals = sp.blocks[0].vals;
    }
    account(g, nodes);                                // the statistics of a REPLAYED step: without the hoisted block launch
    if (!use_graph || nodes.size() > 100000) { g->block_owner_variant = -1; launch_serial(g, nodes); return; }   // huge steps: not worth a graph
    if (g->step_graphs.size() >= 32) g->drop_step_graphs();
    sg.exec = instantiate(nodes);
    sg.stats = g->stats;
    auto ins = g->step_graphs.emplace(key, sg);
    launch_step_graph(g, ins.first->second, st);
}

void run_amplitudes(qxb_graph* g, const uint8_t* d_bits, int64_t n_amp, int64_t s0, int64_t s1, void* d_out) {
    if (!g->compiled) throw Error(QXB_ERR_STATE, "graph not compiled");
    const int64_t S = num_slices(g->prog);
    if (s0 < 0 || s1 > S || s0 > s1) throw Error(QXB_ERR_ARG, "slice range out of bounds");
    if (n_amp < 0) throw Error(QXB_ERR_ARG, "negative amplitude count");
    StepKey key{d_bits, d_out, n_amp, s0, s1, 0, {}, nullptr};
    run_blocks(g, decompose(g->prog, s0, s1), key, d_bits, n_amp, d_out);
}

void run_subspace(qxb_graph* g, const uint8_t* d_bits, int64_t n_amp, const int32_t* fvars, const int64_t* fvals,
                  int n_fixed, void* d_out) {
    if (!g->compiled) throw Error(QXB_ERR_STATE, "graph not compiled");
    if (n_amp < 0 || n_fixed < 0) throw Error(QXB_ERR_ARG, "negative count");
    const int k = (int)g->prog.vars.size();
    Block blk; blk.free_mask = low_mask(k); blk.vals.assign(k + 1, 0);
    for (int i = 0; i < n_fixed; ++i) {
        const int v = fvars[i];
        if (v < 0 || v >= k) throw Error(QXB_ERR_ARG, "fixed slice variable out of range");
        if (!((blk.free_mask >> v) & 1ull)) throw Error(QXB_ERR_ARG, "slice variable fixed twice");
        if (fvals[i] < 0 || fvals[i] >= g->prog.vars[v].dim) throw Error(QXB_ERR_ARG, "fixed value out of range");
        blk.free_mask &= ~(1ull << v);
        blk.vals[v] = fvals[i];
    }
    StepKey key{d_bits, d_out, n_amp, -1, -1, blk.free_mask, blk.vals, nullptr};
    std::vector<Block> blocks; blocks.push_back(std::move(blk));
    run_blocks(g, std::move(blocks), key, d_bits, n_amp, d_out);
}

void collect_profile(qxb_graph* g) {
    for (size_t i = 0; i < g->events_used; ++i) {
        EventPair& e = g->events[i];
        float ms = 0;
        if (cudaEventElapsedTime(&ms, e.a, e.b) == cudaSuccess) {
            for (auto& kv : g->variants) {
                if (kv.second->key != e.variant) continue;
                if (e.op == kOpRowChunk) kv.second->prof_rows.ms += ms;
                else if (e.op == kOpRowChain) kv.second->prof_chain.ms += ms;
                else if (e.op == kOpRowBlock) kv.second->prof_block_rows.ms += ms;
                else if (e.op >= 0 && e.op < (int)kv.second->prof.size()) kv.second->prof[e.op].ms += ms;
            }
        }
    }
    g->events_used = 0;
}

template <typename F>
int guard(F&& f) {
    try {
        f();
        return QXB_OK;
    } catch (const Error& e) {
        g_err = e.what();
        return e.code;
    } catch (const std::exception& e) {
        g_err = e.what();
        return QXB_ERR_ARG;
    }
}

}  // namespace

void qxb::set_last_error(const std::string& msg) { g_err = msg; }

// =============================================================== C ABI
extern "C" {

int qxb_version(void) { return 100; }

const char* qxb_last_error(void) { return g_err.c_str(); }

int qxb_init(int device) {
    return guard([&] { use_device(device); });
}

int qxb_shutdown(void) {
    return guard([&] {
        for (auto& kv : g_devs) if (kv.second.own) { cudaSetDevice(kv.first); cudaStreamDestroy(kv.second.own); }
        g_devs.clear();
        g_own_stream = nullptr;
        g_device = -1; g_use_ext = false; g_ext_stream = nullptr;
    });
}

int qxb_set_stream(void* s) {
    return guard([&] {
        g_ext_stream = (cudaStream_t)s;
        g_use_ext = true;
        if (s == nullptr) { g_use_ext = false; }
    });
}

int qxb_fma_peak(int dtype, double* tflops) {
    return guard([&] {
        if (!tflops) throw Error(QXB_ERR_ARG, "null argument");
        if (dtype != QXB_C32 && dtype != QXB_C64 && dtype != 2) throw Error(QXB_ERR_ARG, "dtype must be QXB_C32, QXB_C64 or 2 (packed FFMA2)");
        ensure_init();
        *tflops = fma_peak_tflops(dtype, g_num_sms, stream());
        CUDA_OK(cudaGetLastError());
    });
}

int qxb_device_synchronize(void) {
    return guard([&] { ensure_init(); CUDA_OK(cudaStreamSynchronize(stream())); });
}

int qxb_graph_create(qxb_graph** g, int dtype) {
    return guard([&] {
        if (!g) throw Error(QXB_ERR_ARG, "null graph pointer");
        if (dtype != QXB_C32 && dtype != QXB_C64) throw Error(QXB_ERR_ARG, "dtype must be QXB_C32 or QXB_C64");
        *g = new qxb_graph();
        (*g)->dtype = dtype;
    });
}

void qxb_graph_destroy(qxb_graph* g) { delete g; }

static void need_graph(qxb_graph* g) {
    if (!g) throw Error(QXB_ERR_ARG, "null graph");
    if (g->compiled) throw Error(QXB_ERR_STATE, "graph already compiled");
}

int qxb_graph_load(qxb_graph* g, const char* name, const char* label, const int64_t* dims, int rank) {
    return guard([&] {
        need_graph(g);
        if (!name || !label || rank < 0 || (rank > 0 && !dims)) throw Error(QXB_ERR_ARG, "bad load arguments");
        Cmd c; c.kind = CMD_LOAD; c.name = name; c.label = label; c.dims.assign(dims, dims + rank);
        add_cmd(g->prog, c);
    });
}

int qxb_graph_output(qxb_graph* g, const char* name, int64_t idx, int64_t dim) {
    return guard([&] {
        need_graph(g);
        if (!name) throw Error(QXB_ERR_ARG, "bad output arguments");
        Cmd c; c.kind = CMD_OUTPUT; c.name = name; c.idx = idx; c.dim = dim;
        add_cmd(g->prog, c);
    });
}

int qxb_graph_view(qxb_graph* g, const char* name, const char* target, const char* sym, int64_t pos, int64_t dim) {
    return guard([&] {
        need_graph(g);
        if (!name || !target || !sym) throw Error(QXB_ERR_ARG, "bad view arguments");
        Cmd c; c.kind = CMD_VIEW; c.name = name; c.a = target; c.label = sym; c.idx = pos; c.dim = dim;
        add_cmd(g->prog, c);
    });
}

int qxb_graph_ncon(qxb_graph* g, const char* out, const int64_t* cl, int nc, const char* a, const int64_t* al, int na,
                   const char* b, const int64_t* bl, int nb) {
    return guard([&] {
        need_graph(g);
        if (!out || !a || !b || nc < 0 || na < 0 || nb < 0) throw Error(QXB_ERR_ARG, "bad ncon arguments");
        Cmd c; c.kind = CMD_NCON; c.name = out; c.a = a; c.b = b;
        if (nc) c.cl.assign(cl, cl + nc);
        if (na) c.al.assign(al, al + na);
        if (nb) c.bl.assign(bl, bl + nb);
        add_cmd(g->prog, c);
    });
}

int qxb_graph_save(qxb_graph* g, const char* label, const char* name) {
    return guard([&] {
        need_graph(g);
        if (!label || !name) throw Error(QXB_ERR_ARG, "bad save arguments");
        Cmd c; c.kind = CMD_SAVE; c.name = label; c.a = name;
        add_cmd(g->prog, c);
    });
}

int qxb_graph_parse_dsl(qxb_graph* g, const char* text, size_t n) {
    return guard([&] {
        need_graph(g);
        if (!text) throw Error(QXB_ERR_ARG, "null program text");
        parse_dsl(g->prog, text, n);
    });
}

int qxb_graph_set_data(qxb_graph* g, const char* label, const void* data, const int64_t* dims, int rank) {
    return guard([&] {
        need_graph(g);
        if (!label || !data || rank < 0 || (rank > 0 && !dims)) throw Error(QXB_ERR_ARG, "bad set_data arguments");
        HostData h;
        int64_t n = 1;
        for (int i = 0; i < rank; ++i) {
            if (dims[i] < 1) throw Error(QXB_ERR_ARG, "bad dimension");
            h.dims.push_back(dims[i]); n *= dims[i];
        }
        const double* d = (const double*)data;
        h.v.resize(n);
        for (int64_t i = 0; i < n; ++i) h.v[i] = std::complex<double>(d[2 * i], d[2 * i + 1]);
        g->data[label] = std::move(h);
    });
}

int qxb_graph_num_outputs(const qxb_graph* g, int* n) {
    return guard([&] {
        if (!g || !n) throw Error(QXB_ERR_ARG, "null argument");
        ensure_analysed(const_cast<qxb_graph*>(g));
        *n = g->prog.n_outputs;
    });
}

int qxb_graph_root_dims(const qxb_graph* g, int* rank, int64_t* dims) {
    return guard([&] {
        if (!g || !rank) throw Error(QXB_ERR_ARG, "null argument");
        ensure_analysed(const_cast<qxb_graph*>(g));
        const std::vector<Mode>& ms = g->prog.defs[g->prog.root].modes;
        *rank = (int)ms.size();
        if (dims) for (size_t i = 0; i < ms.size(); ++i) dims[i] = ms[i].ext;
    });
}

int qxb_graph_num_slice_vars(const qxb_graph* g, int* k, int64_t* dims) {
    return guard([&] {
        if (!g || !k) throw Error(QXB_ERR_ARG, "null argument");
        ensure_analysed(const_cast<qxb_graph*>(g));
        *k = (int)g->prog.vars.size();
        if (dims) for (size_t i = 0; i < g->prog.vars.size(); ++i) dims[i] = g->prog.vars[i].dim;
    });
}

int qxb_graph_num_slices(const qxb_graph* g, int64_t* n) {
    return guard([&] {
        if (!g || !n) throw Error(QXB_ERR_ARG, "null argument");
        ensure_analysed(const_cast<qxb_graph*>(g));
        *n = num_slices(g->prog);
    });
}

int qxb_slice_values(const qxb_graph* g, int64_t s, int64_t* values) {
    return guard([&] {
        if (!g || !values) throw Error(QXB_ERR_ARG, "null argument");
        ensure_analysed(const_cast<qxb_graph*>(g));
        if (s < 0 || s >= num_slices(g->prog)) throw Error(QXB_ERR_ARG, "slice id out of range");
        slice_values(g->prog, s, values);
    });
}

int64_t qxb_graph_describe(qxb_graph* g, int n_free, char* buf, int64_t buflen) {
    int64_t need = 0;
    int rc = guard([&] {
        if (!g) throw Error(QXB_ERR_ARG, "null graph");
        ensure_analysed(g);
        const int k = (int)g->prog.vars.size();
        Lowered L = lower(g->prog, low_mask(n_free < 0 || n_free > k ? k : n_free), !g->opts.sum_at_root);
        plan_memory(L);
        std::string s = describe_json(g->prog, L);
        need = (int64_t)s.size() + 1;
        if (buf && buflen >= need) memcpy(buf, s.c_str(), need);
    });
    return rc == QXB_OK ? need : rc;
}

// test hook: the per-op launch templates (thread / register-tile / hi-bit split of every contraction, with the
// address maps composed for the kernel) exactly as build_templates hands them to contract_kernel, so that
// tests/template_emulator.py can replay the kernel's index arithmetic on the CPU.  Pure host logic.
int64_t qxb_debug_templates(qxb_graph* g, int n_free, void* buf, int64_t buflen) {
    int64_t need = 0;
    int rc = guard([&] {
        if (!g) throw Error(QXB_ERR_ARG, "null graph");
        ensure_analysed(g);
        const int k = (int)g->prog.vars.size();
        Variant v;
        v.L = lower(g->prog, low_mask(n_free < 0 || n_free > k ? k : n_free), !g->opts.sum_at_root);
        plan_memory(v.L);
        build_templates(v, g->dtype, g->opts);
        need = (int64_t)(v.tmpl.size() * sizeof(OpParams));
        if (buf && buflen >= need && need) memcpy(buf, v.tmpl.data(), (size_t)need);
    });
    return rc == QXB_OK ? need : rc;
}

int qxb_graph_replan(qxb_graph* g, int candidates, int64_t n_amp_model, double* given_bytes, double* new_bytes) {
    return guard([&] {
        need_graph(g);
        ensure_analysed(g);
        double a = 0, b = 0;
        replan(g->prog, candidates, 0x9E3779B97F4A7C15ull, (double)std::max<int64_t>(n_amp_model, 1), !g->opts.sum_at_root,
               &a, &b, (double)g->es());
        if (given_bytes) *given_bytes = a;
        if (new_bytes) *new_bytes = b;
    });
}

int qxb_graph_replan_ex(qxb_graph* g, int candidates, int64_t n_amp_model, int n_free, int64_t budget_bytes,
                        uint64_t seed, int* n_free_out, double* seconds_per_block, double* given_bytes, double* new_bytes) {
    return guard([&] {
        need_graph(g);
        ensure_analysed(g);
        double a = 0, b = 0, sec = 0;
        int nf = 0;
        replan(g->prog, candidates, seed ? seed : 0x9E3779B97F4A7C15ull, (double)std::max<int64_t>(n_amp_model, 1),
               !g->opts.sum_at_root, &a, &b, (double)g->es(), n_free, (double)budget_bytes, &nf, &sec);
        if (given_bytes) *given_bytes = a;
        if (new_bytes) *new_bytes = b;
        if (n_free_out) *n_free_out = nf;
        if (seconds_per_block) *seconds_per_block = sec;
    });
}

int64_t qxb_graph_program_text(qxb_graph* g, char* buf, int64_t buflen) {
    int64_t need = 0;
    int rc = guard([&] {
        if (!g) throw Error(QXB_ERR_ARG, "null graph");
        std::string s = program_text(g->prog);
        need = (int64_t)s.size() + 1;
        if (buf && buflen >= need) memcpy(buf, s.c_str(), need);
    });
    return rc == QXB_OK ? need : rc;
}

int qxb_graph_configure(qxb_graph* g, const qxb_options* opts) {
    return guard([&] {
        need_graph(g);
        if (!opts) throw Error(QXB_ERR_ARG, "null options");
        g->opts = *opts;
    });
}

int qxb_graph_compile(qxb_graph* g, const qxb_options* opts) {
    return guard([&] {
        need_graph(g);
        ensure_analysed(g);
        if (opts) g->opts = *opts;
        ensure_init();
        g->device = g_device;
        upload_leaves(g);
        g->compiled = true;
        try {
            get_variant(g, low_mask((int)g->prog.vars.size()));     // lower + fold constants for the all-free case
        } catch (const Error& e) {
            // heavily sliced programs (2^31-element tensors per slice) cannot batch every variable at all: the
            // variants with fewer batched variables are lowered on demand by the first call that needs them
            if (e.code != QXB_ERR_UNSUPP || g->prog.vars.empty()) { g->compiled = false; throw; }
        } catch (...) {
            g->compiled = false;
            throw;
        }
    });
}

int qxb_amplitudes_device(qxb_graph* g, const uint8_t* d_bits, int64_t n_amp, int64_t s0, int64_t s1, void* d_out) {
    return guard([&] {
        if (!g) throw Error(QXB_ERR_ARG, "null graph");
        if (n_amp > 0 && (!d_out || (!d_bits && g->prog.n_outputs > 0))) throw Error(QXB_ERR_ARG, "null buffer");
        ensure_init();
        if (g->compiled) use_device(g->device);
        run_amplitudes(g, d_bits, n_amp, s0, s1, d_out);
        if (g->opts.profile) {
            CUDA_OK(cudaStreamSynchronize(stream()));
            collect_profile(g);
        }
    });
}

int qxb_amplitudes(qxb_graph* g, const uint8_t* bits, int64_t n_amp, int64_t s0, int64_t s1, void* out) {
    return guard([&] {
        if (!g) throw Error(QXB_ERR_ARG, "null graph");
        if (n_amp < 0) throw Error(QXB_ERR_ARG, "negative amplitude count");
        if (n_amp == 0) return;
        if (!out || (!bits && g->prog.n_outputs > 0)) throw Error(QXB_ERR_ARG, "null buffer");
        if (!g->compiled) throw Error(QXB_ERR_STATE, "graph not compiled");
        ensure_init();
        use_device(g->device);
        const size_t nb = (size_t)n_amp * std::max(1, g->prog.n_outputs);
        cudaStream_t st = stream();
        g->d_bits.reserve(nb);
        g->d_out.reserve((size_t)n_amp * (size_t)root_elems(g) * g->es());
        if (g->prog.n_outputs > 0)
            CUDA_OK(cudaMemcpyAsync(g->d_bits.p, bits, (size_t)n_amp * g->prog.n_outputs, cudaMemcpyHostToDevice, st));
        run_amplitudes(g, (const uint8_t*)g->d_bits.p, n_amp, s0, s1, g->d_out.p);
        CUDA_OK(cudaMemcpyAsync(out, g->d_out.p, (size_t)n_amp * (size_t)root_elems(g) * g->es(), cudaMemcpyDeviceToHost, st));
        // The entries are validated on the host WHILE the device works (a pass over n_amp x n_outputs bytes: 0.3 ms for
        // the 6.4 MB of a 131072-bitstring call, which used to sit in front of the copy).  The kernels read any byte > 2
        // as '-', so nothing goes out of bounds; an invalid call still fails with QXB_ERR_ARG, `out` is then unspecified.
        const bool valid = bits_valid(bits, (size_t)n_amp * g->prog.n_outputs);
        CUDA_OK(cudaStreamSynchronize(st));
        if (!valid) throw Error(QXB_ERR_ARG, "bitstring entries must be 0, 1, 2 ('+') or 3 ('-')");
        if (g->opts.profile) collect_profile(g);
    });
}

int qxb_amplitudes_subspace(qxb_graph* g, const uint8_t* bits, int64_t n_amp, const int32_t* fixed_vars,
                            const int64_t* fixed_vals, int n_fixed, void* out, int on_device) {
    return guard([&] {
        if (!g) throw Error(QXB_ERR_ARG, "null graph");
        if (n_amp < 0) throw Error(QXB_ERR_ARG, "negative amplitude count");
        if (n_amp == 0) return;
        if (!out || (!bits && g->prog.n_outputs > 0) || (n_fixed > 0 && (!fixed_vars || !fixed_vals)))
            throw Error(QXB_ERR_ARG, "null buffer");
        if (!g->compiled) throw Error(QXB_ERR_STATE, "graph not compiled");
        ensure_init();
        use_device(g->device);
        cudaStream_t st = stream();
        if (on_device) {
            run_subspace(g, bits, n_amp, fixed_vars, fixed_vals, n_fixed, out);
            if (g->opts.profile) { CUDA_OK(cudaStreamSynchronize(st)); collect_profile(g); }
            return;
        }
        g->d_bits.reserve((size_t)n_amp * std::max(1, g->prog.n_outputs));
        g->d_out.reserve((size_t)n_amp * (size_t)root_elems(g) * g->es());
        if (g->prog.n_outputs > 0)
            CUDA_OK(cudaMemcpyAsync(g->d_bits.p, bits, (size_t)n_amp * g->prog.n_outputs, cudaMemcpyHostToDevice, st));
        run_subspace(g, (const uint8_t*)g->d_bits.p, n_amp, fixed_vars, fixed_vals, n_fixed, g->d_out.p);
        CUDA_OK(cudaMemcpyAsync(out, g->d_out.p, (size_t)n_amp * (size_t)root_elems(g) * g->es(), cudaMemcpyDeviceToHost, st));
        const bool valid = bits_valid(bits, (size_t)n_amp * g->prog.n_outputs);     // while the device works (see qxb_amplitudes)
        CUDA_OK(cudaStreamSynchronize(st));
        if (!valid) throw Error(QXB_ERR_ARG, "bitstring entries must be 0, 1, 2 ('+') or 3 ('-')");
        if (g->opts.profile) collect_profile(g);
    });
}

int qxb_partition_vars(qxb_graph* g, int n_parts, int32_t* vars_out, int* n_vars_out) {
    return guard([&] {
        if (!g || !n_vars_out) throw Error(QXB_ERR_ARG, "null argument");
        if (n_parts < 1) throw Error(QXB_ERR_ARG, "n_parts must be >= 1");
        ensure_analysed(g);
        std::vector<int> v = partition_vars(g->prog, n_parts, !g->opts.sum_at_root);
        *n_vars_out = (int)v.size();
        if (vars_out) for (size_t i = 0; i < v.size(); ++i) vars_out[i] = v[i];
    });
}

int qxb_graph_cost_bytes(qxb_graph* g, uint64_t free_mask, int64_t n_amp, double* bytes_out) {
    return guard([&] {
        if (!g || !bytes_out) throw Error(QXB_ERR_ARG, "null argument");
        ensure_analysed(g);
        Lowered L = lower(g->prog, free_mask, !g->opts.sum_at_root);
        *bytes_out = lowered_cost_bytes(L, (double)std::max<int64_t>(n_amp, 1), (double)g->es());
    });
}

int64_t qxb_graph_describe_mask(qxb_graph* g, uint64_t free_mask, char* buf, int64_t buflen) {
    int64_t need = 0;
    int rc = guard([&] {
        if (!g) throw Error(QXB_ERR_ARG, "null graph");
        ensure_analysed(g);
        Lowered L = lower(g->prog, free_mask, !g->opts.sum_at_root);
        plan_memory(L);
        std::string s = describe_json(g->prog, L);
        need = (int64_t)s.size() + 1;
        if (buf && buflen >= need) memcpy(buf, s.c_str(), need);
    });
    return rc == QXB_OK ? need : rc;
}

int qxb_last_stats(const qxb_graph* g, qxb_stats* st) {
    return guard([&] {
        if (!g || !st) throw Error(QXB_ERR_ARG, "null argument");
        *st = g->stats;
    });
}

int qxb_profile_dump(qxb_graph* g, const char* path) {
    return guard([&] {
        if (!g || !path) throw Error(QXB_ERR_ARG, "null argument");
        FILE* f = fopen(path, "w");
        if (!f) throw Error(QXB_ERR_ARG, std::string("cannot open ") + path);
        fprintf(f, "{\"dtype\":\"%s\",\"variants\":[", g->dtype == QXB_C32 ? "c32" : "c64");
        bool firstv = true;
        for (auto& kv : g->variants) {
            Variant& v = *kv.second;
            fprintf(f, "%s{\"n_free\":%d,\"free_mask\":%llu,\"ops\":[", firstv ? "" : ",", v.L.n_free,
                    (unsigned long long)kv.first);
            firstv = false;
            bool first = true;
            for (size_t i = 0; i < v.L.ops.size(); ++i) {
                const LOp& op = v.L.ops[i];
                const OpProfile& pr = v.prof[i];
                if (pr.launches == 0) continue;
                fprintf(f, "%s{\"name\":\"%s\",\"phase\":%d,\"nC\":%d,\"nK\":%d,\"batch_bits\":%d,\"m_bits\":%d,\"n_bits\":%d,"
                           "\"kernel\":\"%s\",\"launches\":%lld,\"flops\":%.6g,\"bytes\":%.6g,\"ms\":%.6g}",
                        first ? "" : ",", op.name.c_str(), (int)op.phase, op.nC, op.nK, op.n_batch, op.n_m, op.n_n,
                        pr.kernel, pr.launches, pr.flops, pr.bytes, pr.ms);
                first = false;
            }
            // the fused launches of a row-program variant: one pseudo-op per phase (the sums over the ops they cover)
            const OpProfile* fp[2] = {&v.prof_block_rows, &v.prof_rows};
            const char* fname[2] = {"ROWPROG_BLOCK", "ROWPROG_CHUNK"};
            for (int q = 0; q < 2; ++q) {
                if (fp[q]->launches == 0) continue;
                const RowProgramHost& rp = q ? v.rp_chunk : v.rp_block;
                std::string lc = "[]";
                if (q == 1 && v.row_timing.p) {
                    // cycles CTA 0 spent in each level, summed over every row it took since the graph was compiled
                    std::vector<long long> cyc(rp.n_levels, 0);
                    if (cudaMemcpy(cyc.data(), v.row_timing.p, sizeof(long long) * rp.n_levels, cudaMemcpyDeviceToHost) == cudaSuccess) {
                        lc = "[";
                        for (int l = 0; l < rp.n_levels; ++l) lc += (l ? "," : "") + std::to_string(cyc[l]);
                        lc += "]";
                    }
                }
                fprintf(f, "%s{\"name\":\"%s\",\"phase\":%d,\"fused_ops\":%d,\"levels\":%d,\"units\":%d,\"arena_bytes\":%lld,"
                           "\"level_cycles_cta0\":%s,"
                           "\"launches\":%lld,\"flops\":%.6g,\"bytes\":%.6g,\"ms\":%.6g}",
                        first ? "" : ",", fname[q], q ? 2 : 1, (int)rp.ops.size(), rp.n_levels, (int)rp.units.size(),
                        (long long)rp.arena_elems * (long long)g->es(), lc.c_str(), fp[q]->launches, fp[q]->flops, fp[q]->bytes, fp[q]->ms);
                first = false;
            }
            if (v.prof_chain.launches) {
                const RowProgramHost& rp = v.rp_chain;
                std::string names;
                for (int ci : v.chain) names += (names.empty() ? "\"" : ",\"") + v.L.ops[ci].name + "\"";
                // DRAM bytes the fused launch has to move per bitstring: its staged inputs and its results
                double io = 0;
                for (size_t j = 0; j < rp.ops.size(); ++j) {
                    if (rp.lop[j] < 0) { if (v.L.tensors[rp.ref_a[j]].amp) io += std::ldexp(1.0, v.L.tensors[rp.ref_a[j]].span_bits); }
                    else if (!rp.in_arena_c[j]) io += std::ldexp(1.0, v.L.tensors[rp.ref_c[j]].span_bits);
                }
                std::string lc = "[]";
                if (v.row_timing.p) {
                    std::vector<long long> cyc(rp.n_levels, 0);
                    if (cudaMemcpy(cyc.data(), v.row_timing.p, sizeof(long long) * rp.n_levels, cudaMemcpyDeviceToHost) == cudaSuccess) {
                        lc = "[";
                        for (int l = 0; l < rp.n_levels; ++l) lc += (l ? "," : "") + std::to_string(cyc[l]);
                        lc += "]";
                    }
                }
                fprintf(f, "%s{\"name\":\"ROWPROG_CHAIN\",\"phase\":2,\"kernel\":\"chain\",\"fused_ops\":%d,\"fused\":[%s],\"levels\":%d,"
                           "\"units\":%d,\"arena_bytes\":%lld,\"level_cycles_cta0\":%s,\"io_bytes_per_row\":%.0f,\"launches\":%lld,\"flops\":%.6g,\"bytes\":%.6g,\"ms\":%.6g}",
                        first ? "" : ",", (int)v.chain.size(), names.c_str(), rp.n_levels, (int)rp.units.size(),
                        (long long)rp.arena_elems * (long long)g->es(), lc.c_str(), io * (double)g->es(), v.prof_chain.launches,
                        v.prof_chain.flops, v.prof_chain.bytes, v.prof_chain.ms);
                first = false;
            }
            fprintf(f, "]}");
        }
        fprintf(f, "]}\n");
        fclose(f);
    });
}

}  // extern "C"

// =============================================================== several GPUs of one node behind the C ABI
// ("qxb_init(n_devices, ids) + one tiny reduce").  One host thread, one compiled replica of the
// graph per device.  A call splits its work two ways, the reference's two levels of parallelism
// (docs/src/users_guide.md:11-20; bin/qxrun.jl:40-46 `-m`, `-s`): the devices form n_devices / sub_comm_size groups;
// every group takes a contiguous share of the BITSTRINGS, the devices of a group split the SLICE range and their
// partial amplitudes are summed.  Disjoint bitstring shards land in disjoint ranges of the caller's buffer: with
// sub_comm_size = 1 there is nothing to reduce at all; inside a group the sum runs over sub_comm_size partial
// vectors of the group's amplitudes on the host (the "single tiny reduce": 16 bytes per amplitude and device).
struct HostPinned {
    void* p = nullptr; size_t bytes = 0;
    void reserve(size_t n) {
        if (n <= bytes) return;
        if (p) cudaFreeHost(p);
        p = nullptr; bytes = 0;
        if (cudaHostAlloc(&p, n, cudaHostAllocPortable) != cudaSuccess) { p = nullptr; throw Error(QXB_ERR_MEM, "cudaHostAlloc failed"); }
        bytes = n;
    }
    ~HostPinned() { if (p) cudaFreeHost(p); }
};
struct qxb_multi {
    std::vector<int> devs;
    std::vector<qxb_graph*> graphs;
    std::vector<std::unique_ptr<HostPinned>> h_bits, h_out;
    ~qxb_multi() { for (qxb_graph* g : graphs) delete g; }
};

extern "C" {

int qxb_multi_create(qxb_multi** m, const qxb_graph* proto, int n_devices, const int* device_ids, const qxb_options* opts) {
    return guard([&] {
        if (!m || !proto) throw Error(QXB_ERR_ARG, "null argument");
        if (proto->compiled) throw Error(QXB_ERR_STATE, "qxb_multi_create takes an uncompiled graph (it compiles one replica per device)");
        int avail = 0;
        cudaError_t e = cudaGetDeviceCount(&avail);
        if (e != cudaSuccess || avail == 0)
            throw Error(QXB_ERR_CUDA, "no CUDA device available (libqxb200 has no CPU fallback)");
        if (n_devices <= 0) n_devices = avail;
        std::unique_ptr<qxb_multi> mm(new qxb_multi());
        for (int i = 0; i < n_devices; ++i) {
            const int d = device_ids ? device_ids[i] : i;
            if (d < 0 || d >= avail) throw Error(QXB_ERR_ARG, "device index out of range");
            if (std::find(mm->devs.begin(), mm->devs.end(), d) != mm->devs.end()) throw Error(QXB_ERR_ARG, "device listed twice");
            mm->devs.push_back(d);
        }
        const int saved = g_device;
        for (int d : mm->devs) {
            use_device(d);
            std::unique_ptr<qxb_graph> g(new qxb_graph());
            g->dtype = proto->dtype;
            g->prog.cmds = proto->prog.cmds;             // the (possibly re-planned) program, statement by statement
            g->data = proto->data;
            g->opts = opts ? *opts : proto->opts;
            int rc = qxb_graph_compile(g.get(), nullptr);
            if (rc != QXB_OK) throw Error(rc, g_err);
            mm->graphs.push_back(g.release());
            mm->h_bits.emplace_back(new HostPinned());
            mm->h_out.emplace_back(new HostPinned());
        }
        if (saved >= 0) use_device(saved);
        *m = mm.release();
    });
}

void qxb_multi_destroy(qxb_multi* m) { delete m; }

int qxb_multi_num_devices(const qxb_multi* m, int* n) {
    return guard([&] {
        if (!m || !n) throw Error(QXB_ERR_ARG, "null argument");
        *n = (int)m->devs.size();
    });
}

int qxb_multi_amplitudes(qxb_multi* m, const uint8_t* bits, int64_t n_amp, int64_t s0, int64_t s1, int sub_comm_size, void* out) {
    return guard([&] {
        if (!m) throw Error(QXB_ERR_ARG, "null multi-GPU handle");
        if (n_amp < 0) throw Error(QXB_ERR_ARG, "negative amplitude count");
        if (n_amp == 0) return;
        qxb_graph* g0 = m->graphs[0];
        const int n_out = g0->prog.n_outputs;
        if (!out || (!bits && n_out > 0)) throw Error(QXB_ERR_ARG, "null buffer");
        const int64_t S = num_slices(g0->prog);
        if (s0 < 0 || s1 > S || s0 > s1) throw Error(QXB_ERR_ARG, "slice range out of bounds");
        if (root_elems(g0) != 1) throw Error(QXB_ERR_UNSUPP, "qxb_multi_amplitudes: tensor-valued save (open network)");
        {
            unsigned char seen = 0;
            const size_t nbits = (size_t)n_amp * n_out;
            for (size_t i = 0; i < nbits; ++i) seen |= bits[i];
            if (seen > 3) throw Error(QXB_ERR_ARG, "bitstring entries must be 0, 1, 2 ('+') or 3 ('-')");
        }
        const int n_dev = (int)m->devs.size();
        // auto: enough bitstrings -> every device its own shard (nothing to sum); a handful of amplitudes over many
        // slices (the Sycamore-like configuration) -> all devices share the bitstrings and split the slices
        int k = sub_comm_size > 0 ? sub_comm_size : (n_amp >= n_dev ? 1 : n_dev);
        if (k > n_dev || n_dev % k) throw Error(QXB_ERR_ARG, "sub_comm_size must divide the number of devices");
        const int groups = n_dev / k;
        const size_t es = g0->es();
        const int saved = g_device;
        const bool saved_ext = g_use_ext;
        g_use_ext = false;                                // the replicas run on their own devices' streams
        struct Part { int64_t a0, a1, b0, b1; };
        std::vector<Part> parts(n_dev);
        try {
            for (int i = 0; i < n_dev; ++i) {
                const int grp = i / k, r = i % k;
                Part& p = parts[i];
                p.a0 = n_amp * grp / groups; p.a1 = n_amp * (grp + 1) / groups;
                p.b0 = s0 + (s1 - s0) * r / k; p.b1 = s0 + (s1 - s0) * (r + 1) / k;
                const int64_t n = p.a1 - p.a0;
                if (n == 0 || (p.b1 == p.b0 && s1 > s0)) continue;
                qxb_graph* g = m->graphs[i];
                use_device(m->devs[i]);
                cudaStream_t st = stream();
                HostPinned &hb = *m->h_bits[i], &ho = *m->h_out[i];
                hb.reserve((size_t)n * std::max(1, n_out)); ho.reserve((size_t)n * es);
                g->d_bits.reserve((size_t)n * std::max(1, n_out));
                g->d_out.reserve((size_t)n * es);
                if (n_out > 0) {
                    memcpy(hb.p, bits + (size_t)p.a0 * n_out, (size_t)n * n_out);
                    CUDA_OK(cudaMemcpyAsync(g->d_bits.p, hb.p, (size_t)n * n_out, cudaMemcpyHostToDevice, st));
                }
                run_amplitudes(g, (const uint8_t*)g->d_bits.p, n, p.b0, p.b1, g->d_out.p);
                CUDA_OK(cudaMemcpyAsync(ho.p, g->d_out.p, (size_t)n * es, cudaMemcpyDeviceToHost, st));
            }
            // every device is busy now; collect in order
            if (k > 1) memset(out, 0, (size_t)n_amp * es);
            for (int i = 0; i < n_dev; ++i) {
                const Part& p = parts[i];
                const int64_t n = p.a1 - p.a0;
                if (n == 0 || (p.b1 == p.b0 && s1 > s0)) continue;
                use_device(m->devs[i]);
                CUDA_OK(cudaStreamSynchronize(stream()));
                char* dst = (char*)out + (size_t)p.a0 * es;
                if (k == 1) { memcpy(dst, m->h_out[i]->p, (size_t)n * es); continue; }
                if (g0->dtype == QXB_C32) {
                    float* o = (float*)dst; const float* v = (const float*)m->h_out[i]->p;
                    for (int64_t j = 0; j < 2 * n; ++j) o[j] += v[j];
                } else {
                    double* o = (double*)dst; const double* v = (const double*)m->h_out[i]->p;
                    for (int64_t j = 0; j < 2 * n; ++j) o[j] += v[j];
                }
            }
        } catch (...) {
            g_use_ext = saved_ext;
            if (saved >= 0) { try { use_device(saved); } catch (...) {} }
            throw;
        }
        g_use_ext = saved_ext;
        if (saved >= 0) use_device(saved);
    });
}

}  // extern "C"

extern "C" {

// test hook: the row program of one phase (qxb_rowprog.h) exactly as the executor would hand it to rowprog_kernel
// (unit descriptors and slot table of build_row_tables; tensors outside the arena keep a null pointer and are named
// by the reference arrays), serialised for tests/rowprog_emulator.py.  Pure host logic.  Layout: int32 header[8] =
// {ok, n_ops, n_descs, n_levels, n_leaves, arena_elems, root_off, root_span}, int32 n_slots, int32
// level_start[n_levels + 1] (in slots), uint16 slots[n_slots] (+ pad to 4 bytes), RowUnitDesc descs[n_descs],
// int32 desc_op[n_descs], RowLeaf leaves[n_leaves], int32 lop / ref_a / ref_b / ref_c / in_arena_a / in_arena_b /
// in_arena_c [n_ops] each.
int64_t qxb_debug_rowprog(qxb_graph* g, uint64_t free_mask, int phase, void* buf, int64_t buflen) {
    int64_t need = 0;
    int rc = guard([&] {
        if (!g) throw Error(QXB_ERR_ARG, "null graph");
        if (phase != PH_BLOCK && phase != PH_CHUNK && phase != 3)
            throw Error(QXB_ERR_ARG, "phase must be 1 (block), 2 (chunk) or 3 (the fused chain of the chunk phase)");
        ensure_analysed(g);
        Lowered L = lower(g->prog, free_mask, !g->opts.sum_at_root);
        std::vector<int> chain;
        RowPlanOptions co;
        co.bank_search_tiles = knob(0, "QXB_CHAIN_BANK_TILES", 1) != 0;
        co.bank_opt = g->opts.row_bank_opt != 1;
        co.chain_side = g->opts.chain_side;
        if (phase == 3) {
            co.min_tt_bits = knob(0, "QXB_CHAIN_MIN_TT", 7);
            co.max_arena_bytes = (233472 / 2 - 1024) - (long long)row_fixed_smem_bytes(512);
            chain = select_chain(L, g->dtype, co);
            if (!chain.empty()) chain = make_contiguous(L, chain);
        }
        plan_memory(L);
        RowPlanOptions o;
        o.bank_opt = g->opts.row_bank_opt != 1;
        o.min_tt_bits = knob(g->opts.row_min_tt_bits, "QXB_ROW_MIN_TT", 6);
        o.tile_reg_budget = knob(g->opts.row_tile_regs, "QXB_ROW_TILE_REGS", 100);
        o.alap = knob(0, "QXB_ROW_ALAP", 1) != 0;
        o.max_tile_bits = knob(0, "QXB_ROW_MAX_TILE", 4);
        o.stage_shared = knob(0, "QXB_ROW_STAGE", 1) != 0;
        o.dmma = co.dmma = (g->opts.row_dmma == 2 || (g->opts.row_dmma == 0 && knob(0, "QXB_ROW_DMMA", 0) != 0));
        o.max_arena_bytes = 227 * 1024 - (long long)row_fixed_smem_bytes(2048 + kRowWarps * kRowMaxLevels);
        RowProgramHost rp;
        if (phase == 3) {
            if (chain.empty()) { rp.ok = false; rp.why = "no chain worth fusing"; }
            else rp = build_row_program(L, PH_CHUNK, g->dtype, co, &chain);
        } else {
            rp = build_row_program(L, (Phase)phase, g->dtype, o);
        }
        if (!rp.ok) { set_last_error(rp.why); }
        std::vector<char> out;
        auto put = [&](const void* p, size_t n) { const char* c = (const char*)p; out.insert(out.end(), c, c + n); };
        if (!rp.ok) {
            int32_t hdr[8] = {0, 0, 0, 0, 0, 0, 0, 0};
            put(hdr, sizeof(hdr));
        } else {
            std::vector<RowOp> ops = rp.ops;
            for (size_t j = 0; j < ops.size(); ++j) {       // tensors outside the arena: base offset 0, pointer unresolved
                if (!rp.in_arena_a[j]) ops[j].oA = 0;
                if (!rp.in_arena_b[j]) ops[j].oB = 0;
                if (!rp.in_arena_c[j]) ops[j].oC = 0;
            }
            RowDeviceTables t = build_row_tables(rp, ops);
            int32_t hdr[8] = {1, (int)rp.ops.size(), (int)t.descs.size(), rp.n_levels, (int)rp.leaves.size(), rp.arena_elems,
                              rp.root_off, rp.root_span};
            put(hdr, sizeof(hdr));
            int32_t n_slots = (int32_t)t.slots.size();
            put(&n_slots, 4);
            std::vector<int32_t> ls(t.level_start.begin(), t.level_start.end());
            put(ls.data(), ls.size() * 4);
            std::vector<uint16_t> sl = t.slots;
            if (sl.size() % 2) sl.push_back(0);
            put(sl.data(), sl.size() * 2);
            put(t.descs.data(), t.descs.size() * sizeof(RowUnitDesc));
            auto puti = [&](const std::vector<int>& v) { std::vector<int32_t> w(v.begin(), v.end()); put(w.data(), w.size() * 4); };
            auto putc = [&](const std::vector<char>& v) { std::vector<int32_t> w(v.begin(), v.end()); put(w.data(), w.size() * 4); };
            puti(t.desc_op);
            put(rp.leaves.data(), rp.leaves.size() * sizeof(RowLeaf));
            puti(rp.lop); puti(rp.ref_a); puti(rp.ref_b); puti(rp.ref_c);
            putc(rp.in_arena_a); putc(rp.in_arena_b); putc(rp.in_arena_c);
            if (phase == 3) {                                // names of the fused ops, '\n'-separated, after the fixed layout
                std::string names;
                for (int ci : chain) names += L.ops[ci].name + "\n";
                put(names.data(), names.size());
            }
        }
        need = (int64_t)out.size();
        if (buf && buflen >= need) memcpy(buf, out.data(), out.size());
    });
    return rc == QXB_OK ? need : rc;
}

// test hook: the HBM arena plan of the chunk phase as the executor builds it for a variant WITH its fused chain
// (select_chain -> make_contiguous -> plan_memory).  Text, one record per line:
//   fused <first op> <last op>            (-1 -1: no chain)
//   op <index> <name> <a> <b> <c>         chunk-phase ops in launch order (tensor indices)
//   tensor <index> <offset> <elems> <amp> <leaf>     per-row element offset / size in the chunk arena (offset -1: not there)
int64_t qxb_debug_fused_plan(qxb_graph* g, uint64_t free_mask, char* buf, int64_t buflen) {
    int64_t need = 0;
    int rc = guard([&] {
        if (!g) throw Error(QXB_ERR_ARG, "null graph");
        ensure_analysed(g);
        Lowered L = lower(g->prog, free_mask, !g->opts.sum_at_root);
        RowPlanOptions co;
        co.bank_opt = g->opts.row_bank_opt != 1;
        co.chain_side = g->opts.chain_side;
        co.min_tt_bits = knob(0, "QXB_CHAIN_MIN_TT", 7);
        co.max_arena_bytes = (233472 / 2 - 1024) - (long long)row_fixed_smem_bytes(512);
        std::vector<int> chain = select_chain(L, g->dtype, co);
        if (!chain.empty()) chain = make_contiguous(L, chain);
        plan_memory(L);
        std::ostringstream o;
        o << "fused " << L.fused_first << " " << L.fused_last << "\n";
        std::set<int> ts;
        for (size_t i = 0; i < L.ops.size(); ++i) {
            const LOp& op = L.ops[i];
            if (op.phase != PH_CHUNK) continue;
            o << "op " << i << " " << op.name << " " << op.a << " " << op.b << " " << op.c << "\n";
            ts.insert(op.a); ts.insert(op.b); ts.insert(op.c);
        }
        for (int t : ts) {
            const LTensor& T = L.tensors[t];
            const bool here = T.phase == PH_CHUNK && !T.is_leaf;
            o << "tensor " << t << " " << (here ? (long long)T.offset : -1ll) << " " << std::max<int64_t>(2, int64_t(1) << T.span_bits) << " "
              << (T.amp ? 1 : 0) << " " << (T.is_leaf ? 1 : 0) << "\n";
        }
        const std::string out = o.str();
        need = (int64_t)out.size() + 1;
        if (buf && buflen >= need) memcpy(buf, out.c_str(), out.size() + 1);
    });
    return rc == QXB_OK ? need : rc;
}

// test hook: byte offset a tile-index bit contributes in the tcgen05 kernel's canonical K-major staging layout
int qxb_debug_tc5_smem_bit(int tile_bits, int bit) { return qxb::gemm_tc5_smem_bit(tile_bits, bit); }

// test hook: shared-memory layout table of the tensor-core GEMM kernels (tests/test_mma_layout.py)
int qxb_debug_mma_smem_bit(int dtype, int is_b, int tile_bits, int bit) {
    return qxb::gemm_mma_smem_bit(dtype, is_b != 0, tile_bits, bit);
}

}  // extern "C"
