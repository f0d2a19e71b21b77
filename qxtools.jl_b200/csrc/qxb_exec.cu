// qxb200 -- executor and C ABI (include/qxb200.h).
//
// Replaces the executors behind the reference's hot path: QXTns.contract_tn!
// (call site /root/reference/src/simulation.jl:89) and the "Simulation" section of
// QXContexts.execute (call site /root/reference/bin/qxrun.jl:83-87):
//     for bitstring: for slice: run every ncon; acc += scalar
// Here the two loops are batch axes of each node's single launch (qxb_lower.cpp).
#include <cuda_runtime.h>

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <memory>
#include <sstream>
#include <tuple>

#include "../../include/qxb200.h"
#include "qxb_ir.h"
#include "qxb_kernels.cuh"

using namespace qxb;

namespace {

thread_local std::string g_err;
int g_device = -1;
cudaStream_t g_own_stream = nullptr;
cudaStream_t g_ext_stream = nullptr;
bool g_use_ext = false;
int g_num_sms = 148;

inline cudaStream_t stream() { return g_use_ext ? g_ext_stream : g_own_stream; }

#define CUDA_OK(expr)                                                                              \
    do {                                                                                           \
        cudaError_t e_ = (expr);                                                                   \
        if (e_ != cudaSuccess)                                                                     \
            throw Error(QXB_ERR_CUDA, std::string(#expr) + ": " + cudaGetErrorString(e_));         \
    } while (0)

void ensure_init() {
    if (g_device >= 0) return;
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess || n == 0)
        throw Error(QXB_ERR_CUDA, "no CUDA device available (libqxb200 has no CPU fallback)");
    CUDA_OK(cudaSetDevice(0));
    g_device = 0;
    CUDA_OK(cudaStreamCreateWithFlags(&g_own_stream, cudaStreamNonBlocking));
    cudaDeviceProp prop;
    CUDA_OK(cudaGetDeviceProperties(&prop, g_device));
    g_num_sms = prop.multiProcessorCount;
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

struct OpProfile { double ms = 0; double flops = 0; double bytes = 0; long long launches = 0; };

struct Variant {
    Lowered L;
    DevBuf const_arena;
    DevBuf outleaf_desc;
    std::vector<OpParams> tmpl;          // per op, pointers unset
    std::vector<OpProfile> prof;
};

struct EventPair { cudaEvent_t a, b; int variant, op; };

// A whole qxb_amplitudes step captured as a CUDA graph, keyed by everything the
// launches depend on.  Replay removes ~550 launch gaps per step.
struct StepKey {
    const void* bits; void* out; int64_t n_amp, s0, s1; cudaStream_t st;
    bool operator<(const StepKey& o) const {
        return std::tie(bits, out, n_amp, s0, s1, st) < std::tie(o.bits, o.out, o.n_amp, o.s0, o.s1, o.st);
    }
};
struct StepGraph { cudaGraphExec_t exec = nullptr; qxb_stats stats{}; };

}  // namespace

struct qxb_graph {
    int dtype = QXB_C64;
    Program prog;
    std::map<std::string, HostData> data;
    bool compiled = false;
    qxb_options opts{};
    std::map<std::string, DevBuf> leafbuf;
    std::map<int, std::unique_ptr<Variant>> variants;
    DevBuf block_arena, chunk_arena, acc, d_bits, d_out;
    qxb_stats stats{};
    std::vector<EventPair> events;
    size_t events_used = 0;
    std::map<StepKey, StepGraph> step_graphs;
    void drop_step_graphs() {
        for (auto& kv : step_graphs) if (kv.second.exec) cudaGraphExecDestroy(kv.second.exec);
        step_graphs.clear();
    }
    size_t es() const { return dtype == QXB_C32 ? 8 : 16; }
    ~qxb_graph() {
        for (auto& kv : leafbuf) kv.second.release();
        for (auto& kv : variants) { kv.second->const_arena.release(); kv.second->outleaf_desc.release(); }
        block_arena.release(); chunk_arena.release(); acc.release(); d_bits.release(); d_out.release();
        for (auto& e : events) { cudaEventDestroy(e.a); cudaEventDestroy(e.b); }
        drop_step_graphs();
    }
};

namespace {

void ensure_analysed(qxb_graph* g) {
    if (!g->prog.analysed) analyse(g->prog);
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
        // a rank-0 / all-ones-extent mismatch is tolerated only when shapes agree exactly
        if (h.dims != dims) {
            int64_t n1 = 1, n2 = 1;
            for (auto x : h.dims) n1 *= x;
            for (auto x : dims) n2 *= x;
            if (n1 != n2 || h.dims.size() != dims.size())
                throw Error(QXB_ERR_ARG, "load " + d.name + ": dims do not match the data of '" + d.data_label + "'");
            throw Error(QXB_ERR_ARG, "load " + d.name + ": dims do not match the data of '" + d.data_label + "'");
        }
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
void build_templates(Variant& v, int dtype) {
    v.tmpl.resize(v.L.ops.size());
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
        p.lob = std::min(nC, 8);
        // register tile: lowest M-only / N-only bits above the thread bits
        std::vector<int> mbits, nbits;
        for (int b = p.lob; b < nC; ++b) {
            if (mapA[b] >= 0 && mapB[b] < 0 && mbits.size() < 2) mbits.push_back(b);
            else if (mapB[b] >= 0 && mapA[b] < 0 && nbits.size() < 2) nbits.push_back(b);
        }
        // K chunk: keep (2^ma + 2^nb) * 2^kc operand loads (<= 64 registers) in flight
        const int budget_loads = dtype == QXB_C32 ? 32 : 16;
        int kc = std::min(op.nK, 3);
        while (kc > 0 && (int)(((1u << mbits.size()) + (1u << nbits.size())) << kc) > budget_loads) --kc;
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
        // thread-bit and hi-bit segment maps
        std::vector<std::pair<int, std::pair<int, int>>> alo, blo, ahi, bhi, chi;
        for (int b = 0; b < p.lob; ++b) {
            if (mapA[b] >= 0) alo.push_back({b, {mapA[b], 1}});
            if (mapB[b] >= 0) blo.push_back({b, {mapB[b], 1}});
        }
        int h = 0;
        for (int b = p.lob; b < nC; ++b) {
            if (is_tile[b]) continue;
            chi.push_back({h, {b, 1}});
            if (mapA[b] >= 0) ahi.push_back({h, {mapA[b], 1}});
            if (mapB[b] >= 0) bhi.push_back({h, {mapB[b], 1}});
            ++h;
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
    }
}

void run_phase(const RunCtx& c, Phase ph) {
    qxb_graph* g = c.g;
    Lowered& L = c.v->L;
    cudaStream_t st = stream();
    for (size_t i = 0; i < L.ops.size(); ++i) {
        const LOp& op = L.ops[i];
        if (op.phase != ph) continue;
        const LTensor &A = L.tensors[op.a], &B = L.tensors[op.b], &C = L.tensors[op.c];
        OpParams p = c.v->tmpl[i];
        p.A = tensor_ptr(c, A); p.B = tensor_ptr(c, B); p.C = tensor_ptr(c, C);
        p.sUA = A.amp ? (1ll << A.span_bits) : 0;
        p.sUB = B.amp ? (1ll << B.span_bits) : 0;
        p.sUC = C.amp ? (1ll << C.span_bits) : 0;
        p.U = C.amp ? (int)c.n : 1;
        p.tiles = (long long)p.U << p.hb;
        const int sub_bits = 8 - p.lob;
        long long blocks = (p.tiles + (1ll << sub_bits) - 1) >> sub_bits;
        const long long cap = (long long)g_num_sms * 8;
        const int grid = (int)std::max<long long>(1, std::min(blocks, cap));
        const double u = (double)p.U;
        const double flops = 8.0 * op.macs_per_amp * u;
        const double bytes = (double)g->es() * (op.elems_a * (A.amp ? c.n : 1) + op.elems_b * (B.amp ? c.n : 1) +
                                                op.elems_c * u);
        EventPair* ev = nullptr;
        if (g->opts.profile) {
            if (g->events_used == g->events.size()) {
                EventPair e{};
                CUDA_OK(cudaEventCreate(&e.a)); CUDA_OK(cudaEventCreate(&e.b));
                g->events.push_back(e);
            }
            ev = &g->events[g->events_used++];
            ev->variant = c.variant_key; ev->op = (int)i;
            CUDA_OK(cudaEventRecord(ev->a, st));
        }
        launch_contract(g->dtype, p, grid, st);
        if (ev) CUDA_OK(cudaEventRecord(ev->b, st));
        g->stats.kernel_launches++; g->stats.contract_launches++;
        g->stats.flops += flops; g->stats.bytes += bytes;
        if (g->opts.profile) {
            OpProfile& pr = c.v->prof[i];
            pr.flops += flops; pr.bytes += bytes; pr.launches++;
        }
    }
    CUDA_OK(cudaGetLastError());
}

Variant* get_variant(qxb_graph* g, int n_free) {
    auto it = g->variants.find(n_free);
    if (it != g->variants.end()) return it->second.get();
    std::unique_ptr<Variant> v(new Variant());
    v->L = lower(g->prog, n_free, !g->opts.sum_at_root);
    plan_memory(v->L, 1);
    build_templates(*v, g->dtype);
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
    RunCtx c{g, v.get(), n_free, zeros.data(), 1};
    const int prof = g->opts.profile;
    g->opts.profile = 0;
    run_phase(c, PH_CONST);
    g->opts.profile = prof;
    CUDA_OK(cudaStreamSynchronize(stream()));
    Variant* raw = v.get();
    g->variants[n_free] = std::move(v);
    return raw;
}

struct Block { int n_free; std::vector<int64_t> vals; };

std::vector<Block> decompose(const Program& p, int64_t b, int64_t e) {
    const int k = (int)p.vars.size();
    std::vector<int64_t> place(k + 1, 1);
    for (int i = 0; i < k; ++i) place[i + 1] = place[i] * p.vars[i].dim;
    std::vector<Block> out;
    while (b < e) {
        int j = 0;
        while (j < k && b % place[j + 1] == 0 && b + place[j + 1] <= e) ++j;
        Block blk; blk.n_free = j; blk.vals.assign(k + 1, 0);
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

// Everything that may allocate, synchronise or query the device happens here,
// outside stream capture.
StepPlan prepare_step(qxb_graph* g, int64_t n_amp, int64_t s0, int64_t s1) {
    StepPlan sp;
    sp.blocks = decompose(g->prog, s0, s1);
    size_t free_b = 0, total_b = 0;
    CUDA_OK(cudaMemGetInfo(&free_b, &total_b));
    const size_t es = g->es();
    int64_t budget = g->opts.hbm_budget_bytes > 0
                         ? g->opts.hbm_budget_bytes
                         : (int64_t)(0.6 * (double)(free_b + g->block_arena.bytes + g->chunk_arena.bytes));
    int64_t chunk = n_amp;
    if (g->opts.amp_batch > 0) chunk = std::min<int64_t>(chunk, g->opts.amp_batch);
    int64_t max_block = 2 * (int64_t)es, max_per_amp = 2 * (int64_t)es;
    for (const Block& blk : sp.blocks) {
        Variant* v = get_variant(g, blk.n_free);
        sp.variants.push_back(v);
        const int64_t block_bytes = std::max<int64_t>(v->L.block_elems, 2) * es;
        const int64_t per_amp = std::max<int64_t>(v->L.chunk_elems_per_amp, 2) * es;
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
    bool moved = g->acc.reserve(sizeof(double) * 2 * n_amp);
    moved |= g->block_arena.reserve(max_block);
    moved |= g->chunk_arena.reserve(max_per_amp * chunk);
    if (moved) g->drop_step_graphs();
    g->stats.workspace_bytes += max_per_amp * chunk;
    g->stats.amp_batch = chunk;
    g->stats.n_blocks = (int64_t)sp.blocks.size();
    return sp;
}

void issue_step(qxb_graph* g, const StepPlan& sp, const uint8_t* d_bits, int64_t n_amp, void* d_out) {
    cudaStream_t st = stream();
    CUDA_OK(cudaMemsetAsync(g->acc.p, 0, sizeof(double) * 2 * n_amp, st));
    for (size_t bi = 0; bi < sp.blocks.size(); ++bi) {
        const Block& blk = sp.blocks[bi];
        Variant* v = sp.variants[bi];
        Lowered& L = v->L;
        RunCtx c{g, v, blk.n_free, blk.vals.data(), 1};
        run_phase(c, PH_BLOCK);
        const LTensor& R = L.tensors[L.root];
        for (int64_t a0 = 0; a0 < n_amp; a0 += sp.chunk) {
            c.n = std::min(sp.chunk, n_amp - a0);
            if (!L.output_leaves.empty()) {
                launch_output_leaves(g->dtype, g->chunk_arena.p, (const OutLeafDesc*)v->outleaf_desc.p,
                                     (int)L.output_leaves.size(), d_bits, g->prog.n_outputs, a0, c.n, st);
                g->stats.kernel_launches++;
            }
            run_phase(c, PH_CHUNK);
            launch_reduce_root(g->dtype, tensor_ptr(c, R), R.amp ? (1ll << R.span_bits) : 0, R.span_bits, c.n,
                               L.root_scale, (double*)g->acc.p, a0, st);
            g->stats.kernel_launches++;
        }
    }
    launch_finalize(g->dtype, (const double*)g->acc.p, d_out, n_amp, st);
    g->stats.kernel_launches++;
}

void run_amplitudes(qxb_graph* g, const uint8_t* d_bits, int64_t n_amp, int64_t s0, int64_t s1, void* d_out) {
    if (!g->compiled) throw Error(QXB_ERR_STATE, "graph not compiled");
    const int64_t S = num_slices(g->prog);
    if (s0 < 0 || s1 > S || s0 > s1) throw Error(QXB_ERR_ARG, "slice range out of bounds");
    if (n_amp < 0) throw Error(QXB_ERR_ARG, "negative amplitude count");
    g->stats = qxb_stats{};
    g->events_used = 0;
    for (auto& kv : g->variants) kv.second->prof.assign(kv.second->L.ops.size(), OpProfile{});
    if (n_amp == 0) return;
    cudaStream_t st = stream();
    StepPlan sp = prepare_step(g, n_amp, s0, s1);
    const bool use_graph = !g->opts.profile && !g->opts.no_cuda_graph;
    if (!use_graph) {
        issue_step(g, sp, d_bits, n_amp, d_out);
        CUDA_OK(cudaGetLastError());
        return;
    }
    StepKey key{d_bits, d_out, n_amp, s0, s1, st};
    auto it = g->step_graphs.find(key);
    if (it == g->step_graphs.end()) {
        if (g->step_graphs.size() >= 32) g->drop_step_graphs();
        const qxb_stats pre = g->stats;
        CUDA_OK(cudaStreamBeginCapture(st, cudaStreamCaptureModeThreadLocal));
        cudaGraph_t graph = nullptr;
        try {
            issue_step(g, sp, d_bits, n_amp, d_out);
        } catch (...) {
            cudaStreamEndCapture(st, &graph);
            if (graph) cudaGraphDestroy(graph);
            g->stats = pre;
            throw;
        }
        CUDA_OK(cudaStreamEndCapture(st, &graph));
        StepGraph sg;
        cudaError_t e = cudaGraphInstantiate(&sg.exec, graph, 0);
        cudaGraphDestroy(graph);
        if (e != cudaSuccess) throw Error(QXB_ERR_CUDA, std::string("cudaGraphInstantiate: ") + cudaGetErrorString(e));
        sg.stats = g->stats;
        it = g->step_graphs.emplace(key, sg).first;
    }
    g->stats = it->second.stats;
    CUDA_OK(cudaGraphLaunch(it->second.exec, st));
}

void collect_profile(qxb_graph* g) {
    for (size_t i = 0; i < g->events_used; ++i) {
        EventPair& e = g->events[i];
        float ms = 0;
        if (cudaEventElapsedTime(&ms, e.a, e.b) == cudaSuccess) {
            auto it = g->variants.find(e.variant);
            if (it != g->variants.end() && e.op < (int)it->second->prof.size()) it->second->prof[e.op].ms += ms;
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

// =============================================================== C ABI
extern "C" {

int qxb_version(void) { return 100; }

const char* qxb_last_error(void) { return g_err.c_str(); }

int qxb_init(int device) {
    return guard([&] {
        int n = 0;
        cudaError_t e = cudaGetDeviceCount(&n);
        if (e != cudaSuccess || n == 0)
            throw Error(QXB_ERR_CUDA, "no CUDA device available (libqxb200 has no CPU fallback)");
        if (device < 0 || device >= n) throw Error(QXB_ERR_ARG, "device index out of range");
        CUDA_OK(cudaSetDevice(device));
        if (g_own_stream && g_device != device) { cudaStreamDestroy(g_own_stream); g_own_stream = nullptr; }
        g_device = device;
        if (!g_own_stream) CUDA_OK(cudaStreamCreateWithFlags(&g_own_stream, cudaStreamNonBlocking));
        cudaDeviceProp prop;
        CUDA_OK(cudaGetDeviceProperties(&prop, device));
        g_num_sms = prop.multiProcessorCount;
    });
}

int qxb_shutdown(void) {
    return guard([&] {
        if (g_own_stream) { cudaStreamDestroy(g_own_stream); g_own_stream = nullptr; }
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
        Lowered L = lower(g->prog, n_free, !g->opts.sum_at_root);
        plan_memory(L, 1);
        std::string s = describe_json(g->prog, L);
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
        upload_leaves(g);
        g->compiled = true;
        try {
            get_variant(g, (int)g->prog.vars.size());     // lower + fold constants for the all-free case
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
        const size_t nb = (size_t)n_amp * std::max(1, g->prog.n_outputs);
        for (size_t i = 0; i < (size_t)n_amp * g->prog.n_outputs; ++i)
            if (bits[i] > 3) throw Error(QXB_ERR_ARG, "bitstring entries must be 0, 1, 2 ('+') or 3 ('-')");
        cudaStream_t st = stream();
        g->d_bits.reserve(nb);
        g->d_out.reserve((size_t)n_amp * g->es());
        if (g->prog.n_outputs > 0)
            CUDA_OK(cudaMemcpyAsync(g->d_bits.p, bits, (size_t)n_amp * g->prog.n_outputs, cudaMemcpyHostToDevice, st));
        run_amplitudes(g, (const uint8_t*)g->d_bits.p, n_amp, s0, s1, g->d_out.p);
        CUDA_OK(cudaMemcpyAsync(out, g->d_out.p, (size_t)n_amp * g->es(), cudaMemcpyDeviceToHost, st));
        CUDA_OK(cudaStreamSynchronize(st));
        if (g->opts.profile) collect_profile(g);
    });
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
            fprintf(f, "%s{\"n_free\":%d,\"ops\":[", firstv ? "" : ",", kv.first);
            firstv = false;
            bool first = true;
            for (size_t i = 0; i < v.L.ops.size(); ++i) {
                const LOp& op = v.L.ops[i];
                const OpProfile& pr = v.prof[i];
                if (pr.launches == 0) continue;
                fprintf(f, "%s{\"name\":\"%s\",\"phase\":%d,\"nC\":%d,\"nK\":%d,\"batch_bits\":%d,\"m_bits\":%d,\"n_bits\":%d,"
                           "\"launches\":%lld,\"flops\":%.6g,\"bytes\":%.6g,\"ms\":%.6g}",
                        first ? "" : ",", op.name.c_str(), (int)op.phase, op.nC, op.nK, op.n_batch, op.n_m, op.n_n,
                        pr.launches, pr.flops, pr.bytes, pr.ms);
                first = false;
            }
            fprintf(f, "]}");
        }
        fprintf(f, "]}\n");
        fclose(f);
    });
}

}  // extern "C"
