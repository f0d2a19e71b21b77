// qxb200 -- host-side IR of a .qx program and its bit-level lowering.
//
// The op set is the one build_compute_graph emits
// (/root/reference/src/compute_graph/compute_graph.jl:15-98) with the textual
// form of /root/reference/docs/src/users_guide.md:93-164.
#pragma once
#include <complex>
#include <cstdint>
#include <map>
#include <set>
#include <stdexcept>
#include <string>
#include <vector>

namespace qxb {

struct Error : std::runtime_error {
    int code;
    Error(int c, const std::string& m) : std::runtime_error(m), code(c) {}
};

// message returned by qxb_last_error() on this thread (defined next to the C ABI, qxb_exec.cu)
void set_last_error(const std::string& msg);

// ----------------------------------------------------------------- DSL level
enum CmdKind { CMD_LOAD, CMD_OUTPUT, CMD_VIEW, CMD_NCON, CMD_SAVE };

struct Cmd {
    CmdKind kind;
    std::string name;              // defined symbol (load/output/view/ncon) or save label
    std::string a, b;              // view target / ncon operands / save source
    std::string label;             // load: data label; view: slice symbol
    std::vector<int64_t> dims;     // load dims
    std::vector<int64_t> cl, al, bl;
    int64_t idx = 0;               // output: 1-based qubit; view: 1-based mode
    int64_t dim = 0;               // output dim / view bond dim
};

struct HostData {
    std::vector<int64_t> dims;
    std::vector<std::complex<double>> v;   // column-major
};

struct SliceVar {
    std::string sym;
    int64_t dim;
    int nbits;                     // ceil(log2(dim)): free variables are padded to a power of two
};

// One DSL-level mode of a tensor.
struct Mode {
    int64_t ext;                   // extent seen by ncon label lists (1 once sliced)
    int64_t full_ext;              // extent of the stored mode
    int nbits;                     // ceil(log2(full_ext))
    int var;                       // -1, or the slice variable this mode was viewed with
};

enum TKind { T_LOAD, T_OUTPUT, T_VIEW, T_NCON };

struct TensorDef {
    TKind kind;
    std::string name;
    std::vector<Mode> modes;
    int leaf = -1;                 // load/output/view: index of the underlying load/output def
    std::string data_label;        // load
    int64_t out_idx = 0;           // output (1-based)
    int a = -1, b = -1;            // ncon operands (TensorDef indices)
    std::vector<int64_t> cl, al, bl;
    std::set<int> vars;            // slice variables the value depends on
    bool amp = false;              // depends on the output bitstring
    int uses = 0;
};

struct Program {
    std::vector<Cmd> cmds;
    std::vector<TensorDef> defs;
    std::map<std::string, int> by_name;
    std::vector<SliceVar> vars;    // v1..vk in numeric order
    std::map<std::string, int> var_by_sym;
    int n_outputs = 0;             // max output index
    int root = -1;                 // def index of the saved tensor
    bool analysed = false;
};

// parsing / construction ------------------------------------------------------
void parse_dsl(Program& p, const char* text, size_t n);
void add_cmd(Program& p, const Cmd& c);
// resolve names, modes, dependency sets; validates labels and extents
void analyse(Program& p);
int64_t num_slices(const Program& p);
void slice_values(const Program& p, int64_t s, int64_t* out);

// --------------------------------------------------------------- bit level
// A lowered tensor is a 2^n array; "entries" say which address bits belong to
// which logical mode.  key >= 0: DSL mode index of the def; key < 0: ~var.
struct LayEntry { int key; int nbits; int pos; };

enum Phase { PH_CONST = 0, PH_BLOCK = 1, PH_CHUNK = 2 };

struct Seg { uint8_t src, dst, len, pad; };

struct LTensor {
    int def = -1;
    std::vector<LayEntry> lay;
    int span_bits = 0;             // address span per amplitude (dense for intermediates)
    bool amp = false;
    Phase phase = PH_CONST;
    bool is_leaf = false;          // operand reads the uploaded leaf buffer directly
    std::string data_label;        // leaf: which buffer
    bool is_output_leaf = false;   // materialised per batch from the bitstrings
    int64_t out_idx = 0;
    std::vector<std::pair<int, int>> fixed;   // leaf: (var, bit position) of fixed variables
    int64_t offset = -1;           // element offset in its phase arena
    int first_use = -1, last_use = -1;        // op indices (within the lowered op list)
    bool persistent = false;       // consumed by a later phase
};

struct LOp {
    int c, a, b;                   // LTensor indices
    Phase phase;
    int nC = 0, nK = 0;
    int n_batch = 0, n_m = 0, n_n = 0;   // bit counts of the (batch, M, N) split; nK bits of K
    std::vector<Seg> segA, segB, segKA, segKB;
    std::vector<int> deps;         // ops that must finish first: producers of A/B and, after plan_memory,
                                   // the last readers of any arena region C overwrites (write-after-read)
    double macs_per_amp = 0;       // complex MACs per amplitude row (2^(nC+nK))
    double elems_a = 0, elems_b = 0, elems_c = 0;   // stored elements per amplitude row
    std::string name;
};

struct Lowered {
    uint64_t free_mask = 0;        // bit v set: slice variable v is batched; clear: fixed by the caller
    int n_free = 0;                // popcount(free_mask)
    bool early_sum = true;         // batched variables summed at the lowest covering node (else at the root)
    std::vector<LTensor> tensors;
    std::vector<LOp> ops;
    std::vector<int> output_leaves;          // LTensor indices materialised from the bitstrings
    int root = -1;
    std::vector<int> root_vars;              // batched variables still open in the root (summed by reduce_root)
    // open (tensor-valued) root: one entry per DSL mode of the saved tensor, in the order of its label list =
    // Julia dimension order of the result; empty for a closed network (scalar root)
    struct RootMode { int64_t ext; int nbits; int pos; };
    std::vector<RootMode> root_modes;
    int64_t root_elems = 1;                  // prod of the mode extents: values per bitstring in the result
    double root_scale = 1.0;                 // prod of extents of free vars the root does not depend on
    // arena sizes in elements: const, block, and chunk = per_amp * n_amp
    int64_t const_elems = 0, block_elems = 0, chunk_elems_per_amp = 0;
    // ops [fused_first, fused_last] run as ONE launch (a fused chain, qxb_rowplan.h make_contiguous): every row of the
    // batch reads the launch's inputs at its own time, so plan_memory releases no operand of these ops before the last one
    int fused_first = -1, fused_last = -1;
};

// Lower for a given set of free (batched) slice variables.
Lowered lower(const Program& p, uint64_t free_mask, bool early_sum = true);
inline uint64_t low_mask(int n) { return n >= 64 ? ~0ull : ((1ull << n) - 1ull); }
// Greedy choice of slice variables to FIX when sharding the slice space over n_parts ranks:
// at each step the variable whose fixing leaves the least work (bytes moved) per rank.
std::vector<int> partition_vars(const Program& p, int n_parts, bool early_sum = true);
double lowered_cost_bytes(const Lowered& L, double n_amp, double elem_bytes);
// Batch-aware re-planning of the ncon tree (qxb_replan.cpp).  Returns true when a cheaper
// program replaced p.cmds (leaf statements are kept verbatim).
// n_free: how many slice variables (v1..v_n_free) the plan is optimised and scored for as batched, the rest
// being fixed per block; -1 = all, -2 = choose the count that minimises blocks x modelled seconds under
// budget_bytes (largest node with its operands).  seconds_out: modelled seconds of one block.
bool replan(Program& p, int candidates, uint64_t seed, double n_amp, bool early_sum, double* given_bytes,
            double* new_bytes, double elem_bytes, int n_free = -1, double budget_bytes = 0, int* n_free_out = nullptr,
            double* seconds_out = nullptr);
std::string program_text(const Program& p);
// Plan arena offsets (fills LTensor::offset, the arena sizes and the write-after-read edges).
// Chunk-phase offsets are per bitstring row and scale with the batch at launch time.
void plan_memory(Lowered& L);
std::string describe_json(const Program& p, const Lowered& L);

inline int ceil_log2(int64_t x) { int b = 0; while (b < 62 && (int64_t(1) << b) < x) ++b; return b; }

}  // namespace qxb
