// qxb200 -- contraction-tree search (qxb_treeopt.cpp); used by the re-planner (qxb_replan.cpp).
#pragma once
#include <cstdint>
#include <utility>
#include <vector>

namespace qxb {

// The tensor network behind a program: index classes (wires, bonds, hyper-edges, sliced hyper-edges,
// and one class for the bitstring axis shared by all output leaves).
struct TreeNet {
    int ncls = 0;
    int amp = -1;                               // class of the bitstring axis (-1: no output leaves)
    std::vector<double> wbits;                  // log2 of the stored extent per class (extents are padded to
                                                // powers of two; 0 for a slice variable the caller fixes)
    std::vector<int> total;                     // number of leaves carrying each class
    std::vector<std::vector<int>> leaf_ids;     // classes per leaf (amp included for output leaves)
    std::vector<char> leaf_var;                 // leaf depends on a slice variable (never constant-folded)
};

struct TreeCostModel {
    double elem_bytes = 16;                     // 8 for ComplexF32
    double bandwidth = 6.0e12;                  // B/s the streaming kernels reach on B200 (93 % of the measured peak)
    double flop_rate = 27e12;                   // flop/s of the tensor-core GEMM kernels on GEMM-shaped nodes (c64 DMMA 27e12, c32 ~40e12)
    // Every other node runs on the register-tiled SIMT kernel, which is bound by operand loads through L1:
    // (2^tm + 2^tn) / 2^(tm + tn) loads per complex MAC for a 2^tm x 2^tn register tile, at l1_bandwidth bytes/s --
    // 19.9 TB/s is the median over the L1-bound nodes of the tree-searched 7x7 plan (profiles/r1p_ops.md, fitted by
    // scripts/model_min_lob.py).  With one rate for all nodes (l1_bandwidth = 0: flops / flop_rate, the r1p model) the
    // search bought a few per cent of bytes with long-K untiled nodes that ran at a tenth of the HBM peak.
    double l1_bandwidth = 19.9e12;
    int gemm_min_mn_bits = 6, gemm_min_k_bits = 3;   // shape the tensor-core kernels take (qxb_exec.cu: build_templates)
    int thread_bits = 8;                        // C bits per bitstring row that stay thread bits (QXB_MIN_LOB)
    // compute seconds of a node with 2^macs_bits complex MACs whose C has c_bits bits per bitstring row, of which
    // m_bits come from A only and n_bits from B only; k_bits summed
    double compute_seconds(double macs_bits, double c_bits, double m_bits, double n_bits, double k_bits) const;
    double launch_s = 0.0;
    bool shared_reread = false;                 // charge an operand shared by all bitstrings once per bitstring row (experiment)
    int bisection_restarts = 0;                 // > 0: also seed the pool with recursive-bisection trees (Fiduccia-Mattheyses)
    double const_weight = 1.0;                  // share of a constant-folded node's cost that counts (1 = as if run once per step)
    // amp_bits: log2 of the bitstring batch when the operand / result carries that axis, else 0
    double time(double a_bits, double b_bits, double c_bits, double union_bits, double a_amp_bits = 0, double b_amp_bits = 0) const;
};

struct TreeReport { double seconds = 0, flops = 0, bytes = 0, max_bits = 0; };

// Randomised greedy construction (`restarts` trees) + the given seed plans, the best few refined by
// subtree reconfiguration.  Returns the modelled seconds of the best tree; `plan` lists its pairwise
// steps over tensor ids [0, n_leaves) + intermediates in creation order, `root` the last tensor.
double optimize_tree(const TreeNet& net, const TreeCostModel& cm, int restarts, int refine_rounds, uint64_t seed,
                     const std::vector<std::vector<std::pair<int, int>>>& seed_plans, const std::vector<int>& seed_roots,
                     std::vector<std::pair<int, int>>& plan, int& root, TreeReport* report);

// GPU-aware slicing: starting from `plan`, zero the extent of one index class at a time (the class whose slicing
// costs the least total work among those of the largest node) until no tensor exceeds 2^max_node_bits elements;
// the tree is re-refined after every choice.  Returns the chosen classes in order; plan/root are updated.
std::vector<int> slice_tree(TreeNet net, const TreeCostModel& cm, std::vector<std::pair<int, int>>& plan, int& root,
                            double max_node_bits, const std::vector<char>& sliceable, int max_slices, uint64_t seed,
                            TreeReport* report);

}  // namespace qxb
