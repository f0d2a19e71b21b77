/* libqxb200 -- C ABI of the B200-native executor for the QXTools contraction hot path.
 *
 * The reference has no FFI: its hot path sits behind three Julia-level seams, and
 * these entry points are what a `ccall` binding for each of them needs
 * (INTEGRATION.md shows the Julia side):
 *
 *   B1  single_amplitude(tnc, plan, amplitude)        /root/reference/src/simulation.jl:86-91
 *       -> QXTns.contract_tn!                         /root/reference/src/simulation.jl:89
 *   B2  build_compute_graph(tnc, plan, bond_groups)   /root/reference/src/compute_graph/compute_graph.jl:15-98
 *       -> ComputeGraph(root, tensors): Load/Output/View/Contract/Save commands
 *   B3  QXContexts.execute(dsl, input, param, output; ...)   /root/reference/bin/qxrun.jl:83-87
 *       -> the per-bitstring x per-slice contraction loop + reduction
 *
 * Conventions: plain C types only; every function returns 0 on success and a
 * negative code on error, with the message available from qxb_last_error()
 * (thread-local, owned by the library).  Names are NUL-terminated.  Dims and
 * labels are int64 arrays in Julia (column-major) order.  All tensor data
 * crossing the boundary is interleaved complex (re, im).  No exceptions cross
 * the boundary and no callbacks are taken.
 *
 * There is NO CPU fallback: every compute entry point fails with QXB_ERR_CUDA
 * when no CUDA device is usable.
 */
#ifndef QXB200_H
#define QXB200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define QXB_OK            0
#define QXB_ERR_ARG      -1   /* bad argument / malformed program            */
#define QXB_ERR_STATE    -2   /* call out of order (e.g. not compiled)        */
#define QXB_ERR_CUDA     -3   /* CUDA runtime error or no device              */
#define QXB_ERR_UNSUPP   -4   /* valid program the executor does not handle   */
#define QXB_ERR_MEM      -5   /* does not fit the HBM budget                  */

#define QXB_C32 0             /* ComplexF32 */
#define QXB_C64 1             /* ComplexF64 */

typedef struct qxb_graph qxb_graph;

typedef struct qxb_options {
    int64_t hbm_budget_bytes;  /* workspace budget per device; 0 = 60% of free memory            */
    int64_t amp_batch;         /* max bitstrings contracted per launch; 0 = as many as fit       */
    int32_t profile;           /* 1 = record a CUDA-event pair around every op (qxb_profile_dump) */
    int32_t no_cuda_graph;     /* 1 = launch every kernel directly instead of replaying a captured step */
    int32_t sum_at_root;       /* 1 = keep every batched slice variable open until the root and sum there;
                                  0 (default) = sum each one at the lowest node covering all its leaves   */
    int32_t no_smem_stage;     /* 1 = never use the shared-memory-staged kernel for broadcast-type nodes */
    int32_t no_gemm;           /* 1 = never use the tiled GEMM kernel for GEMM-shaped nodes               */
    int32_t gemm_mode;         /* GEMM-shaped nodes: 0 = auto (= 2), 1 = SIMT FMA kernel only, 2 = tensor cores wherever the
                                  tile shape allows: ComplexF32 -> tcgen05 / TMEM 3xTF32 kernel (2^7 x 2^6 x 2^4 tiles),
                                  else the mma.sync 3xTF32 kernel; ComplexF64 -> DMMA (tcgen05 has no FP64 kind);
                                  4 = mma.sync kernels only (no tcgen05) */
    /* ---- round 2: every kernel-selection knob is an option (0 = the library's default; the QXB_* environment
     *      variables of round 1 are still read, as overrides for experiments, only where the option is 0) ---- */
    int32_t row_programs;      /* row programs (csrc/qxb_rowprog.h: a whole phase of the tree as ONE persistent kernel).
                                  0 = auto: the block phase always, the chunk phase (a bitstring row's intermediates
                                  in shared memory) for calls of <= row_chunk_max_amps bitstrings; 1 = never;
                                  2 = both phases whenever the row fits; 3 = block phase only                      */
    int32_t min_lob;           /* 5..8: thread bits contract_kernel keeps (register tiles for small nodes); 0 = 6  */
    int32_t kc_regs_multi;     /* register budget of the K chunk, multi-chunk nodes; 0 = 160                      */
    int32_t kc_regs_one;       /* register budget of the K chunk, single-chunk nodes; 0 = 128                     */
    int32_t smem_tma;          /* TMA-staged contract_tma_kernel for broadcast-type nodes: 0 = auto (on), 1 = on, 2 = off */
    int32_t row_min_tt_bits;   /* row programs: keep >= 2^n thread-tiles per node when choosing the register tile; 0 = 6 */
    int32_t row_tile_regs;     /* row programs: registers for staged operands + accumulators; 0 = 100            */
    int32_t row_ctas_per_sm;   /* row programs: resident CTAs per SM; 0 = as many as the arena allows, at most 2  */
    int32_t ring;              /* 0 = auto: nodes whose operand and result rows are small dense per-bitstring rows run on
                                  the TMA ring kernel (bulk copies in, bulk store out, csrc/qxb_rowprog.h); 1 = never  */
    int32_t chain;             /* 0 = auto: the chain of dominant contractions (each one the only consumer of the previous
                                  result, small rows) runs as ONE row-program launch, its intermediates never reach HBM;
                                  1 = never                                                                          */
    int32_t row_dmma;          /* ComplexF64 nodes with >= 3 M-only and >= 3 N-only bits inside row programs, fused chains and
                                  the ring kernel on the FP64 tensor pipe (mma.sync.m8n8k4.f64, 8 x 8 tiles per warp):
                                  0 = auto (= off: measured 8 % slower than the SIMT register tiles on the headline chain,
                                  both are latency-bound, profiles/r2_summary.md), 1 = off, 2 = on                     */
    int32_t row_chunk_max_amps;/* auto mode: largest call (bitstrings) whose chunk phase runs as a row program; 0 = 512 */
    int32_t streaming;         /* "huge x tiny" nodes (one operand >= 2^20 elements, the other <= 64 KB, K <= 2^5, no batch bits)
                                  on bigsmall_kernel (csrc/qxb_kred.cu: a thread owns one position of the big operand and all its
                                  outputs, DRAM traffic |big| + |C| once): 0 = auto (on), 1 = off, 2 = on with the big operand
                                  staged by TMA bulk copies where its layout allows, 3 = on with packed FFMA2 accumulators
                                  (ComplexF32); 2 and 3 measured no faster than the default, profiles/r2_summary.md */
    int32_t row_bank_opt;      /* row programs / fused chain / ring kernel: choose the lane bits of a unit (and, inside a fused
                                  chain, the arena layout of the intermediates) by the shared-memory bank-conflict model:
                                  0 = auto (on), 1 = off (lanes = the lowest C bits outside the register tile, as lowered) */
    int32_t chain_side;        /* fused chain: also take the producers of the chain ops' other operands, up to this depth
                                  (they run in the levels' idle warps); 0 = off (measured slower on the headline plan) */
} qxb_options;

/* library */
int         qxb_version(void);
const char* qxb_last_error(void);
/* Select the CUDA device this process drives (one process per GPU). */
int         qxb_init(int device);
int         qxb_shutdown(void);
/* Launch on an existing CUDA stream (a cudaStream_t, e.g. torch's current stream); NULL = library stream. */
int         qxb_set_stream(void* cuda_stream);
int         qxb_device_synchronize(void);
/* Measured peak of the FMA pipe of the current device in TFLOP/s (QXB_C32: FFMA, QXB_C64: DFMA, 2: packed FFMA2; 2 flops per FMA, no
 * memory traffic): the roofline denominator for the fused launches whose intermediates live in shared memory. */
int         qxb_fma_peak(int dtype, double* tflops);

/* graph construction: one call per DSL statement, in DSL (post-order) order.
 * Replaces the Load/Output/View/Contract/Save command objects built at
 * compute_graph.jl:27,33,49,66,94. */
int  qxb_graph_create(qxb_graph** g, int dtype);
void qxb_graph_destroy(qxb_graph* g);
int  qxb_graph_load  (qxb_graph* g, const char* name, const char* data_label, const int64_t* dims, int rank);
int  qxb_graph_output(qxb_graph* g, const char* name, int64_t output_idx /*1-based qubit*/, int64_t dim);
int  qxb_graph_view  (qxb_graph* g, const char* name, const char* target, const char* slice_sym,
                      int64_t index_pos /*1-based mode*/, int64_t dim);
int  qxb_graph_ncon  (qxb_graph* g, const char* out, const int64_t* out_labels, int n_out,
                      const char* a, const int64_t* a_labels, int n_a,
                      const char* b, const int64_t* b_labels, int n_b);   /* n == 0 is the DSL's "0" scalar */
int  qxb_graph_save  (qxb_graph* g, const char* label, const char* name);
/* Same as the calls above, from the text of a .qx file (docs/src/users_guide.md:93-164). */
int  qxb_graph_parse_dsl(qxb_graph* g, const char* qx_text, size_t nbytes);
/* Leaf tensor values: ComplexF64, column-major (tensor_cache.jl:52-53,90-94).  Copied. */
int  qxb_graph_set_data(qxb_graph* g, const char* data_label, const void* c64_colmajor,
                        const int64_t* dims, int rank);

/* queries (pure host logic; usable without a GPU) */
int  qxb_graph_num_outputs(const qxb_graph* g, int* n_outputs);
/* Shape of the saved tensor: rank 0 for a closed network (one amplitude per bitstring).  An OPEN network -- built
 * with convert_to_tnc(...; no_output=true), which is how the reference's own tests drive contract_tn!
 * (test/test_contraction_planning.jl:50-61) -- saves a tensor: rank > 0 and dims (may be NULL) in Julia order;
 * qxb_amplitudes* then write prod(dims) values per bitstring, column-major, bitstring-major. */
int  qxb_graph_root_dims(const qxb_graph* g, int* rank, int64_t* dims);
/* k slice symbols v1..vk and their extents; dims may be NULL to query k only. */
int  qxb_graph_num_slice_vars(const qxb_graph* g, int* k, int64_t* dims);
int  qxb_graph_num_slices(const qxb_graph* g, int64_t* n_slices);
/* Linear slice id -> 0-based values of v1..vk (v1 fastest).  Bit-exact bookkeeping, see DESIGN.md. */
int  qxb_slice_values(const qxb_graph* g, int64_t slice_id, int64_t* values /*[k]*/);
/* Lowering report (JSON): per contraction the bit-level (batch, M, N, K) split, layouts,
 * dependency class, algorithmic flops/bytes.  `n_free` = how many low slice variables are
 * batched (the rest are fixed); -1 = all.  Returns the number of bytes needed (incl. NUL). */
int64_t qxb_graph_describe(qxb_graph* g, int n_free, char* buf, int64_t buflen);

/* Batch-aware re-planning (before compile): keep every load/output/view statement, recover the
 * tensor network behind the ncon tree and replace the tree by the cheapest of `candidates`
 * seeded min-fill orders of the UNSLICED network under the executor's cost model for batches of
 * n_amp_model bitstrings.  Exact re-association: the value of the program does not change.
 * The planner-side counterpart is contraction_scheme, src/contraction_planning.jl:219-299, which
 * plans for one slice at a time.  given/new_bytes (may be NULL) report the model cost. */
int  qxb_graph_replan(qxb_graph* g, int candidates, int64_t n_amp_model, double* given_bytes, double* new_bytes);
/* Same, for a run that batches only the first n_free slice variables (the others are fixed per block, as
 * qxb_amplitudes does when the all-batched workspace exceeds the HBM budget).  n_free = -1: all; -2: choose the
 * count that minimises blocks x modelled seconds with the largest node (and its operands) inside budget_bytes;
 * -3: GPU-aware slicing -- ADD slice variables (views on every leaf of the chosen index classes, exactly what
 * build_compute_graph emits for a bond group, compute_graph.jl:39-58) until the largest tensor of the searched
 * tree fits budget_bytes / 3; the new variables come after the file's own (v_{k+1}...).
 * Outputs (each may be NULL): the count used, the modelled seconds of one block, algorithmic bytes before/after. */
int  qxb_graph_replan_ex(qxb_graph* g, int candidates, int64_t n_amp_model, int n_free, int64_t budget_bytes,
                         uint64_t seed, int* n_free_out, double* seconds_per_block, double* given_bytes, double* new_bytes);
/* The program as .qx text (after re-planning: the re-planned one).  Returns bytes needed incl. NUL. */
int64_t qxb_graph_program_text(qxb_graph* g, char* buf, int64_t buflen);
/* Set options before qxb_graph_describe / qxb_graph_compile (compile(opts != NULL) overrides). */
int  qxb_graph_configure(qxb_graph* g, const qxb_options* opts);
/* Lower the program, upload leaves, fold constants, size the workspace. */
int  qxb_graph_compile(qxb_graph* g, const qxb_options* opts);

/* THE HOT PATH.  amplitude[a] = sum_{s in [slice_begin, slice_end)} root(bits[a], s).
 * bits: [n_amp][n_outputs] bytes, 0/1 (2 = '+', 3 = '-', basics.md:62), char i <-> qubit i.
 * out:  [n_amp] interleaved complex of the graph's dtype ([n_amp][prod(root dims)] for an open network,
 *       see qxb_graph_root_dims).
 * Host-pointer form: H2D of bits, compute, D2H of out, synchronised on return
 * (replaces the "Simulation" section of QXContexts.execute and contract_tn!).  The entries of `bits` are validated on
 * the host while the device works: a call with an entry > 3 returns QXB_ERR_ARG and leaves `out` unspecified. */
int  qxb_amplitudes(qxb_graph* g, const uint8_t* bits, int64_t n_amp,
                    int64_t slice_begin, int64_t slice_end, void* out);
/* Device-pointer form: bits and out already in HBM; asynchronous on the stream. */
int  qxb_amplitudes_device(qxb_graph* g, const uint8_t* d_bits, int64_t n_amp,
                           int64_t slice_begin, int64_t slice_end, void* d_out);

/* Slice SUB-SPACE form: every slice assignment in which the n_fixed variables fixed_vars[i]
 * (0-based: 0 = v1) take the values fixed_vals[i] and all other variables run over their full
 * extent.  This is the unit of multi-GPU sharding: the ranks fix different values of the same
 * variables and one all-reduce of `out` over the ranks gives the full sum.  on_device != 0:
 * bits/out are device pointers and the call is asynchronous. */
int  qxb_amplitudes_subspace(qxb_graph* g, const uint8_t* bits, int64_t n_amp, const int32_t* fixed_vars,
                             const int64_t* fixed_vals, int n_fixed, void* out, int on_device);
/* Which slice variables to fix when sharding over n_parts ranks: chosen greedily so that the
 * work left per rank is smallest (with dependency tracking, fixing a variable that few nodes
 * depend on saves nothing).  n_vars_out = 0 when the extents cannot factor n_parts (shard by
 * contiguous slice ranges instead).  Pure host logic. */
int  qxb_partition_vars(qxb_graph* g, int n_parts, int32_t* vars_out /*may be NULL*/, int* n_vars_out);

/* ---- several GPUs of one node behind the ABI (replaces the MPI layer of QXContexts.execute,
 * bin/qxrun.jl:40-46 `-m` / `-s`, docs/src/users_guide.md:11-20).  One host thread, one compiled replica per device.
 * qxb_multi_create takes an UNCOMPILED graph (built -- and, if wanted, re-planned -- through the calls above; it is
 * not modified and may be destroyed afterwards), clones it onto n_devices devices (device_ids NULL: 0..n-1;
 * n_devices <= 0: all visible) and compiles each replica there.
 * qxb_multi_amplitudes: host buffers as qxb_amplitudes.  The devices form n_devices / sub_comm_size groups: each
 * group takes a contiguous share of the bitstrings, the devices of a group split [slice_begin, slice_end) and their
 * partial amplitudes are summed (on the host: 16 bytes per amplitude and device).  sub_comm_size 0 = auto: 1 (disjoint
 * bitstring shards, nothing to reduce) when n_amp >= n_devices, else n_devices (all devices share the bitstrings). */
typedef struct qxb_multi qxb_multi;
int  qxb_multi_create(qxb_multi** m, const qxb_graph* g, int n_devices, const int* device_ids, const qxb_options* opts);
void qxb_multi_destroy(qxb_multi* m);
int  qxb_multi_num_devices(const qxb_multi* m, int* n);
int  qxb_multi_amplitudes(qxb_multi* m, const uint8_t* bits, int64_t n_amp, int64_t slice_begin, int64_t slice_end,
                          int sub_comm_size, void* out);
/* Cost model of one step: algorithmic bytes moved by the non-constant nodes when the variables in
 * free_mask are batched (the others fixed) and n_amp bitstrings are contracted.  Host logic only;
 * used to choose between sharding bitstrings and sharding slice variables over the ranks. */
int  qxb_graph_cost_bytes(qxb_graph* g, uint64_t free_mask, int64_t n_amp, double* bytes_out);
/* qxb_graph_describe for an arbitrary set of batched variables (bit v of free_mask = v(v+1)). */
int64_t qxb_graph_describe_mask(qxb_graph* g, uint64_t free_mask, char* buf, int64_t buflen);

/* Counters of the last qxb_amplitudes* call. */
typedef struct qxb_stats {
    int64_t kernel_launches;     /* kernels of this library launched                   */
    int64_t contract_launches;   /* of which the contraction kernel                    */
    double  flops;               /* algorithmic real flops executed (8 per complex MAC) */
    double  bytes;               /* algorithmic bytes: s*(|A|+|B|+|C|) summed over launches */
    int64_t workspace_bytes;     /* arena high-water mark                              */
    int64_t amp_batch;           /* bitstrings per batch actually used                 */
    int64_t n_blocks;            /* aligned slice blocks the range was split into      */
} qxb_stats;
int  qxb_last_stats(const qxb_graph* g, qxb_stats* st);
/* Per-op table (name, shape, flops, bytes, ms) of the last profiled call, as JSON. */
int  qxb_profile_dump(qxb_graph* g, const char* json_path);

/* ---- the file seam (B3): data files, parameter files, a whole triple --------------------------------
 * JLD2 (HDF5-subset) reader/writer for the `.jld2` files of a simulation triple: one dataset per data label
 * holding the N-d ComplexF64 array (tensor_cache.jl:90-106), and the results file bin/qxrun.jl -o names
 * (docs/src/distributed.md:30-33).  Handles what JLD2.jl emits for numeric arrays (512-byte header, superblock
 * v2, OHDR v2, link messages, contiguous/compact layout, committed compound {re, im}) and files of the HDF5 C
 * library (superblock v0/v1, OHDR v1, symbol-table groups); chunked/compressed datasets are QXB_ERR_UNSUPP. */
typedef struct qxb_jld2 qxb_jld2;
#define QXB_JLD2_MAX_RANK 32
#define QXB_ELEM_C64    0     /* {re: f64, im: f64} */
#define QXB_ELEM_C32    1     /* {re: f32, im: f32} */
#define QXB_ELEM_F64    2
#define QXB_ELEM_F32    3
#define QXB_ELEM_INT    4     /* elem_size bytes */
#define QXB_ELEM_STRING 5     /* fixed length, elem_size bytes */
#define QXB_ELEM_OTHER  6     /* present but not decoded (vlen strings, references, Julia structs) */
int  qxb_jld2_open(const char* path, qxb_jld2** f);
void qxb_jld2_close(qxb_jld2* f);
/* datasets reachable from the root group (JLD2's `_types` group skipped); checksum_failures (may be NULL) counts
 * version-2 structures whose lookup3 checksum did not match. */
int  qxb_jld2_count(const qxb_jld2* f, int* n, int* checksum_failures);
/* name is owned by the handle; dims (Julia / column-major order, room for QXB_JLD2_MAX_RANK) and the rest may be NULL */
int  qxb_jld2_info(const qxb_jld2* f, int i, const char** name, int* elem_kind, int* elem_size, int* rank, int64_t* dims);
/* as_c64 != 0: numeric elements converted to ComplexF64 (16 bytes each); 0: the stored bytes (elem_size each) */
int  qxb_jld2_read(const qxb_jld2* f, int i, void* out, int as_c64);
/* commit_types != 0 stores the complex datatypes under `_types/` and references them (JLD2.jl's layout);
 * 0 writes them inline (any HDF5 tool reads it; JLD2.jl sees NamedTuple{(:re, :im)} elements). */
int  qxb_jld2_write(const char* path, int n, const char* const* names, const int* elem_kinds, const int* elem_sizes,
                    const int* ranks, const int64_t* const* dims, const void* const* data, int commit_types);
/* qxb_graph_set_data for every numeric dataset of the file (dataset name = data label). */
int  qxb_graph_load_jld2(qxb_graph* g, const char* path, int* n_set /*may be NULL*/);

/* Parameter file (src/outputs.jl:47-78, docs/src/users_guide.md:22-35). */
#define QXB_METHOD_LIST      0
#define QXB_METHOD_UNIFORM   1
#define QXB_METHOD_REJECTION 2
typedef struct qxb_params {
    int32_t method;
    int32_t has_seed;
    int64_t seed;
    int64_t num_qubits;       /* List: length of the strings */
    int64_t num_samples;
    double  M;                /* Rejection (outputs.jl:57-62) */
    int32_t fix_M;
    int32_t reserved;
    int64_t n_bitstrings;     /* List: strings in the file; Uniform: num_samples strings drawn WITH replacement
                                 (simulation.jl:24-28) from the library's documented splitmix64 stream -- Julia's
                                 MersenneTwister stream is not reproduced; Rejection: 0 */
} qxb_params;
/* bitstrings (may be NULL to query): n_bitstrings x (num_qubits + 1) bytes, each string NUL-terminated. */
int  qxb_params_read(const char* yml_path, qxb_params* p, char* bitstrings, int64_t buflen);

/* QXContexts.execute(dsl_file, input_file, param_file, output_file; max_amplitudes, max_slices) (bin/qxrun.jl:83-87)
 * on this process' GPU.  input_file / param_file NULL = the DSL name with .jld2 / .yml (qxrun.jl:21,25);
 * max_amplitudes / max_slices < 0 = all, else the FIRST n (qxrun.jl:32-39); replan_candidates > 0 runs
 * qxb_graph_replan first.  Writes output_file (may be NULL) as JLD2 with datasets "bitstrings"
 * (fixed-length strings) and "amplitudes" (complex of `dtype`).  seconds (may be NULL) receives the reference's
 * timer sections: [0] parse input files, [1] create context, [2] simulation, [3] write results.
 * All three methods of outputs.jl:54-77: List, Uniform, and Rejection (empirical-supremum rejection sampling,
 * docs/src/features.md:68-84, candidates in batches of 1024; the file then also holds the scalars "M" and "drawn"). */
int  qxb_execute_files(const char* dsl_file, const char* input_file, const char* param_file, const char* output_file,
                       int dtype, int64_t max_amplitudes, int64_t max_slices, int replan_candidates,
                       int64_t* n_amplitudes, double* seconds);
/* The same on n_devices GPUs of this node (qxb_multi; List / Uniform methods -- Rejection stays on one device):
 * what `qxrun -m [-s K]` runs.  n_devices <= 0: all visible devices; sub_comm_size as in qxb_multi_amplitudes. */
int  qxb_execute_files_multi(const char* dsl_file, const char* input_file, const char* param_file, const char* output_file,
                             int dtype, int64_t max_amplitudes, int64_t max_slices, int replan_candidates,
                             int n_devices, int sub_comm_size, int64_t* n_amplitudes, double* seconds);

/* test hook (not part of the drop-in surface): shared-memory offset contributed by tile-index bit `bit` in the
 * tensor-core GEMM kernels' staging layouts; tests/test_mma_layout.py replays the kernels' index arithmetic on the CPU */
int  qxb_debug_mma_smem_bit(int dtype, int is_b, int tile_bits, int bit);
/* test hook: the same for the tcgen05 / TMEM kernel (csrc/qxb_gemm_tc5.cu): BYTE offset in the canonical K-major
 * no-swizzle core-matrix layout (8 rows x 16 bytes contiguous, 128 B between 16-byte K chunks, 1024 B between 8-row groups) */
int  qxb_debug_tc5_smem_bit(int tile_bits, int bit);
/* test hook: raw array of the per-contraction launch templates (struct OpParams of csrc/qxb_kernels.cuh, one per
 * lowered op, pointers unset) for n_free batched variables; returns the bytes needed.  Host logic only. */
int64_t qxb_debug_templates(qxb_graph* g, int n_free, void* buf, int64_t buflen);
/* test hook: the row program (csrc/qxb_rowprog.h) of one phase (1 = block, 2 = chunk) for the batched variables in
 * free_mask, serialised (layout in csrc/qxb_exec.cu) for tests/rowprog_emulator.py; returns the bytes needed.
 * Host logic only. */
int64_t qxb_debug_rowprog(qxb_graph* g, uint64_t free_mask, int phase, void* buf, int64_t buflen);
/* Test hook: the chunk-phase HBM arena plan of a variant with its fused chain, as text ("fused first last", "op ...",
 * "tensor index offset elems amp leaf" lines); returns the bytes needed (including the final NUL), negative = error. */
int64_t qxb_debug_fused_plan(qxb_graph* g, uint64_t free_mask, char* buf, int64_t buflen);
/* test hook: the lookup3 checksum of HDF5 version-2 metadata, checked against Jenkins' published vectors */
uint32_t qxb_debug_lookup3(const void* data, size_t n, uint32_t initval);

#ifdef __cplusplus
}
#endif
#endif /* QXB200_H */
