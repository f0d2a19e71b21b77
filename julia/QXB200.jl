# QXB200.jl -- Julia shim over libqxb200.so (C ABI: include/qxb200.h).
#
# Drop-in for the two executors behind QXTools' hot path:
#   * QXTns.contract_tn!      (call site QXTools/src/simulation.jl:89)   -> QXB200.single_amplitude / run_simulation
#   * QXContexts.execute      (call site QXTools/bin/qxrun.jl:83-87)     -> QXB200.execute
# Everything else (convert_to_tnc, flow_cutter_contraction_plan, contraction_scheme,
# build_compute_graph, generate_simulation_files) stays QXTools' own Julia code.
#
# NOTE: written against the ABI but NOT runnable in the build container (no Julia there): UNTESTED until it runs
# under a CI with Julia + the QX packages (INTEGRATION.md says the same).
module QXB200

using QXTools, QXContexts, JLD2, YAML
using AbstractTrees                       # PostOrderDFS, as compute_graph.jl:3 of the reference
using DataStructures: OrderedDict

const LIB = get(ENV, "QXB200_LIB", "libqxb200.so")
const C32, C64 = Cint(0), Cint(1)

struct QXBError <: Exception
    code::Cint
    msg::String
end
check(rc) = rc < 0 ? throw(QXBError(rc, unsafe_string(ccall((:qxb_last_error, LIB), Cstring, ())))) : rc

mutable struct Graph
    h::Ptr{Cvoid}
    dtype::Cint
    function Graph(dtype::Cint)
        r = Ref{Ptr{Cvoid}}(C_NULL)
        check(ccall((:qxb_graph_create, LIB), Cint, (Ref{Ptr{Cvoid}}, Cint), r, dtype))
        g = new(r[], dtype)
        finalizer(x -> ccall((:qxb_graph_destroy, LIB), Cvoid, (Ptr{Cvoid},), x.h), g)
        g
    end
end

# Mirror of `qxb_options` (include/qxb200.h), field for field; all zero = the library's defaults.  The constructors below
# pass C_NULL (defaults); build an `Options`, wrap it in a `Ref` and pass it to `qxb_graph_compile` to change a knob.
Base.@kwdef struct Options
    hbm_budget_bytes::Int64 = 0
    amp_batch::Int64 = 0
    profile::Int32 = 0
    no_cuda_graph::Int32 = 0
    sum_at_root::Int32 = 0
    no_smem_stage::Int32 = 0
    no_gemm::Int32 = 0
    gemm_mode::Int32 = 0            # 0 = auto, 1 = SIMT FMA GEMM only, 2 = tensor cores (tcgen05 / DMMA), 4 = mma.sync only
    row_programs::Int32 = 0         # 0 = auto, 1 = never, 2 = both phases, 3 = block phase only
    min_lob::Int32 = 0
    kc_regs_multi::Int32 = 0
    kc_regs_one::Int32 = 0
    smem_tma::Int32 = 0
    row_min_tt_bits::Int32 = 0
    row_tile_regs::Int32 = 0
    row_ctas_per_sm::Int32 = 0
    ring::Int32 = 0
    chain::Int32 = 0
    row_dmma::Int32 = 0
    row_chunk_max_amps::Int32 = 0
    streaming::Int32 = 0            # huge x tiny nodes on bigsmall_kernel: 0 = auto (on), 1 = off, 2 = TMA staging, 3 = FFMA2
    row_bank_opt::Int32 = 0         # 0 = auto (on), 1 = off
    chain_side::Int32 = 0           # depth of the side branches a fused chain takes in (0 = none)
end

i64(v) = Int64.(collect(v))

# one qxb_graph_* call per command, in post-order (children before parents) --
# the order QXContexts.generate_dsl_files writes them (users_guide.md:71-90)
function add!(g::Graph, c::QXContexts.LoadCommand)
    d = i64(c.dims)
    check(ccall((:qxb_graph_load, LIB), Cint, (Ptr{Cvoid}, Cstring, Cstring, Ptr{Int64}, Cint),
                g.h, string(c.name), string(c.label), d, length(d)))
end
add!(g::Graph, c::QXContexts.OutputCommand) =
    check(ccall((:qxb_graph_output, LIB), Cint, (Ptr{Cvoid}, Cstring, Int64, Int64), g.h, string(c.name), c.idx, c.dim))
add!(g::Graph, c::QXContexts.ViewCommand) =
    check(ccall((:qxb_graph_view, LIB), Cint, (Ptr{Cvoid}, Cstring, Cstring, Cstring, Int64, Int64),
                g.h, string(c.name), string(c.target), string(c.slice_sym), c.bond_index, c.bond_dim))
function add!(g::Graph, c::QXContexts.ContractCommand)
    o, l, r = i64(c.output_idxs), i64(c.left_idxs), i64(c.right_idxs)
    check(ccall((:qxb_graph_ncon, LIB), Cint,
                (Ptr{Cvoid}, Cstring, Ptr{Int64}, Cint, Cstring, Ptr{Int64}, Cint, Cstring, Ptr{Int64}, Cint),
                g.h, string(c.output_name), o, length(o), string(c.left_name), l, length(l),
                string(c.right_name), r, length(r)))
end
add!(g::Graph, c::QXContexts.SaveCommand) =
    check(ccall((:qxb_graph_save, LIB), Cint, (Ptr{Cvoid}, Cstring, Cstring), g.h, string(c.label), string(c.name)))

function set_data!(g::Graph, tensors::AbstractDict)
    for (label, a) in tensors
        data = convert(Array{ComplexF64}, a)            # column-major, as tensor_cache.jl:52-53
        d = i64(size(data))
        check(ccall((:qxb_graph_set_data, LIB), Cint, (Ptr{Cvoid}, Cstring, Ptr{ComplexF64}, Ptr{Int64}, Cint),
                    g.h, string(label), data, d, length(d)))
    end
end

function Graph(cg::QXContexts.ComputeGraph; dtype::Cint=C64)
    g = Graph(dtype)
    for node in AbstractTrees.PostOrderDFS(cg.root)
        add!(g, node.op)
    end
    set_data!(g, cg.tensors)
    check(ccall((:qxb_graph_compile, LIB), Cint, (Ptr{Cvoid}, Ptr{Cvoid}), g.h, C_NULL))
    g
end

function Graph(dsl_text::String, tensors::AbstractDict; dtype::Cint=C32, replan::Int=64, n_amp_model::Int=1024,
               autoslice_budget::Int=0)
    g = Graph(dtype)
    check(ccall((:qxb_graph_parse_dsl, LIB), Cint, (Ptr{Cvoid}, Cstring, Csize_t), g.h, dsl_text, sizeof(dsl_text)))
    set_data!(g, tensors)
    # batch-aware re-planning of the ncon tree (exact; leaves and views untouched).  With autoslice_budget > 0 the
    # library also ADDS slice variables (views, as compute_graph.jl:39-58 emits them) until the largest tensor of
    # the searched tree fits autoslice_budget / 3 bytes: GPU-aware slicing instead of contraction_scheme's `num`.
    if replan > 0
        n_free = autoslice_budget > 0 ? Cint(-3) : Cint(-1)
        check(ccall((:qxb_graph_replan_ex, LIB), Cint,
                    (Ptr{Cvoid}, Cint, Int64, Cint, Int64, UInt64, Ptr{Cint}, Ptr{Cdouble}, Ptr{Cdouble}, Ptr{Cdouble}),
                    g.h, replan, n_amp_model, n_free, autoslice_budget, UInt64(0), C_NULL, C_NULL, C_NULL, C_NULL))
    end
    check(ccall((:qxb_graph_compile, LIB), Cint, (Ptr{Cvoid}, Ptr{Cvoid}), g.h, C_NULL))
    g
end

num_slices(g::Graph) = (n = Ref{Int64}(0); check(ccall((:qxb_graph_num_slices, LIB), Cint, (Ptr{Cvoid}, Ref{Int64}), g.h, n)); n[])
num_outputs(g::Graph) = (n = Ref{Cint}(0); check(ccall((:qxb_graph_num_outputs, LIB), Cint, (Ptr{Cvoid}, Ref{Cint}), g.h, n)); Int(n[]))

const CH = Dict('0' => 0x00, '1' => 0x01, '+' => 0x02, '-' => 0x03)

"amplitudes of `bitstrings` summed over slices [slice_begin, slice_end) (0-based, half-open)"
function amplitudes(g::Graph, bitstrings::Vector{String}; slice_begin::Int=0, slice_end::Int=num_slices(g))
    nq, n = num_outputs(g), length(bitstrings)
    bits = Matrix{UInt8}(undef, nq, n)                 # [n_amp][n_outputs] row-major == (nq, n) column-major
    for (j, s) in enumerate(bitstrings), i in 1:nq
        bits[i, j] = CH[s[i]]                           # char i <-> qubit i (docs/src/basics.md:55)
    end
    T = g.dtype == C32 ? ComplexF32 : ComplexF64
    out = Vector{T}(undef, n)
    check(ccall((:qxb_amplitudes, LIB), Cint, (Ptr{Cvoid}, Ptr{UInt8}, Int64, Int64, Int64, Ptr{Cvoid}),
                g.h, bits, n, slice_begin, slice_end, out))
    out
end

# ---- seam B1: QXTools.single_amplitude / run_simulation (src/simulation.jl:86-133) ----------
function single_amplitude(tnc::TensorNetworkCircuit, plan::Array{NTuple{3, Symbol}, 1},
                          amplitude::Union{String, Nothing}=nothing)
    cg = build_compute_graph(tnc, plan)                 # the caller's tnc is copied inside (compute_graph.jl:18)
    g = Graph(cg; dtype=C64)
    amplitudes(g, [amplitude === nothing ? "0"^qubits(tnc) : amplitude])[1]
end

function run_simulation(circ; num_amplitudes=nothing, seed=nothing)
    tnc = convert_to_tnc(circ)
    plan = flow_cutter_contraction_plan(tnc; hypergraph=true)
    if num_amplitudes === nothing && qubits(tnc) > 30 num_amplitudes = 1000 end
    amps = num_amplitudes === nothing ? amplitudes_all(qubits(tnc)) : amplitudes_uniform(qubits(tnc), seed, num_amplitudes)
    bs = unique(collect(amps))
    g = Graph(build_compute_graph(tnc, plan); dtype=C64)
    OrderedDict{String, ComplexF64}(zip(bs, amplitudes(g, bs)))   # one batched call instead of one contraction per bitstring
end

# ---- seam B3: QXContexts.execute (bin/qxrun.jl:83-87) ----------------------------------------
function execute(dsl_file::String, input_file::Union{String, Nothing}=nothing,
                 param_file::Union{String, Nothing}=nothing, output_file::String="";
                 use_mpi::Bool=false, sub_comm_size::Int=1, use_gpu::Bool=true,
                 max_amplitudes::Union{Int, Nothing}=nothing, max_slices::Union{Int, Nothing}=nothing,
                 timings::Bool=false, elt::Type=ComplexF32)
    use_gpu || error("QXB200 has no CPU path")
    input_file === nothing && (input_file = splitext(dsl_file)[1] * ".jld2")
    param_file === nothing && (param_file = splitext(dsl_file)[1] * ".yml")
    tensors = load(input_file)                          # JLD2: label => N-d ComplexF64 array
    params = YAML.load_file(param_file)["output"]
    if params["method"] != "List"
        # Uniform / Rejection (src/outputs.jl:57-72): the library's own samplers, through the whole-triple entry point
        output_file == "" && (output_file = tempname() * ".jld2")
        execute_native(dsl_file, input_file, param_file, output_file; max_amplitudes=max_amplitudes, max_slices=max_slices, elt=elt)
        res = load(output_file)            # amplitudes: compound {re, im} -> NamedTuple elements (INTEGRATION.md)
        amps = [Complex(a.re, a.im) for a in res["amplitudes"]]
        return OrderedDict(zip(String.(res["bitstrings"]), amps))
    end
    bs = Vector{String}(params["params"]["bitstrings"])
    max_amplitudes === nothing || (bs = bs[1:min(end, max_amplitudes)])
    rank, nranks = 0, 1
    MPI = nothing
    if use_mpi
        # one process per GPU.  The device is chosen BEFORE the graph is built: compile uploads the leaves and folds
        # constants on the current device.  Local rank = rank modulo the GPUs of the node (QXB200_GPUS_PER_NODE, 8).
        MPI = Base.require(Base.PkgId(Base.UUID("da04e1cc-30fd-572f-bb4f-1f8673147195"), "MPI"))
        MPI.Initialized() || MPI.Init()
        rank, nranks = MPI.Comm_rank(MPI.COMM_WORLD), MPI.Comm_size(MPI.COMM_WORLD)
        nranks % sub_comm_size == 0 || error("sub_comm_size must divide the number of ranks")
        check(ccall((:qxb_init, LIB), Cint, (Cint,), rank % parse(Int, get(ENV, "QXB200_GPUS_PER_NODE", "8"))))
    else
        sub_comm_size = 1
    end
    g = Graph(read(dsl_file, String), tensors; dtype=(elt == ComplexF32 ? C32 : C64))
    S = max_slices === nothing ? num_slices(g) : min(max_slices, num_slices(g))
    # two levels, as docs/src/users_guide.md:11-20: the bitstrings are split across the nranks / sub_comm_size
    # sub-communicators, the slices across the ranks inside one (qxtools.jl_b200/dist.py is the tested mirror)
    n_groups = nranks ÷ sub_comm_size
    group, r = rank ÷ sub_comm_size, rank % sub_comm_size
    a0, a1 = (length(bs) * group) ÷ n_groups, (length(bs) * (group + 1)) ÷ n_groups
    amps = zeros(elt == ComplexF32 ? ComplexF32 : ComplexF64, length(bs))
    if a1 > a0
        amps[a0+1:a1] = amplitudes(g, bs[a0+1:a1]; slice_begin=(S * r) ÷ sub_comm_size, slice_end=(S * (r + 1)) ÷ sub_comm_size)
    end
    # groups own disjoint bitstrings (zeros elsewhere), ranks of a group hold partial sums over slices: one sum does both
    use_mpi && (amps = MPI.Allreduce(amps, +, MPI.COMM_WORLD))
    results = OrderedDict(zip(bs, amps))
    if rank == 0 && output_file != ""
        jldopen(output_file, "w") do io
            io["bitstrings"] = bs
            io["amplitudes"] = amps
        end
    end
    results
end

"""
    execute_native(dsl_file, input_file, param_file, output_file; max_amplitudes, max_slices, elt, replan)

The whole file triple inside the library (`qxb_execute_files`): `.qx` parser, native JLD2 and YAML readers, compile,
contraction, JLD2 results file.  Same defaults as `bin/qxrun.jl:21,25`.  Returns the number of amplitudes and the
four timer sections of the reference's table (parse, context, simulation, write), in seconds.
"""
function execute_native(dsl_file::String, input_file::Union{String, Nothing}=nothing,
                        param_file::Union{String, Nothing}=nothing, output_file::String="";
                        max_amplitudes::Union{Int, Nothing}=nothing, max_slices::Union{Int, Nothing}=nothing,
                        elt::Type=ComplexF32, replan::Int=0)
    n, sec = Ref{Int64}(0), zeros(Cdouble, 4)
    opt(x) = x === nothing ? C_NULL : x
    check(ccall((:qxb_execute_files, LIB), Cint,
                (Cstring, Cstring, Cstring, Cstring, Cint, Int64, Int64, Cint, Ref{Int64}, Ptr{Cdouble}),
                dsl_file, opt(input_file), opt(param_file), output_file, elt == ComplexF32 ? C32 : C64,
                max_amplitudes === nothing ? -1 : max_amplitudes, max_slices === nothing ? -1 : max_slices, replan, n, sec))
    n[], sec
end

end # module
