# B200Backend.jl -- the reference-side binding of libpq_b200.so.
#
# A PicoQuant maintainer drops this file next to src/backends/interactive.jl and adds
# `include("backends/b200.jl")` to src/backends.jl.  It defines a backend type that plugs
# into the existing interface (src/backends.jl:3,62-66), so TensorNetworkCircuit
# (src/layer3.jl:124-134), contract_pair!/contract_network!/full_wavefunction_contraction!
# (src/layer2.jl) and the slicing code (src/layer2/slicing.jl) run unchanged on top of it:
#
#     tn = convert_qiskit_circ_to_network(circ, B200Backend{ComplexF64}())
#
# UNTESTED IN THE BUILD ENVIRONMENT: the image this backend was developed in has no Julia
# toolchain; the same C ABI is exercised end-to-end through the ctypes twin
# (picoquant.jl_b200/host/b200_backend.py), whose tests mirror PicoQuant's own.

export B200Backend

const libpq_b200 = get(ENV, "PQ_B200_LIB", "libpq_b200.so")

const PQ_C64, PQ_C128 = Cint(0), Cint(1)
const PQ_HOST_F32, PQ_HOST_F64, PQ_HOST_C64, PQ_HOST_C128 = Cint(0), Cint(1), Cint(2), Cint(3)
const PQ_ERR_NOT_FOUND = Cint(-2)
const PQ_ERR_SHAPE = Cint(-3)
const PQ_MAX_RANK = 64

mutable struct B200Backend{T<:Union{ComplexF32,ComplexF64}} <: AbstractBackend
    handle::Ptr{Cvoid}
    metrics::Metrics

    function B200Backend{T}(device::Integer=0) where {T}
        h = Ref{Ptr{Cvoid}}(C_NULL)
        rc = ccall((:pq_create, libpq_b200), Cint, (Cint, Cint, Ref{Ptr{Cvoid}}),
                   device, T === ComplexF64 ? PQ_C128 : PQ_C64, h)
        rc == 0 || error("pq_create failed ($rc): a CUDA device is required (no CPU fallback)")
        backend = new{T}(h[], Metrics())
        finalizer(b -> ccall((:pq_destroy, libpq_b200), Cint, (Ptr{Cvoid},), b.handle), backend)
        backend
    end
end
B200Backend() = B200Backend{ComplexF32}()   # same default element type as InteractiveBackend()

function pq_check(backend::B200Backend, rc::Cint)
    rc == 0 && return
    msg = unsafe_string(ccall((:pq_last_error, libpq_b200), Cstring, (Ptr{Cvoid},), backend.handle))
    rc == PQ_ERR_NOT_FOUND && throw(KeyError(msg))
    rc == PQ_ERR_SHAPE && throw(DimensionMismatch(msg))
    error("libpq_b200 error $rc: $msg")
end

host_code(::Type{Float32}) = PQ_HOST_F32
host_code(::Type{Float64}) = PQ_HOST_F64
host_code(::Type{ComplexF32}) = PQ_HOST_C64
host_code(::Type{ComplexF64}) = PQ_HOST_C128

# src/backends/interactive.jl:32-36
function save_tensor_data(backend::B200Backend, tensor_label::Symbol,
                          tensor_data::AbstractArray{S}) where {S<:Number}
    data = S <: Union{Float32,Float64,ComplexF32,ComplexF64} ? Array(tensor_data) :
           Array{ComplexF64}(tensor_data)
    dims = Int64[size(data)...]
    pq_check(backend, ccall((:pq_save_tensor, libpq_b200), Cint,
                            (Ptr{Cvoid}, Cstring, Cint, Ptr{Int64}, Ptr{Cvoid}, Cint),
                            backend.handle, String(tensor_label), length(dims), dims, data,
                            host_code(eltype(data))))
end

# Batched form of save_tensor_data (one staging copy, one host->device transfer, one launch)
# for the O(#gates) tiny node tensors of a network -- e.g. the per-rank upload of
# examples/dist_slicing_example.jl:22-27:
#     save_tensors(backend, [(:node_1, data1), (:node_2, data2), ...])
function save_tensors(backend::B200Backend, items::Vector{<:Tuple{Symbol,AbstractArray}})
    arrays = [eltype(d) <: Union{Float32,Float64,ComplexF32,ComplexF64} ? Array(d) :
              Array{ComplexF64}(d) for (_, d) in items]
    labels = [String(l) for (l, _) in items]
    ranks = Cint[ndims(a) for a in arrays]
    dims = Int64[s for a in arrays for s in size(a)]
    codes = Cint[host_code(eltype(a)) for a in arrays]
    GC.@preserve arrays labels begin
        ptrs = Ptr{Cvoid}[Ptr{Cvoid}(pointer(a)) for a in arrays]
        cstrs = Cstring[Base.unsafe_convert(Cstring, l) for l in labels]
        pq_check(backend, ccall((:pq_save_tensors, libpq_b200), Cint,
                                (Ptr{Cvoid}, Cint, Ptr{Cstring}, Ptr{Cint}, Ptr{Int64},
                                 Ptr{Ptr{Cvoid}}, Ptr{Cint}),
                                backend.handle, length(arrays), cstrs, ranks, dims, ptrs, codes))
    end
end

# src/backends/interactive.jl:44-49 (returns `nothing` when absent)
function load_tensor_data(backend::B200Backend{T}, tensor_label::Symbol) where {T}
    rank = Ref{Cint}(0)
    dims = zeros(Int64, PQ_MAX_RANK)
    rc = ccall((:pq_tensor_info, libpq_b200), Cint, (Ptr{Cvoid}, Cstring, Ref{Cint}, Ptr{Int64}),
               backend.handle, String(tensor_label), rank, dims)
    rc == PQ_ERR_NOT_FOUND && return nothing
    pq_check(backend, rc)
    out = Array{T}(undef, dims[1:rank[]]...)
    pq_check(backend, ccall((:pq_load_tensor, libpq_b200), Cint,
                            (Ptr{Cvoid}, Cstring, Ptr{Cvoid}, Cint),
                            backend.handle, String(tensor_label), out, host_code(T)))
    out
end

# src/backends/interactive.jl:60-75 (stores C, deletes A and B)
function contract_tensors(backend::B200Backend,
                          A_label::Symbol, A_ncon_indices::Array{Int, 1},
                          B_label::Symbol, B_ncon_indices::Array{Int, 1},
                          C_label::Symbol)
    a = Int32.(A_ncon_indices); b = Int32.(B_ncon_indices)
    pq_check(backend, ccall((:pq_contract, libpq_b200), Cint,
                            (Ptr{Cvoid}, Cstring, Ptr{Int32}, Cint, Cstring, Ptr{Int32}, Cint, Cstring),
                            backend.handle, String(A_label), a, length(a),
                            String(B_label), b, length(b), String(C_label)))
end

# src/backends/interactive.jl:84-88
function save_output(backend::B200Backend, node::Symbol, name::String="result")
    pq_check(backend, ccall((:pq_save_output, libpq_b200), Cint, (Ptr{Cvoid}, Cstring, Cstring),
                            backend.handle, String(node), name))
end

# src/backends/interactive.jl:97-102
function reshape_tensor(backend::B200Backend, tensor::Symbol, groups::Array{Array{Int, 1}, 1})
    flat = Int32.(vcat(groups...)); sizes = Int32.(length.(groups))
    pq_check(backend, ccall((:pq_reshape, libpq_b200), Cint,
                            (Ptr{Cvoid}, Cstring, Ptr{Int32}, Ptr{Int32}, Cint),
                            backend.handle, String(tensor), flat, sizes, length(sizes)))
end

# src/backends/interactive.jl:111-115
function permute_tensor(backend::B200Backend, tensor::Symbol, axes::Array{Int, 1})
    ax = Int32.(axes)
    pq_check(backend, ccall((:pq_permute, libpq_b200), Cint, (Ptr{Cvoid}, Cstring, Ptr{Int32}, Cint),
                            backend.handle, String(tensor), ax, length(ax)))
end

# src/backends/interactive.jl:159-161
function delete_tensor!(backend::B200Backend, tensor_label::Symbol)
    ccall((:pq_delete, libpq_b200), Cint, (Ptr{Cvoid}, Cstring), backend.handle, String(tensor_label))
end

# src/backends/interactive.jl:169-172
function view_tensor!(backend::B200Backend, view_node, node, bond_idx, bond_range)
    idx = Int32.(collect(bond_range))
    pq_check(backend, ccall((:pq_view, libpq_b200), Cint,
                            (Ptr{Cvoid}, Cstring, Cstring, Cint, Ptr{Int32}, Cint),
                            backend.handle, String(view_node), String(node), bond_idx, idx, length(idx)))
end

# src/backends/interactive.jl:130-152 -> src/layer1.jl:146-184 (SVD split on the device)
function decompose_tensor!(backend::B200Backend,
                           tensor::Symbol,
                           left_positions::Array{Int, 1},
                           right_positions::Array{Int, 1};
                           threshold::AbstractFloat=1e-13,
                           max_rank::Int=0,
                           left_label::Symbol,
                           right_label::Symbol)
    lp = Int32.(left_positions); rp = Int32.(right_positions)
    chi = Ref{Cint}(0)
    pq_check(backend, ccall((:pq_decompose, libpq_b200), Cint,
                            (Ptr{Cvoid}, Cstring, Ptr{Int32}, Cint, Ptr{Int32}, Cint, Cdouble, Cint,
                             Cstring, Cstring, Ref{Cint}),
                            backend.handle, String(tensor), lp, length(lp), rp, length(rp),
                            Float64(threshold), max_rank, String(left_label), String(right_label), chi))
    Int(chi[])
end

# ---- beyond the nine: the sliced loop of examples/dist_slicing_example.jl -----------------

"dst += src on the device (local accumulation of slice partials)"
accumulate!(backend::B200Backend, dst::Symbol, src::Symbol) =
    pq_check(backend, ccall((:pq_accumulate, libpq_b200), Cint, (Ptr{Cvoid}, Cstring, Cstring),
                            backend.handle, String(dst), String(src)))

"128-byte NCCL unique id; broadcast it (e.g. MPI.Bcast!) and call `comm_init!` on every rank"
function comm_unique_id()
    id = zeros(UInt8, 128)
    rc = ccall((:pq_comm_unique_id, libpq_b200), Cint, (Ptr{UInt8},), id)
    rc == 0 || error("pq_comm_unique_id failed ($rc)")
    id
end
comm_init!(backend::B200Backend, id::Vector{UInt8}, rank::Integer, nranks::Integer) =
    pq_check(backend, ccall((:pq_comm_init, libpq_b200), Cint, (Ptr{Cvoid}, Ptr{UInt8}, Cint, Cint),
                            backend.handle, id, rank, nranks))
"replaces MPI.Reduce!(wf_part, MPI.SUM, 0, comm) of dist_slicing_example.jl:30 (all ranks get the sum)"
allreduce_sum!(backend::B200Backend, label::Symbol) =
    pq_check(backend, ccall((:pq_allreduce_sum, libpq_b200), Cint, (Ptr{Cvoid}, Cstring),
                            backend.handle, String(label)))

"execute_dsl_file on the device: compile a .tl stream once, replay it per slice"
function compile_program(backend::B200Backend, tl_text::String)
    p = Ref{Ptr{Cvoid}}(C_NULL)
    pq_check(backend, ccall((:pq_program_compile, libpq_b200), Cint,
                            (Ptr{Cvoid}, Cstring, Ref{Ptr{Cvoid}}), backend.handle, tl_text, p))
    p[]
end
function run_program(backend::B200Backend, program::Ptr{Cvoid}, view_starts::Vector{Int}=Int[];
                     accumulate_into::Union{Symbol,Nothing}=nothing)
    vs = Int32.(view_starts)
    acc = accumulate_into === nothing ? C_NULL : String(accumulate_into)
    pq_check(backend, ccall((:pq_program_run, libpq_b200), Cint,
                            (Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Int32}, Cint, Cstring),
                            backend.handle, program, isempty(vs) ? C_NULL : vs, length(vs), acc))
end

"the slice loop in one call: one run per column of `view_starts` (nviews x nslices), partial
sums added in slice order, up to `lanes` slices in flight on the device"
function run_program_slices(backend::B200Backend, program::Ptr{Cvoid}, view_starts::Matrix{Int};
                            accumulate_into::Union{Symbol,Nothing}=nothing, lanes::Integer=2)
    vs = Int32.(vec(view_starts))
    acc = accumulate_into === nothing ? C_NULL : String(accumulate_into)
    pq_check(backend, ccall((:pq_program_run_slices, libpq_b200), Cint,
                            (Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Int32}, Cint, Cint, Cstring, Cint),
                            backend.handle, program, vs, size(view_starts, 2), size(view_starts, 1),
                            acc, lanes))
end
