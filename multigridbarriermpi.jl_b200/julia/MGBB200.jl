# MGBB200.jl - the Julia `ccall` shim over libmgb_b200.so (include/mgb_b200.h).
#
# UNTESTED HERE: neither julia nor MPI exists in the build image (SURVEY.md 0.3).  This file is what a
# maintainer of MultiGridBarrierMPI.jl would add as a package extension so that the existing
# fem{1,2,3}d_mpi_solve / amgb entry points (src/MultiGridBarrierMPI.jl:594-600, 661-667, 735-745) run
# their per-Newton-step assembly through the B200 kernels.  It hooks the same seam the reference uses:
# methods of MultiGridBarrier generics specialised on HPCSparseArrays types (src:62-192).
#
# Integration point: upstream `barrier(F; F1, F2)` builds closures f0/f1/f2(z, x, w, c, R, D, z0).
# `b200_barrier(Q)` below returns a Barrier whose three closures call mgb_assemble on a per-(D,R)
# cached plan instead of hcat/map_rows/spdiagm/SpGEMM.
module MGBB200

using SparseArrays, LinearAlgebra
using CUDA                      # device arrays only; no CUDA.jl kernels on the hot path
using HPCSparseArrays: HPCVector, HPCMatrix, HPCSparseMatrix
import MPI                      # already a dependency of the reference (Project.toml: MPI = "0.20")
import MultiGridBarrier
import MultiGridBarrier: Barrier

const LIB = get(ENV, "MGB_B200_LIB", joinpath(@__DIR__, "..", "libmgb_b200.so"))

struct MgbCsr
    nrows::Int64; ncols::Int64; nnz::Int64
    rowptr::Ptr{Int32}; colidx::Ptr{Int32}; vals::Ptr{Float64}
    index_base::Int32
end
struct MgbBarrier          # field-for-field the C struct mgb_barrier (include/mgb_b200.h)
    kind::Int32; nidx::Int32; idx::NTuple{8,Int32}; p::Float64; slack::Int32
    nidx2::Int32; idx2::NTuple{8,Int32}; p2::Float64          # optional second cone (parabolic_solve); nidx2 = 0: none
end
MgbBarrier(idx::Vector{Int}, p::Float64, slack::Bool; idx2::Vector{Int} = Int[], p2::Float64 = 2.0) =
    MgbBarrier(1, length(idx), ntuple(i -> i <= length(idx) ? Int32(idx[i] - 1) : Int32(0), 8), p, slack,
               length(idx2), ntuple(i -> i <= length(idx2) ? Int32(idx2[i] - 1) : Int32(0), 8), p2)

lasterr() = unsafe_string(ccall((:mgb_last_error, LIB), Cstring, ()))
check(rc) = rc == 0 ? nothing : error("libmgb_b200: " * lasterr())

mutable struct Ctx; h::Ptr{Cvoid}; end
function Ctx(dev::Integer = CUDA.deviceid())
    r = Ref{Ptr{Cvoid}}(C_NULL)
    # the task-local CUDA.jl stream keeps ordering with the caller's other device work
    check(ccall((:mgb_ctx_create, LIB), Cint, (Cint, Ptr{Cvoid}, Ref{Ptr{Cvoid}}), dev, CUDA.stream().handle, r))
    c = Ctx(r[]); finalizer(c -> ccall((:mgb_ctx_destroy, LIB), Cint, (Ptr{Cvoid},), c.h), c); c
end

# An HPCSparseMatrix holds ONLY this rank's rows, with COMPRESSED column ids (rowval indexes col_indices,
# src/MultiGridBarrierMPI.jl:216-221) - its arrays are not a global CSR.  Two ways into the library:
#   * hpc_block(A, rank) below: the rank's block exactly as stored, zero-copy (mgb_plan_create_local), for the operators D;
#   * GatheredCsr(A): the WHOLE matrix on every rank, through the collective gather the reference itself uses in
#     mpi_to_native (`SparseMatrixCSC(A)`, src:357-371), converted to CSR (= CSC of the transpose).  The symbolic phase
#     needs R whole, once per level (a few MB: nnz(R) = 4.6e5 at fem2d L=8); no numeric call gathers anything.
struct GatheredCsr
    nrows::Int; ncols::Int
    rowptr::Vector{Int32}; colidx::Vector{Int32}; vals::Vector{Float64}
end
function GatheredCsr(A::HPCSparseMatrix)
    At = SparseMatrixCSC(transpose(SparseMatrixCSC(A)))      # collective; CSC of A' = CSR of A
    GatheredCsr(size(A, 1), size(A, 2), Int32.(At.colptr), Int32.(At.rowval), Vector{Float64}(At.nzval))
end
csr(G::GatheredCsr) = MgbCsr(G.nrows, G.ncols, length(G.vals), pointer(G.rowptr), pointer(G.colidx), pointer(G.vals), Int32(1))

mutable struct Plan; h::Ptr{Cvoid}; m::Int; nnzH::Int; rowptr::Vector{Int32}; colidx::Vector{Int32}; end

function Plan(ctx::Ctx, D::Vector{<:HPCSparseMatrix}, R::HPCSparseMatrix, x::Matrix{Float64}, w::Vector{Float64};
              idx::Vector{Int}, p::Float64, slack::Bool = false, rows = (0, size(D[1], 1)))
    Dg = [GatheredCsr(d) for d in D]; Rg = GatheredCsr(R)     # collective gathers (setup, once per level)
    Ds = [csr(d) for d in Dg]; Rs = Ref(csr(Rg))
    bar = Ref(MgbBarrier(idx, p, slack))
    r = Ref{Ptr{Cvoid}}(C_NULL)
    GC.@preserve Dg Rg x w begin
        check(ccall((:mgb_plan_create, LIB), Cint,
                    (Ptr{Cvoid}, Int64, Int32, Ptr{MgbCsr}, Ref{MgbCsr}, Int32, Ptr{Float64}, Ptr{Float64},
                     Ref{MgbBarrier}, Int64, Int64, Int32, Ref{Ptr{Cvoid}}),
                    ctx.h, size(D[1], 1), length(D), Ds, Rs, size(x, 2), x, w, bar, rows[1], rows[2], 0, r))
    end
    info = zeros(Int64, 15)
    check(ccall((:mgb_plan_info, LIB), Cint, (Ptr{Cvoid}, Ptr{Int64}, Int32), r[], info, 15))
    m, nnzH = info[4], info[5]
    rp = zeros(Int32, m + 1); ci = zeros(Int32, nnzH)
    check(ccall((:mgb_plan_pattern, LIB), Cint, (Ptr{Cvoid}, Ptr{Int32}, Ptr{Int32}), r[], rp, ci))
    pl = Plan(r[], m, nnzH, rp, ci)
    finalizer(p -> ccall((:mgb_plan_destroy, LIB), Cint, (Ptr{Cvoid},), p.h), pl); pl
end

# local part of dot / sum / norm^2 / max|.| on device vectors (deterministic); the caller all-reduces the scalar
function reduce_local(ctx::Ctx, op::Integer, x::CuVector{Float64}, y::Union{CuVector{Float64},Nothing} = nothing)
    out = Ref{Float64}(0.0)
    GC.@preserve x y begin
        check(ccall((:mgb_reduce, LIB), Cint, (Ptr{Cvoid}, Int32, CuPtr{Float64}, CuPtr{Float64}, Int64, CuPtr{Float64}, Ref{Float64}),
                    ctx.h, op, pointer(x), y === nothing ? CU_NULL : pointer(y), length(x), CU_NULL, out))
    end
    out[]
end
LinearAlgebra.dot(ctx::Ctx, x::HPCVector, y::HPCVector, comm) = MPI.Allreduce(reduce_local(ctx, 0, x.v, y.v), +, comm)

# The rank's own storage of a row-partitioned HPCSparseMatrix, field for field (constructor order
# src/MultiGridBarrierMPI.jl:216-221): colptr indexes LOCAL rows, rowval holds COMPRESSED column ids,
# col_indices maps them to global columns.  Passed zero-copy to mgb_plan_create_local (struct mgb_hpc_block).
struct MgbHpcBlock
    nrows_local::Int64; ncols_compressed::Int64; ncols_global::Int64; row0::Int64
    colptr::Ptr{Int32}; rowval::Ptr{Int32}; nzval::Ptr{Float64}; col_indices::Ptr{Int32}
    index_base::Int32
end
hpc_block(A::HPCSparseMatrix, rank::Integer) =
    MgbHpcBlock(A.nrows_local, A.ncols_compressed, size(A, 2), A.row_partition[rank + 1] - 1,
                pointer(A.colptr), pointer(A.rowval), pointer(A.nzval), pointer(A.col_indices), Int32(1))

# Plan from the local blocks: no rank ever needs the global operators D (only R is gathered, once).
function LocalPlan(ctx::Ctx, comm, D::Vector{<:HPCSparseMatrix}, R::HPCSparseMatrix, x::HPCMatrix, w::HPCVector;
                   idx::Vector{Int}, p::Float64, slack::Bool = false)
    rank = MPI.Comm_rank(comm)
    Rg = GatheredCsr(R)
    Ds = [hpc_block(d, rank) for d in D]; Rs = Ref(csr(Rg))
    bar = Ref(MgbBarrier(idx, p, slack))
    r = Ref{Ptr{Cvoid}}(C_NULL)
    xl = Array(x.A); wl = Array(w.v)          # local rows (HPCMatrix.A / HPCVector.v, src:176)
    GC.@preserve D Rg xl wl begin
        check(ccall((:mgb_plan_create_local, LIB), Cint,
                    (Ptr{Cvoid}, Int64, Int32, Ptr{MgbHpcBlock}, Ref{MgbCsr}, Int32, Ptr{Float64}, Ptr{Float64},
                     Ref{MgbBarrier}, Int32, Ref{Ptr{Cvoid}}),
                    ctx.h, size(D[1], 1), length(D), Ds, Rs, size(xl, 2), xl, wl, bar, 0, r))
    end
    info = zeros(Int64, 16)
    check(ccall((:mgb_plan_info, LIB), Cint, (Ptr{Cvoid}, Ptr{Int64}, Int32), r[], info, 16))
    m, nnzH = info[4], info[5]
    rp = zeros(Int32, m + 1); ci = zeros(Int32, nnzH)
    check(ccall((:mgb_plan_pattern, LIB), Cint, (Ptr{Cvoid}, Ptr{Int32}, Ptr{Int32}), r[], rp, ci))
    pl = Plan(r[], m, nnzH, rp, ci)
    finalizer(p -> ccall((:mgb_plan_destroy, LIB), Cint, (Ptr{Cvoid},), p.h), pl); pl
end

const WANT_F0, WANT_GRAD, WANT_HESS, STORE_DZ = 1, 2, 4, 8

"numeric phase: everything stays on the device (CuArray pointers are borrowed for the call)"
function assemble!(pl::Plan, s::CuVector{Float64}, Dz0::CuMatrix{Float64}, c::CuMatrix{Float64}, t::Float64, flags;
                   scal::CuVector{Float64}, grad = nothing, hval = nothing, Dz = nothing)
    ptr(a) = a === nothing ? CU_NULL : pointer(a)
    GC.@preserve s Dz0 c scal grad hval Dz begin
        check(ccall((:mgb_assemble, LIB), Cint,
                    (Ptr{Cvoid}, CuPtr{Float64}, CuPtr{Float64}, CuPtr{Float64}, Float64, Int32,
                     CuPtr{Float64}, CuPtr{Float64}, CuPtr{Float64}, CuPtr{Float64}),
                    pl.h, pointer(s), pointer(Dz0), pointer(c), t, flags, pointer(scal), ptr(grad), ptr(hval), ptr(Dz)))
    end
end

# ---- the ten overloads of src/MultiGridBarrierMPI.jl:62-192 that change meaning on this backend ----
# amgb_all_isfinite (src:121-133): device reduction + MPI/NCCL AND stays as in the reference; the local
# part can call mgb_all_isfinite.
function all_isfinite_local(ctx::Ctx, v::CuVector{Float64})
    flag = Ref{Int32}(0)
    check(ccall((:mgb_all_isfinite, LIB), Cint, (Ptr{Cvoid}, CuPtr{Float64}, Int64, Ref{Int32}), ctx.h, pointer(v), length(v), flag))
    flag[] == 1
end
# map_rows_gpu (src:168-170) for the barrier closures -> mgb_map_barrier; amgb_diag (src:137-147) is never
# called by b200_barrier (w .* y is applied inside the kernels); amgb_zeros / amgb_blockdiag / vertex_indices /
# _raw_array / _rows_to_svectors / _to_cpu_array keep the reference definitions.

"""
    b200_barrier(; idx, p) -> MultiGridBarrier.Barrier

Drop-in for `MultiGridBarrier.barrier(F; F1, F2)` on HPC types: f0/f1/f2 share one plan per (D, R)
pair (symbolic phase once per level) and return an HPCVector / HPCSparseMatrix on the plan's frozen
pattern, so `MultiGridBarrier.solve(H, g)` (MUMPS, test/test_newton_matrix_compare.jl:51) is unchanged.
"""
function b200_barrier(; idx::Vector{Int}, p::Float64, ctx::Ctx = Ctx())
    plans = IdDict{Any,Any}()
    # per (R, D): the plan plus its device buffers.  Dz0 = D*z0 is refreshed when z0 changes (once per Newton
    # solve: the operator-only plan built with R = I and MGB_PLAN_NO_HESSIAN gives D*z0 through mgb_apply_D).
    function state(x, w, R, D)
        get!(plans, (R, D)) do
            pl = Plan(ctx, D, R, Matrix(x), Vector(w); idx = idx, p = p)
            n, nD = size(D[1], 1), length(D)
            (pl = pl, s = CUDA.zeros(Float64, pl.m), Dz0 = CUDA.zeros(Float64, n, nD), c = CUDA.zeros(Float64, n, nD),
             scal = CUDA.zeros(Float64, 4), grad = CUDA.zeros(Float64, pl.m), hval = CUDA.zeros(Float64, pl.nnzH),
             z0id = Ref{UInt}(0), cid = Ref{UInt}(0))
        end
    end
    function run(flags, s, x, w, c, R, D, z0)
        st = state(x, w, R, D)
        if objectid(z0) != st.z0id[]                       # new Newton solve: Dz0 = hcat([D_k*z0]...) (test/test_apply_d.jl:44)
            copyto!(st.Dz0, reduce(hcat, [Vector(Dk * z0) for Dk in D])); st.z0id[] = objectid(z0)
        end
        if objectid(c) != st.cid[]
            copyto!(st.c, Matrix(c)); st.cid[] = objectid(c)
        end
        copyto!(st.s, MultiGridBarrier._raw_array(s))      # src/MultiGridBarrierMPI.jl:175-176
        assemble!(st.pl, st.s, st.Dz0, st.c, 1.0, flags; scal = st.scal, grad = st.grad, hval = st.hval)
        st
    end
    backend(s) = s.backend
    f0(s, x, w, c, R, D, z0) = (st = run(WANT_F0, s, x, w, c, R, D, z0); sc = Array(st.scal); sc[2] == 1.0 ? sc[1] : Inf)
    f1(s, x, w, c, R, D, z0) = (st = run(WANT_GRAD, s, x, w, c, R, D, z0); HPCVector(Array(st.grad), backend(s)))
    function f2(s, x, w, c, R, D, z0)
        st = run(WANT_HESS, s, x, w, c, R, D, z0)
        pl = st.pl                                          # frozen CSR pattern (0-based -> 1-based), values of this step
        At = SparseMatrixCSC(pl.m, pl.m, Int32.(pl.rowptr .+ 1), Int32.(pl.colidx .+ 1), Array(st.hval))   # CSC of H' = CSR of H
        HPCSparseMatrix(SparseMatrixCSC(transpose(At)), backend(s))
    end
    Barrier(f0 = f0, f1 = f1, f2 = f2)
end

# ---- multi-GPU: one MPI rank per GPU (test/test_2d.jl:17-20), owner-computes sharding (mgb_dist_*) ----
struct MgbIpcHandle; bytes::NTuple{64,UInt8}; end

mutable struct DistPlan
    h::Ptr{Cvoid}; rank::Int; nranks::Int; info::Vector{Int64}
    rowptr::Vector{Int32}; colidx::Vector{Int32}     # owned rows of the pattern of R'HR (relative rowptr, global columns)
    rows::Vector{Int64}                              # global ids (0-based) of the quadrature rows this rank evaluates
end

"""
    colocated_partition(sizes, nranks) -> (perm, out_part)

Rank-major renumbering of unknowns stacked by variable (`sizes[k]` = columns of block k of R = blockdiag(R_u, R_s, ..)):
`perm[new] = old` (1-based), and in the new numbering rank r owns the contiguous block `out_part[r+1]+1 : out_part[r+2]`
(0-based offsets as the C ABI takes them) = its uniform block of u, then of s, ...  Pass `R[:, perm]` and `s[perm]` to the
plan; the gradient comes back as `g[perm]`, the Hessian rows / columns in the new numbering.  With the reference's
stacked numbering a rank owning a block of `[u | s]` evaluates every element touching its u rows AND every element
touching its s rows (all of them at 2 ranks); renumbered, it evaluates E/P elements plus a halo (same rule as
`mgb_b200.dist.colocated_partition`, which the Python host side applies by default).
"""
function colocated_partition(sizes::Vector{Int}, nranks::Int)
    block(n, r) = (q = divrem(n, nranks); (r * q[1] + min(r, q[2]), (r + 1) * q[1] + min(r + 1, q[2])))   # 0-based [lo, hi)
    offs = cumsum([0; sizes])
    perm = Int[]; out_part = Int64[0]
    for r in 0:nranks-1
        for (k, n) in enumerate(sizes)
            lo, hi = block(n, r)
            append!(perm, (offs[k] + lo + 1):(offs[k] + hi))
        end
        push!(out_part, length(perm))
    end
    perm, out_part
end

"""
    DistPlan(ctx, comm, D, R, x, w; idx, p, row_part, out_part)

Collective over `comm` (MPI.Comm).  The symbolic phase needs the operators whole, once per level: they are gathered
here (`GatheredCsr`; in the reference every rank builds the full native mesh anyway, src:239-240) together with the
0-based partition offsets of the quadrature rows (whole elements) and of the unknowns (HPCSparseArrays row partition).
The plan evaluates `pl.rows` - every element touching this rank's output rows - and completes its rows of R'HR and its
block of the gradient locally.  The 64-byte CUDA IPC handles of the scalar windows travel through one MPI.Allgather;
afterwards no MPI/NCCL call is made per step.
"""
function DistPlan(ctx::Ctx, comm, D, R, x::Matrix{Float64}, w::Vector{Float64}; idx::Vector{Int}, p::Float64,
                  row_part::Vector{Int64}, out_part::Vector{Int64}, slack::Bool = false)
    rank, nranks = MPI.Comm_rank(comm), MPI.Comm_size(comm)
    Dg = [GatheredCsr(d) for d in D]; Rg = GatheredCsr(R)
    Ds = [csr(d) for d in Dg]; Rs = Ref(csr(Rg)); bar = Ref(MgbBarrier(idx, p, slack)); r = Ref{Ptr{Cvoid}}(C_NULL)
    GC.@preserve Dg Rg x w row_part out_part begin
        check(ccall((:mgb_dist_plan_create, LIB), Cint,
                    (Ptr{Cvoid}, Int64, Int32, Ptr{MgbCsr}, Ref{MgbCsr}, Int32, Ptr{Float64}, Ptr{Float64}, Ref{MgbBarrier},
                     Int32, Int32, Ptr{Int64}, Ptr{Int64}, Ref{Ptr{Cvoid}}),
                    ctx.h, size(D[1], 1), length(D), Ds, Rs, size(x, 2), x, w, bar, rank, nranks, row_part, out_part, r))
    end
    mine = Ref(MgbIpcHandle(ntuple(_ -> 0x00, 64)))
    check(ccall((:mgb_dist_export, LIB), Cint, (Ptr{Cvoid}, Ref{MgbIpcHandle}), r[], mine))
    all = MPI.Allgather(mine[], comm)                       # Vector{MgbIpcHandle}, one per rank
    check(ccall((:mgb_dist_attach, LIB), Cint, (Ptr{Cvoid}, Ptr{MgbIpcHandle}), r[], all))
    MPI.Barrier(comm)                                       # every window mapped before the first store
    info = zeros(Int64, 16)
    check(ccall((:mgb_dist_info, LIB), Cint, (Ptr{Cvoid}, Ptr{Int64}, Int32), r[], info, 16))
    rp = zeros(Int32, info[6] - info[5] + 1); ci = zeros(Int32, info[3]); rows = zeros(Int64, info[7])
    check(ccall((:mgb_dist_pattern, LIB), Cint, (Ptr{Cvoid}, Ptr{Int32}, Ptr{Int32}), r[], rp, ci))
    check(ccall((:mgb_dist_rows, LIB), Cint, (Ptr{Cvoid}, Ptr{Int64}), r[], rows))
    DistPlan(r[], rank, nranks, info, rp, ci, rows)         # destroy collectively: MPI.Barrier, then mgb_plan_destroy
end

"""collective, same order on every rank.  `s_own`: this rank's block of the Newton unknown (the local storage of the
HPCVector, `MultiGridBarrier._raw_array(s)`, src:175-176) - the library all-gathers it over NVLink peer memory;
`Dz0`, `c`: rows `pl.rows .+ 1` of the n x nD blocks.  Returns device arrays (owned H values in `pl.colidx` order, owned
gradient block, the 4 global scalars)."""
function dist_assemble_s!(pl::DistPlan, s_own::CuVector{Float64}, Dz0::CuMatrix{Float64}, c::CuMatrix{Float64}, t::Float64, flags)
    hp, gp, sp = Ref{CuPtr{Float64}}(), Ref{CuPtr{Float64}}(), Ref{CuPtr{Float64}}()
    GC.@preserve s_own Dz0 c begin
        check(ccall((:mgb_dist_assemble_s, LIB), Cint,
                    (Ptr{Cvoid}, CuPtr{Float64}, CuPtr{Float64}, CuPtr{Float64}, Float64, Int32,
                     Ref{CuPtr{Float64}}, Ref{CuPtr{Float64}}, Ref{CuPtr{Float64}}),
                    pl.h, pointer(s_own), pointer(Dz0), pointer(c), t, flags, hp, gp, sp))
    end
    (unsafe_wrap(CuArray, hp[], pl.info[3]), unsafe_wrap(CuArray, gp[], pl.info[4]), unsafe_wrap(CuArray, sp[], 4))
end

"the same with a replicated unknown `s` (length m) already on every rank"
function dist_assemble!(pl::DistPlan, s::CuVector{Float64}, Dz0::CuMatrix{Float64}, c::CuMatrix{Float64}, t::Float64, flags)
    hp, gp, sp = Ref{CuPtr{Float64}}(), Ref{CuPtr{Float64}}(), Ref{CuPtr{Float64}}()
    GC.@preserve s Dz0 c begin
        check(ccall((:mgb_dist_assemble, LIB), Cint,
                    (Ptr{Cvoid}, CuPtr{Float64}, CuPtr{Float64}, CuPtr{Float64}, Float64, Int32,
                     Ref{CuPtr{Float64}}, Ref{CuPtr{Float64}}, Ref{CuPtr{Float64}}),
                    pl.h, pointer(s), pointer(Dz0), pointer(c), t, flags, hp, gp, sp))
    end
    (unsafe_wrap(CuArray, hp[], pl.info[3]), unsafe_wrap(CuArray, gp[], pl.info[4]), unsafe_wrap(CuArray, sp[], 4))
end

"page-lock a Julia Array once so mgb_assemble_host copies into it at DMA rate"
pin!(a::Array) = (check(ccall((:mgb_host_register, LIB), Cint, (Ptr{Cvoid}, Int64), a, sizeof(a))); a)
unpin!(a::Array) = check(ccall((:mgb_host_unregister, LIB), Cint, (Ptr{Cvoid},), a))

end # module
