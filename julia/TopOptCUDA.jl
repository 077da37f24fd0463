# TopOptCUDA.jl -- Julia glue binding libtopopt_cuda through `ccall` as a new solver type for
# TopOpt.jl's FEASolver(...) factory.  NOT EXECUTED in this repository's CI: Julia is not installed
# in the build image.  It mirrors topopt.jl_b200/_lib.py (the ctypes binding that IS tested) 1:1.
#
# Reference hooks (JuliaTopOpt/TopOpt.jl v0.14.0):
#   AbstractLinearSolver / solve_system!      src/FEA/solvers_api.jl:44,223-282
#   generic call operator (runs CPU assemble!) src/FEA/solvers_api.jl:285-374  -> specialised below
#   FEASolver factory                          src/FEA/solvers_api.jl:468-574,599-603
#   ComplianceFun call / rrule                 src/Functions/compliance.jl:58-76
#   solve_adjoint!                             src/Functions/thermal_compliance.jl:169-210
#   DensityFilterFun / SensFilterFun           src/CheqFilters/density_filter.jl:14-47, sens_filter.jl:52-70
module TopOptCUDA

using TopOpt, TopOpt.FEA, TopOpt.TopOptProblems, TopOpt.Functions, TopOpt.CheqFilters
using TopOpt.Utilities: PowerPenaltyFun, RationalPenaltyFun, SinhPenaltyFun, ProjectedPenaltyFun,
    HeavisideProjectionFun, SigmoidProjectionFun, getpenalty
using LinearAlgebra, ChainRulesCore
import TopOpt.FEA: solve_system!, GenericFEASolver, AbstractLinearSolver

const lib = get(ENV, "LIBTOPOPT_CUDA", "libtopopt_cuda")

struct CUDAMatrixFreeSolver <: AbstractLinearSolver end
struct CUDAAssemblySolver <: AbstractLinearSolver end
const CUDASolvers = Union{CUDAMatrixFreeSolver,CUDAAssemblySolver}
opcode(::Type{CUDAMatrixFreeSolver}) = Cint(0)
opcode(::Type{CUDAAssemblySolver}) = Cint(1)

# ---- C structs (include/topopt_cuda.h) ---------------------------------------------------------
struct Desc
    dim::Int32; ncomp::Int32
    nels::NTuple{3,Int64}; sizes::NTuple{3,Float64}
    Ke::Ptr{Float64}; prescribed::Ptr{Int64}; n_prescribed::Int64
    fixedload::Ptr{Float64}; cellvolumes::Ptr{Float64}; cell_dofs::Ptr{Int64}
    fixed_diag::Float64
    device::Int32; rank::Int32; world::Int32
    nccl_unique_id::Ptr{Cvoid}
end
struct CGOpts
    abstol::Float64; reltol::Float64
    maxiter::Int32; op::Int32; precond::Int32; criteria::Int32; check_every::Int32; reserved::Int32
end
mutable struct CGResult
    iters::Int32; converged::Int32; residual::Float64; tol::Float64; solve_ms::Float64
    CGResult() = new(0, 0, 0.0, 0.0, 0.0)
end

mutable struct Handle
    ptr::Ptr{Cvoid}
    function Handle(p)
        h = new(p)
        finalizer(x -> ccall((:topopt_destroy, lib), Cint, (Ptr{Cvoid},), x.ptr), h)
        return h
    end
end
const HANDLES = WeakKeyDict{Any,Handle}()   # solver => device handle (GenericFEASolver has fixed fields)

function check(rc::Cint, h=C_NULL)
    rc == 0 && return
    msg = unsafe_string(ccall((:topopt_last_error, lib), Cstring, (Ptr{Cvoid},), h))
    rc == -1 && throw(ArgumentError(msg))        # solvers_api.jl:515-525 convention
    rc == -4 && throw(DomainError(NaN, msg))     # convergence_criteria.jl:34-41 convention
    error("libtopopt_cuda error $rc: $msg")
end

penalty_kind(::PowerPenaltyFun) = Cint(0)
penalty_kind(::RationalPenaltyFun) = Cint(1)
penalty_kind(::SinhPenaltyFun) = Cint(2)

function handle(s::GenericFEASolver{T,P,S}) where {T,P,S<:CUDASolvers}
    get!(HANDLES, s) do
        problem = s.problem
        dh = problem.ch.dh
        dim = TopOptProblems.getdim(problem)
        ncomp = size(problem.metadata.node_dofs, 1)
        nels = problem.rect_grid.nels
        sizes = problem.rect_grid.sizes
        Ke = Matrix{Float64}(TopOptProblems.rawmatrix(s.elementinfo.Kes[1]))
        any(!=(0), problem.ch.inhomogeneities) && throw(ArgumentError(
            "CUDA solvers do not support inhomogeneous Dirichlet BCs (same rule as CGMatrixFreeSolver)"))
        pres = Vector{Int64}(problem.ch.prescribed_dofs)
        fl = Vector{Float64}(s.elementinfo.fixedload)
        cv = Vector{Float64}(s.elementinfo.cellvolumes)
        cd = Matrix{Int64}(problem.metadata.cell_dofs)     # cross-checked against the internal numbering
        out = Ref{Ptr{Cvoid}}(C_NULL)
        GC.@preserve Ke pres fl cv cd begin
            d = Desc(dim, ncomp, (nels..., ntuple(_ -> 1, 3 - dim)...), (Float64.(sizes)..., ntuple(_ -> 1.0, 3 - dim)...),
                     pointer(Ke), pointer(pres), length(pres), pointer(fl), pointer(cv), pointer(cd),
                     0.0, 0, 0, 1, C_NULL)
            check(ccall((:topopt_create, lib), Cint, (Ref{Desc}, Ref{Ptr{Cvoid}}), d, out))
        end
        Handle(out[])
    end
end

criteria_code(::FEA.DefaultCriteria) = Cint(0)
criteria_code(::FEA.EnergyCriteria) = Cint(1)
cgopts(s::GenericFEASolver{T,P,S}) where {T,P,S} = CGOpts(s.abstol, sqrt(eps(T)), s.cg_max_iter, opcode(S),
    s.preconditioner === identity ? 0 : 1, criteria_code(s.conv), 0, 0)

projection(p) = (Cint(0), 0.0)
projection(p::ProjectedPenaltyFun{<:Any,<:Any,<:HeavisideProjectionFun}) = (Cint(1), Float64(p.proj.β))
projection(p::ProjectedPenaltyFun{<:Any,<:Any,<:SigmoidProjectionFun}) = (Cint(2), Float64(p.proj.β))
penalty_kind(p::ProjectedPenaltyFun) = penalty_kind(p.penalty)

function upload_density!(s, h)
    pen = getpenalty(s)
    pk, β = projection(pen)
    check(ccall((:topopt_set_projection, lib), Cint, (Ptr{Cvoid}, Cint, Float64), h.ptr, pk, β), h.ptr)
    check(ccall((:topopt_set_density, lib), Cint, (Ptr{Cvoid}, Ptr{Float64}, Cint, Float64, Float64, Cint),
                h.ptr, s.vars, penalty_kind(pen), pen.p, s.xmin, TopOpt.PENALTY_BEFORE_INTERPOLATION ? 1 : 0), h.ptr)
end

# (ii) the documented plug-in method
function solve_system!(::Type{S}, s::GenericFEASolver{T,P,S}, K, f, lhs; kwargs...) where {T,P,S<:CUDASolvers}
    h = handle(s)
    res = CGResult()
    rhs = f === s.globalinfo.f ? C_NULL : pointer(f)   # NULL = fixedload with prescribed entries zeroed
    GC.@preserve f lhs check(ccall((:topopt_solve, lib), Cint,
        (Ptr{Cvoid}, Ptr{Float64}, Ptr{Float64}, Ref{CGOpts}, Ref{CGResult}), h.ptr, rhs, lhs, cgopts(s), res), h.ptr)
    return false
end

# (i) more specific call operator: skips the CPU assemble! the generic one always runs
function (s::GenericFEASolver{T,P,S})(reuse_fact::Bool=false, ::Type{Val{safe}}=Val{false};
        assemble_f=true, rhs=assemble_f ? s.globalinfo.f : s.rhs, lhs=assemble_f ? s.u : s.lhs, kwargs...) where {T,P,S<:CUDASolvers,safe}
    h = handle(s)
    upload_density!(s, h)
    if s.preconditioner !== identity && !s.preconditioner_initialized[]
        check(ccall((:topopt_set_jacobi, lib), Cint, (Ptr{Cvoid}, Ptr{Float64}), h.ptr, C_NULL), h.ptr)
        s.preconditioner_initialized[] = true
    end
    if ndims(rhs) == 2 && size(rhs, 2) > 1          # solvers_api.jl:294-350
        for j in axes(rhs, 2)
            col = Vector{T}(rhs[:, j]); sol = zeros(T, length(col))
            solve_system!(S, s, nothing, col, sol)
            lhs[:, j] .= sol
        end
        return nothing
    end
    solve_system!(S, s, nothing, rhs, lhs)           # the library applies apply_zero! to a caller rhs
    return nothing
end

# (iii) fused compliance + sensitivity filling the fields the unchanged rrule reads
function (o::ComplianceFun{T,<:GenericFEASolver{T,P,S}})(x::TopOpt.PseudoDensities) where {T,P,S<:CUDASolvers}
    s = o.solver
    s.vars .= x.x
    s()
    obj = Ref{Float64}(0.0)
    check(ccall((:topopt_compliance, lib), Cint, (Ptr{Cvoid}, Ptr{Float64}, Ref{Float64}, Ptr{Float64}, Ptr{Float64}),
                handle(s).ptr, C_NULL, obj, o.cell_comp, o.grad), handle(s).ptr)
    return obj[]          # cell_comp and grad are filled in place, as compute_compliance does (compliance.jl:89-93)
end

# (iv) adjoint solves (thermal compliance, DisplacementFun, ...)
function Functions.solve_adjoint!(s::GenericFEASolver{T,P,S}, lhs, rhs) where {T,P,S<:CUDASolvers}
    solve_system!(S, s, nothing, rhs, lhs)
    return nothing
end

# (v) filters: a lazy operator in place of the explicit sparse Jacobian
struct CUDAFilterOperator{T} <: AbstractMatrix{T}
    ptr::Ptr{Cvoid}; n::Int; transpose::Bool
end
Base.size(J::CUDAFilterOperator) = (J.n, J.n)
Base.adjoint(J::CUDAFilterOperator{T}) where {T} = CUDAFilterOperator{T}(J.ptr, J.n, !J.transpose)
function LinearAlgebra.mul!(y::AbstractVector, J::CUDAFilterOperator, x::AbstractVector)
    check(ccall((:topopt_filter_apply, lib), Cint, (Ptr{Cvoid}, Ptr{Float64}, Ptr{Float64}, Cint), J.ptr, x, y, J.transpose ? 1 : 0))
    return y
end
Base.:*(J::CUDAFilterOperator{T}, x::AbstractVector) where {T} = mul!(similar(x, T), J, x)
function CheqFilters.DensityFilterFun(s::GenericFEASolver{T,P,S}, rmin::Real, ::Type{TI}=Int) where {T,P,S<:CUDASolvers,TI<:Integer}
    out = Ref{Ptr{Cvoid}}(C_NULL)
    check(ccall((:topopt_filter_create, lib), Cint, (Ptr{Cvoid}, Float64, Ref{Ptr{Cvoid}}), handle(s).ptr, rmin, out), handle(s).ptr)
    J = CUDAFilterOperator{T}(out[], length(s.vars), false)
    return DensityFilterFun(CheqFilters.FilterMetadata(T, TI), T(rmin), J)   # call + rrule run unchanged
end

FEA.getcompliance(s::GenericFEASolver{T,P,S}) where {T,P,S<:CUDASolvers} = begin
    out = Ref{Float64}(0.0)
    check(ccall((:topopt_dot, lib), Cint, (Ptr{Cvoid}, Ptr{Float64}, Ptr{Float64}, Ref{Float64}), handle(s).ptr, C_NULL, C_NULL, out))
    out[]
end

end # module
