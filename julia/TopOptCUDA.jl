# TopOptCUDA.jl -- Julia glue binding libtopopt_cuda through `ccall` as a new solver type for
# TopOpt.jl's FEASolver(...) factory.
#
# STATUS: NOT EXECUTED in this repository (Julia is not installed in the build image or on the GPU box).
# It is the reference-side binding a TopOpt.jl maintainer would add; the binding that IS tested is its
# 1:1 Python image topopt.jl_b200/_lib.py (ctypes) + topopt.jl_b200/{fea,functions,cheqfilters}.py.
# Every method below overrides or extends a TopOpt.jl method only for the new solver types
# (`S<:CUDASolvers`) -- nothing changes for the stock CPU solvers.
#
# Reference hooks (JuliaTopOpt/TopOpt.jl v0.14.0):
#   AbstractLinearSolver / solve_system!       src/FEA/solvers_api.jl:44,223-282
#   generic call operator (runs CPU assemble!) src/FEA/solvers_api.jl:285-374  -> specialised below (i)
#   FEASolver factory                          src/FEA/solvers_api.jl:468-574,599-603 -> lightweight method (vi)
#   ElementFEAInfo / GlobalFEAInfo             src/TopOptProblems/elementinfo.jl:24-110,121-156
#   ComplianceFun call / rrule                 src/Functions/compliance.jl:58-76
#   solve_adjoint!                             src/Functions/thermal_compliance.jl:169-210
#   DensityFilterFun / SensFilterFun           src/CheqFilters/density_filter.jl:14-47, sens_filter.jl:13-70
module TopOptCUDA

using TopOpt, TopOpt.FEA, TopOpt.TopOptProblems, TopOpt.Functions, TopOpt.CheqFilters
using TopOpt.Utilities: PowerPenaltyFun, RationalPenaltyFun, SinhPenaltyFun, ProjectedPenaltyFun,
    HeavisideProjectionFun, SigmoidProjectionFun, getpenalty
using TopOpt.TopOptProblems: ElementFEAInfo, GlobalFEAInfo, AbstractTopOptProblem, StiffnessTopOptProblem,
    HeatTransferTopOptProblem, floattype, getdim, getE, getν, getk, make_cload, default_quad_order
using TopOpt.CheqFilters: FilterMetadata
using Ferrite: ndofs, getncells, Lagrange, QuadratureRule, FacetQuadratureRule, CellValues, FacetValues, getrefshape
using LinearAlgebra, SparseArrays, StaticArrays, ChainRulesCore
import TopOpt.FEA: solve_system!, GenericFEASolver, AbstractLinearSolver, AbstractPhysics, FEASolver, CGStateVariables

const lib = get(ENV, "LIBTOPOPT_CUDA", "libtopopt_cuda")

struct CUDAMatrixFreeSolver <: AbstractLinearSolver end
struct CUDAAssemblySolver <: AbstractLinearSolver end
const CUDASolvers = Union{CUDAMatrixFreeSolver,CUDAAssemblySolver}
opcode(::Type{CUDAMatrixFreeSolver}) = Cint(0)
opcode(::Type{CUDAAssemblySolver}) = Cint(1)
Base.show(io::IO, ::MIME"text/plain", ::CUDAMatrixFreeSolver) = print(io, "TopOpt CUDA matrix-free CG solver (libtopopt_cuda, sm_100a)")
Base.show(io::IO, ::MIME"text/plain", ::CUDAAssemblySolver) = print(io, "TopOpt CUDA assembled-CSR CG solver (libtopopt_cuda, sm_100a)")

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
struct CGOpts                      # topopt_cg_opts
    abstol::Float64; reltol::Float64
    maxiter::Int32; op::Int32; precond::Int32; criteria::Int32; check_every::Int32
    variant::Int32                 # 0 = IterativeSolvers' recurrence (default), 1 = single-pass recurrence
    warm_start::Int32              # 0 = zero initial guess like the reference
    refresh_precond::Int32         # 0 = preconditioner built once like the reference
    mg_degree::Int32               # multigrid preconditioner (precond = 2): Chebyshev degree, 0 = default
    reserved0::Int32
    mg_ratio::Float64              # ... smoothed part of the spectrum, 0 = default
end
mutable struct CGResult
    iters::Int32; converged::Int32; residual::Float64; tol::Float64; solve_ms::Float64
    CGResult() = new(0, 0, 0.0, 0.0, 0.0)
end

# Opt-in extensions over the reference, process-wide (all off by default so that parity holds)
const OPTIONS = Dict{Symbol,Int32}(:variant => 0, :warm_start => 0, :refresh_precond => 0, :multigrid => 0)

"Device handle of one solver; destroyed by the GC finalizer."
mutable struct Handle
    ptr::Ptr{Cvoid}
    solved_once::Bool
    function Handle(p)
        h = new(p, false)
        finalizer(x -> (x.ptr == C_NULL || ccall((:topopt_destroy, lib), Cint, (Ptr{Cvoid},), x.ptr); x.ptr = C_NULL), h)
        return h
    end
end
Base.show(io::IO, h::Handle) = print(io, "TopOptCUDA.Handle(", h.ptr, ")")
const HANDLES = WeakKeyDict{Any,Handle}()   # solver => device handle (GenericFEASolver has fixed fields)

"Device filter; keeps its solver handle alive (the library's filter points into it) and frees itself first."
mutable struct FilterHandle
    ptr::Ptr{Cvoid}
    owner::Handle
    function FilterHandle(p, owner)
        f = new(p, owner)
        finalizer(f) do x
            if x.ptr != C_NULL && x.owner.ptr != C_NULL
                ccall((:topopt_filter_destroy, lib), Cint, (Ptr{Cvoid},), x.ptr)
            end
            x.ptr = C_NULL
        end
        return f
    end
end
Base.show(io::IO, f::FilterHandle) = print(io, "TopOptCUDA.FilterHandle(", f.ptr, ")")

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

# Kes[1] is an ElementMatrix in the stock ElementFEAInfo and a plain SMatrix in the lightweight one (vi)
kematrix(K::TopOptProblems.ElementMatrix) = Matrix{Float64}(TopOptProblems.rawmatrix(K))
kematrix(K::AbstractMatrix) = Matrix{Float64}(K)

function handle(s::GenericFEASolver{T,P,S}) where {T,P,S<:CUDASolvers}
    get!(HANDLES, s) do
        problem = s.problem
        dim = getdim(problem)
        ncomp = size(problem.metadata.node_dofs, 1)
        nels = problem.rect_grid.nels
        sizes = problem.rect_grid.sizes
        Ke = kematrix(s.elementinfo.Kes[1])
        any(!=(0), problem.ch.inhomogeneities) && throw(ArgumentError(
            "CUDA solvers do not support inhomogeneous Dirichlet BCs (same rule as CGMatrixFreeSolver)"))
        pres = Vector{Int64}(problem.ch.prescribed_dofs)
        fl = Vector{Float64}(s.elementinfo.fixedload)
        cv = s.elementinfo.cellvolumes isa RepeatedVector ? Float64[] : Vector{Float64}(s.elementinfo.cellvolumes)
        # the cross-check of the numbering costs 24 x nel Int64 on the host: skipped for the lightweight path
        cd = s.elementinfo.Kes isa RepeatedVector ? Matrix{Int64}(undef, 0, 0) : Matrix{Int64}(problem.metadata.cell_dofs)
        out = Ref{Ptr{Cvoid}}(C_NULL)
        GC.@preserve Ke pres fl cv cd begin
            d = Desc(dim, ncomp, (nels..., ntuple(_ -> 1, 3 - dim)...), (Float64.(sizes)..., ntuple(_ -> 1.0, 3 - dim)...),
                     pointer(Ke), pointer(pres), length(pres), pointer(fl),
                     isempty(cv) ? Ptr{Float64}(C_NULL) : pointer(cv), isempty(cd) ? Ptr{Int64}(C_NULL) : pointer(cd),
                     0.0, 0, 0, 1, C_NULL)
            check(ccall((:topopt_create, lib), Cint, (Ref{Desc}, Ref{Ptr{Cvoid}}), d, out))
        end
        Handle(out[])
    end
end

criteria_code(::FEA.DefaultCriteria) = Cint(0)
criteria_code(::FEA.EnergyCriteria) = Cint(1)
cgopts(s::GenericFEASolver{T,P,S}; warm=false) where {T,P,S} = CGOpts(s.abstol, sqrt(eps(T)), s.cg_max_iter, opcode(S),
    OPTIONS[:multigrid] != 0 ? 2 : (s.preconditioner === identity ? 0 : 1), criteria_code(s.conv), 0,
    OPTIONS[:variant], (warm && OPTIONS[:warm_start] != 0) ? 1 : 0, OPTIONS[:refresh_precond], 0, 0, 0.0)

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
    default_rhs = f === s.globalinfo.f
    rhs = default_rhs ? C_NULL : pointer(f)   # NULL = fixedload with prescribed entries zeroed
    GC.@preserve f lhs check(ccall((:topopt_solve, lib), Cint,
        (Ptr{Cvoid}, Ptr{Float64}, Ptr{Float64}, Ref{CGOpts}, Ref{CGResult}), h.ptr, rhs, lhs,
        cgopts(s; warm=default_rhs && h.solved_once), res), h.ptr)
    default_rhs && (h.solved_once = true)
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

# (iv) adjoint solves (thermal compliance, DisplacementFun, TemperatureFun, ...)
function Functions.solve_adjoint!(s::GenericFEASolver{T,P,S}, lhs, rhs) where {T,P,S<:CUDASolvers}
    solve_system!(S, s, nothing, rhs, lhs)
    return nothing
end

# (v) filters -------------------------------------------------------------------------------------
function make_filter(s::GenericFEASolver, rmin)
    out = Ref{Ptr{Cvoid}}(C_NULL)
    h = handle(s)
    check(ccall((:topopt_filter_create, lib), Cint, (Ptr{Cvoid}, Float64, Ref{Ptr{Cvoid}}), h.ptr, Float64(rmin), out), h.ptr)
    return FilterHandle(out[], h)
end
filter_apply!(y, f::FilterHandle, x, mode) = (GC.@preserve x y check(ccall((:topopt_filter_apply, lib), Cint,
    (Ptr{Cvoid}, Ptr{Float64}, Ptr{Float64}, Cint), f.ptr, x, y, mode), f.owner.ptr); y)

# density filter: a lazy operator in place of the explicit sparse Jacobian; the stock call + rrule
# (density_filter.jl:35-47: `jacobian * x`, `jacobian' * Δ`) run unchanged on it
struct CUDAFilterOperator{T} <: AbstractMatrix{T}
    f::FilterHandle; n::Int; transpose::Bool
end
Base.size(J::CUDAFilterOperator) = (J.n, J.n)
Base.adjoint(J::CUDAFilterOperator{T}) where {T} = CUDAFilterOperator{T}(J.f, J.n, !J.transpose)
Base.getindex(::CUDAFilterOperator, ::Int, ::Int) = error("CUDAFilterOperator is matrix-free: use mul! / *")
Base.show(io::IO, ::MIME"text/plain", J::CUDAFilterOperator) = print(io, J.n, "×", J.n, " matrix-free TopOpt filter operator on ", J.f)
LinearAlgebra.mul!(y::AbstractVector, J::CUDAFilterOperator, x::AbstractVector) = filter_apply!(y, J.f, x, J.transpose ? 1 : 0)
Base.:*(J::CUDAFilterOperator{T}, x::AbstractVector) where {T} = mul!(similar(x, T), J, x)
function CheqFilters.DensityFilterFun(s::GenericFEASolver{T,P,S}, rmin::Real, ::Type{TI}=Int) where {T,P,S<:CUDASolvers,TI<:Integer}
    J = CUDAFilterOperator{T}(make_filter(s, rmin), length(s.vars), false)
    return DensityFilterFun(FilterMetadata(T, TI), T(rmin), J)
end

# sensitivity filter: identity forward (stock method, sens_filter.jl:49-51); the pullback is the nodal
# smoothing map, i.e. the library's forward stencil.  The device filter rides in the FilterMetadata type
# parameter so that the rrule below is more specific than the stock one (sens_filter.jl:52-70).
struct CUDAFilterRef
    f::FilterHandle
end
const CUDASensFilter{T,TV,TE} = SensFilterFun{T,TV,TE,<:FilterMetadata{CUDAFilterRef}}
function CheqFilters.SensFilterFun(s::GenericFEASolver{T,P,S}, rmin::Real, ::Type{TI}=Int) where {T,P,S<:CUDASolvers,TI<:Integer}
    meta = FilterMetadata(CUDAFilterRef(make_filter(s, rmin)), nothing)
    return SensFilterFun(s.elementinfo, meta, T(rmin), zeros(T, 0), zeros(T, length(s.vars)), zeros(T, 0))
end
function ChainRulesCore.rrule(cf::CUDASensFilter, x::TopOpt.PseudoDensities)
    return cf(x), Δ -> begin
        Δ = ChainRulesCore.unthunk(Δ)
        d = Vector{Float64}(hasproperty(Δ, :x) ? Δ.x : Δ)
        newΔ = filter_apply!(similar(d), cf.metadata.cell_neighbouring_nodes.f, d, 0)
        copyto!(cf.last_grad, newΔ)
        (NoTangent(), Tangent{typeof(x)}(; x=newΔ))
    end
end

# u'Ku on the device with the current stiffness (src/FEA/FEA.jl:40)
FEA.getcompliance(s::GenericFEASolver{T,P,S}) where {T,P,S<:CUDASolvers} = begin
    out = Ref{Float64}(0.0)
    check(ccall((:topopt_quadratic_form, lib), Cint, (Ptr{Cvoid}, Ptr{Float64}, Ref{Float64}), handle(s).ptr, C_NULL, out), handle(s).ptr)
    out[]
end

# (vi) lightweight construction ----------------------------------------------------------------------
# The stock factory (solvers_api.jl:468-574) builds one ElementMatrix per element (19 GB at 256x128x128)
# and the 1 G-nnz sparsity pattern (GlobalFEAInfo -> allocate_matrix).  On a uniform grid the library needs
# ONE element matrix and no pattern, so the CUDA solver types get their own method: a repeating Kes / fes /
# cellvolumes view and a 1x1 dummy K.  Everything that reads these fields elementwise keeps working.
struct RepeatedVector{T} <: AbstractVector{T}
    v::T
    n::Int
end
Base.size(r::RepeatedVector) = (r.n,)
Base.getindex(r::RepeatedVector, i::Int) = (@boundscheck checkbounds(r, i); r.v)
Base.IndexStyle(::Type{<:RepeatedVector}) = IndexLinear()
Base.sum(r::RepeatedVector) = r.n * r.v

physics_code(::StiffnessTopOptProblem) = Cint(0)
physics_code(::HeatTransferTopOptProblem) = Cint(1)
material(p::StiffnessTopOptProblem) = (Float64(getE(p)), Float64(getν(p)))
material(p::HeatTransferTopOptProblem) = (Float64(getk(p)), 0.0)

function lightweight_elementinfo(problem::AbstractTopOptProblem, quad_order)
    T = floattype(problem)
    dim = getdim(problem)
    ncomp = size(problem.metadata.node_dofs, 1)
    nel = getncells(problem.ch.dh.grid)
    ks = ncomp * 2^dim
    sizes = Float64[problem.rect_grid.sizes...]
    a, b = material(problem)
    Ke = Matrix{Float64}(undef, ks, ks)
    check(ccall((:topopt_element_matrix, lib), Cint, (Cint, Cint, Ptr{Float64}, Float64, Float64, Cint, Ptr{Float64}),
                dim, physics_code(problem), sizes, a, b, quad_order, Ke))
    interpolation = TopOptProblems._base_interpolation(problem.ch.dh.subdofhandlers[1].field_interpolations[1])
    refshape = getrefshape(interpolation)
    cellvalues = CellValues(QuadratureRule{refshape}(quad_order), interpolation, interpolation)
    facevalues = FacetValues(FacetQuadratureRule{refshape}(quad_order), interpolation, interpolation)
    Kes = RepeatedVector(SMatrix{ks,ks,T}(Ke), nel)
    fes = RepeatedVector(zero(SVector{ks,T}), nel)                   # no body force (zero in every BASELINE config)
    fixedload = Vector{T}(make_cload(problem))                       # point loads; distributed loads -> use the stock factory
    cellvolumes = RepeatedVector(T(prod(sizes)), nel)                # elementinfo.jl:180-192 on a uniform grid
    return ElementFEAInfo{dim,T}(Kes, fes, fixedload, cellvolumes, cellvalues, facevalues, problem.metadata, problem.ch.dh.grid.cells)
end

function FEASolver(::Type{Physics}, ::Type{S}, problem::AbstractTopOptProblem;
        quad_order=default_quad_order(problem), xmin=nothing, penalty=nothing, prev_penalty=nothing, qr=false,
        cg_max_iter=700, abstol=nothing, preconditioner=identity, conv=FEA.DefaultCriteria(), kwargs...) where {Physics<:AbstractPhysics,S<:CUDASolvers}
    T = floattype(problem)
    hasproperty(problem, :rect_grid) || throw(ArgumentError("CUDA solvers need a structured RectilinearGrid problem"))
    _xmin = xmin === nothing ? T(1) / 1000 : T(xmin)
    _penalty = penalty === nothing ? PowerPenaltyFun{T}(1) : penalty
    _prev_penalty = prev_penalty === nothing ? deepcopy(_penalty) : prev_penalty
    _abstol = abstol === nothing ? T(1e-7) : T(abstol)
    elementinfo = lightweight_elementinfo(problem, quad_order)
    n = ndofs(problem.ch.dh)
    globalinfo = GlobalFEAInfo(spzeros(T, 1, 1), zeros(T, n))         # K is never formed on the host
    u = zeros(T, n)
    cg_statevars = CGStateVariables{T,typeof(u)}(T[], T[], T[])       # the CG state lives on the device
    vars = fill(one(T), getncells(problem.ch.dh.grid))
    meandiag = T(sum(diag(elementinfo.Kes[1])) * length(vars))        # solvers_api.jl:526-527
    return GenericFEASolver{T,Physics,S,typeof(_penalty),typeof(problem),typeof(globalinfo),typeof(elementinfo),typeof(u),
                            typeof(cg_max_iter),typeof(cg_statevars),typeof(preconditioner),typeof(conv)}(
        problem, globalinfo, elementinfo, u, similar(u), similar(u), vars, _penalty, _prev_penalty, _xmin, qr, cg_max_iter, _abstol,
        cg_statevars, preconditioner, Ref(false), conv, meandiag, Int[], Int[], Vector{Vector{T}}[])
end

function Base.show(io::IO, ::MIME"text/plain", s::GenericFEASolver{T,P,S}) where {T,P,S<:CUDASolvers}
    print(io, "TopOpt FEA solver on libtopopt_cuda (", S, "): ", length(s.u), " dofs, ", length(s.vars), " elements, abstol=", s.abstol,
          ", cg_max_iter=", s.cg_max_iter)
end

end # module
