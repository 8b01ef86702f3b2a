# PiccoloB200.jl -- thin `ccall` shim that plugs libpiccolo_b200.so (include/piccolo_b200.h) into
# Piccolo.jl's existing integrator extension point.
#
# NOT EXECUTED IN THIS REPOSITORY'S CI: the build image has no Julia toolchain and DirectTrajOpt.jl
# is not vendored in the reference, so the exact abstract-method signatures below are the ones
# visible from the reference's call sites:
#   evaluate!(δ, B, traj)                    src/control/integrators.jl:311, display/inspect.jl:623
#   eval_jacobian(B, traj) / *_structure     integrators.jl:780-782, test/aqua.jl:6-9
#   B.dim, B.x_dim, B.x_name                 integrators.jl:307-309,552
#   integrator= kwarg                        templates/smooth_pulse_problem.jl:123,213-233
#   register_integrator!                     src/specs/registries.jl:112,191
# Keep it mechanical: every numeric call is one ccall; no arithmetic happens in Julia.
module PiccoloB200

using SparseArrays
using Piccolo
using DirectTrajOpt
import DirectTrajOpt: AbstractIntegrator

const LIB = get(ENV, "PICCOLO_B200_LIB", "libpiccolo_b200")

const PB2_KET, PB2_UNITARY, PB2_DENSITY = Cint(0), Cint(1), Cint(2)
const PB2_HOST = Cint(0)

# mirrors `pb2_desc` in include/piccolo_b200.h (field order and widths must match)
struct PB2Desc
    kind::Int32; b::Int32; n_b::Int32; m::Int32; K::Int32; D::Int32
    x_off::Int32; dt_off::Int32; u_off::Int32; global_dim::Int32
    knot0::Int64; device::Int32; algorithm::Int32
    G0::Ptr{Float64}; Gj::Ptr{Float64}
end

check(rc) = rc == 0 || error("libpiccolo_b200: " * unsafe_string(ccall((:pb2_last_error, LIB), Cstring, ())))

mutable struct B200BilinearIntegrator <: AbstractIntegrator
    handle::Ptr{Cvoid}
    x_name::Symbol
    u_name::Symbol
    x_dim::Int
    dim::Int                      # x_dim * (N - 1)            integrators.jl:309
    nnz_jac::Int
    nnz_hess::Int
    ncols::Int                    # traj.dim * traj.N + traj.global_dim   integrators.jl:780-782
    function B200BilinearIntegrator(kind, G0::Matrix{Float64}, Gj::Vector{Matrix{Float64}},
                                    traj, x_name::Symbol, u_name::Symbol; device = 0)
        b = size(G0, 1)
        n_b = kind == PB2_UNITARY ? b ÷ 2 : 1
        Gjflat = isempty(Gj) ? zeros(1) : reduce(vcat, vec.(Gj))
        comps = traj.components
        desc = PB2Desc(kind, b, n_b, length(Gj), traj.N, traj.dim,
                       first(comps[x_name]) - 1, first(comps[traj.timestep]) - 1, first(comps[u_name]) - 1,
                       traj.global_dim, 0, device, 0, pointer(G0), pointer(Gjflat))
        h = Ref{Ptr{Cvoid}}(C_NULL)
        GC.@preserve G0 Gjflat check(ccall((:pb2_create, LIB), Cint, (Ref{PB2Desc}, Ref{Ptr{Cvoid}}), desc, h))
        B = new(h[], x_name, u_name, b * n_b,
                ccall((:pb2_dim, LIB), Int64, (Ptr{Cvoid},), h[]),
                ccall((:pb2_nnz_jac, LIB), Int64, (Ptr{Cvoid},), h[]),
                ccall((:pb2_nnz_hess, LIB), Int64, (Ptr{Cvoid},), h[]),
                traj.dim * traj.N + traj.global_dim)
        finalizer(B -> ccall((:pb2_destroy, LIB), Cvoid, (Ptr{Cvoid},), B.handle), B)
        return B
    end
end

# ---- generator factors: exactly what the reference's closures add up per knot --------------------
# G(u) = G_drift + Σ_j u_j G_drives[j]                   src/quantum/systems/quantum_systems.jl:212-227
generator_parts(sys::QuantumSystem) = (Matrix(sys.G_drift), [Matrix(G) for G in sys.G_drives])
# compact Lindbladian factors P·(G(ad_vec H) [+ Σ iso_D(L)])·L   open_quantum_systems.jl:541-562
function generator_parts(sys::OpenQuantumSystem)
    𝒢d, 𝒢s, _ = Piccolo.compact_lindbladian_parts(sys)
    return Matrix(𝒢d), [Matrix(G) for G in 𝒢s]
end

linear_drives_only(sys) = all(d -> d isa Piccolo.LinearDrive, sys.drives)   # drives.jl:93-99

# ---- constructors mirroring src/control/integrators.jl:35-95 ---------------------------------------
# Anything the C ABI cannot express (nonlinear / time-dependent drives) falls through to the
# reference's own integrator, decided here at construction time.
function Piccolo.BilinearIntegrator(qtraj::UnitaryTrajectory, traj::NamedTrajectory, ::Val{:b200})
    sys = qtraj.system
    (sys.time_dependent || !linear_drives_only(sys)) && return BilinearIntegrator(qtraj, traj.N)
    G0, Gj = generator_parts(sys)
    B200BilinearIntegrator(PB2_UNITARY, G0, Gj, traj, Piccolo.state_name(qtraj), Piccolo.drive_name(qtraj))
end
function Piccolo.BilinearIntegrator(qtraj::KetTrajectory, traj::NamedTrajectory, ::Val{:b200})
    sys = qtraj.system
    (sys.time_dependent || !linear_drives_only(sys)) && return BilinearIntegrator(qtraj, traj.N)
    G0, Gj = generator_parts(sys)
    B200BilinearIntegrator(PB2_KET, G0, Gj, traj, Piccolo.state_name(qtraj), Piccolo.drive_name(qtraj))
end
function Piccolo.BilinearIntegrator(qtraj::DensityTrajectory, traj::NamedTrajectory, ::Val{:b200})
    sys = qtraj.system
    (sys.time_dependent || !linear_drives_only(sys)) && return BilinearIntegrator(qtraj, traj.N)
    G0, Gj = generator_parts(sys)
    B200BilinearIntegrator(PB2_DENSITY, G0, Gj, traj, Piccolo.state_name(qtraj), Piccolo.drive_name(qtraj))
end

# ---- the AbstractIntegrator interface --------------------------------------------------------------
function DirectTrajOpt.evaluate!(δ::AbstractVector{Float64}, B::B200BilinearIntegrator, traj::NamedTrajectory)
    check(ccall((:pb2_residual, LIB), Cint, (Ptr{Cvoid}, Ptr{Float64}, Ptr{Float64}, Cint),
                B.handle, traj.datavec, δ, PB2_HOST))
    return nothing
end

function DirectTrajOpt.jacobian_structure(B::B200BilinearIntegrator)
    rows = Vector{Int64}(undef, B.nnz_jac); cols = similar(rows)
    check(ccall((:pb2_structure_jac, LIB), Cint, (Ptr{Cvoid}, Ptr{Int64}, Ptr{Int64}), B.handle, rows, cols))
    return collect(zip(rows, cols))          # (row, col) tuples, 1-based: test/test_utils.jl:17-30
end

function DirectTrajOpt.jacobian!(vals::AbstractVector{Float64}, B::B200BilinearIntegrator, traj::NamedTrajectory)
    check(ccall((:pb2_jacobian, LIB), Cint, (Ptr{Cvoid}, Ptr{Float64}, Ptr{Float64}, Cint),
                B.handle, traj.datavec, vals, PB2_HOST))
    return nothing
end

function DirectTrajOpt.eval_jacobian(B::B200BilinearIntegrator, traj::NamedTrajectory)
    rows = Vector{Int64}(undef, B.nnz_jac); cols = similar(rows); vals = Vector{Float64}(undef, B.nnz_jac)
    check(ccall((:pb2_structure_jac, LIB), Cint, (Ptr{Cvoid}, Ptr{Int64}, Ptr{Int64}), B.handle, rows, cols))
    DirectTrajOpt.jacobian!(vals, B, traj)
    return sparse(rows, cols, vals, B.dim, B.ncols)       # size pinned at integrators.jl:780-782
end

function DirectTrajOpt.hessian_structure(B::B200BilinearIntegrator)
    rows = Vector{Int64}(undef, B.nnz_hess); cols = similar(rows)
    check(ccall((:pb2_structure_hess, LIB), Cint, (Ptr{Cvoid}, Ptr{Int64}, Ptr{Int64}), B.handle, rows, cols))
    return collect(zip(rows, cols))
end

function DirectTrajOpt.hessian_of_lagrangian!(vals::AbstractVector{Float64}, B::B200BilinearIntegrator,
                                              traj::NamedTrajectory, μ::AbstractVector{Float64})
    check(ccall((:pb2_hess_lagrangian, LIB), Cint, (Ptr{Cvoid}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Cint),
                B.handle, traj.datavec, μ, vals, PB2_HOST))
    return nothing
end

# ---- DerivativeIntegrator(x, ẋ, traj): x_{k+1} - x_k - Δt_k ẋ_k  (smooth_pulse_problem.jl:267-275) -----
# One pair per Julia object (DirectTrajOpt keeps one integrator per pair); pb2_aux_* can also evaluate all
# pairs of a trajectory plus the time-consistency rows in a single launch.
struct PB2AuxDesc
    K::Int32; D::Int32; dt_off::Int32; t_off::Int32; global_dim::Int32; n_pairs::Int32
    x_off::NTuple{8,Int32}; xdot_off::NTuple{8,Int32}; dim::NTuple{8,Int32}
    device::Int32
end

mutable struct B200DerivativeIntegrator <: AbstractIntegrator
    handle::Ptr{Cvoid}
    x_name::Symbol
    xdot_name::Symbol
    dim::Int
    function B200DerivativeIntegrator(x::Symbol, ẋ::Symbol, traj::NamedTrajectory; device = 0)
        comps = traj.components
        pad(v) = ntuple(i -> i == 1 ? Int32(v) : Int32(0), 8)
        desc = PB2AuxDesc(traj.N, traj.dim, first(comps[traj.timestep]) - 1, -1, traj.global_dim, 1,
                          pad(first(comps[x]) - 1), pad(first(comps[ẋ]) - 1), pad(length(comps[x])), device)
        h = Ref{Ptr{Cvoid}}(C_NULL)
        check(ccall((:pb2_aux_create, LIB), Cint, (Ref{PB2AuxDesc}, Ref{Ptr{Cvoid}}), desc, h))
        B = new(h[], x, ẋ, ccall((:pb2_aux_dim, LIB), Int64, (Ptr{Cvoid},), h[]))
        finalizer(B -> ccall((:pb2_aux_destroy, LIB), Cvoid, (Ptr{Cvoid},), B.handle), B)
        return B
    end
end

function DirectTrajOpt.evaluate!(δ::AbstractVector{Float64}, B::B200DerivativeIntegrator, traj::NamedTrajectory)
    check(ccall((:pb2_aux_residual_jacobian, LIB), Cint, (Ptr{Cvoid}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Cint),
                B.handle, traj.datavec, δ, C_NULL, PB2_HOST))
    return nothing
end

function DirectTrajOpt.jacobian!(vals::AbstractVector{Float64}, B::B200DerivativeIntegrator, traj::NamedTrajectory)
    check(ccall((:pb2_aux_residual_jacobian, LIB), Cint, (Ptr{Cvoid}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Cint),
                B.handle, traj.datavec, C_NULL, vals, PB2_HOST))
    return nothing
end

# ---- objective value + gradient (pb2_obj_*) -----------------------------------------------------------
# One device handle for a whole sum of objectives.  A term is the real form of one of
# src/control/objectives.jl's losses (see include/piccolo_b200.h); the constructors below rewrite the
# complex goal as coefficient vectors exactly as piccolo.jl_b200/objectives.py does.
struct PB2ObjTerm
    flags::Int32
    n_rows::Int32
    rows::Ptr{Int32}
    a_re::Ptr{Float64}
    a_im::Ptr{Float64}
    a_sq::Ptr{Float64}
    a_lin::Ptr{Float64}
    scale::Float64
    n_times::Int32
    times::Ptr{Int32}
    Q::Ptr{Float64}
end

struct PB2ObjReg
    n_rows::Int32
    rows::Ptr{Int32}
    R::Ptr{Float64}
    baseline::Ptr{Float64}
    dt_power::Int32
    n_times::Int32
    times::Ptr{Int32}
end

struct PB2ObjDesc
    K::Int32
    D::Int32
    dt_off::Int32
    n_terms::Int32
    n_regs::Int32
    terms::Ptr{PB2ObjTerm}
    regs::Ptr{PB2ObjReg}
    device::Int32
end

mutable struct B200Objective <: DirectTrajOpt.Objectives.AbstractObjective
    handle::Ptr{Cvoid}
    n::Int                      # traj.dim * traj.N
end

# ⟨goal|ψ⟩ for ψ̃ = [Re ψ; Im ψ]
overlap_coeffs(g::AbstractVector{<:Complex}) = (vcat(real(g), imag(g)), vcat(-imag(g), real(g)))

"""
    B200UnitaryInfidelityObjective(U_goal, name, traj; Q, R = (u = 1e-2, ...), dt_power = 0)

`UnitaryInfidelityObjective(U_goal, name, traj; Q) + Σ QuadraticRegularizer(sym, traj, R[sym])`
(smooth_pulse_problem.jl:240-250) as one device handle.
"""
function B200UnitaryInfidelityObjective(U_goal::AbstractMatrix{<:Complex}, name::Symbol, traj::NamedTrajectory;
                                        Q = 100.0, R = NamedTuple(), dt_power = 0, device = 0)
    n = size(U_goal, 1)
    rows = Int32.(collect(traj.components[name]) .- 1)
    a_re = zeros(length(rows)); a_im = zeros(length(rows))
    for c = 1:n
        r, i = overlap_coeffs(U_goal[:, c])
        a_re[(2n*(c-1)+1):(2n*c)] = r
        a_im[(2n*(c-1)+1):(2n*c)] = i
    end
    Qv = [Float64(Q)]
    reg_rows = [Int32.(collect(traj.components[s]) .- 1) for s in keys(R)]
    reg_R = [fill(Float64(R[s]), length(traj.components[s])) for s in keys(R)]
    GC.@preserve rows a_re a_im Qv reg_rows reg_R begin
        terms = [PB2ObjTerm(1, length(rows), pointer(rows), pointer(a_re), pointer(a_im), C_NULL, C_NULL,
                            1 / n^2, 0, C_NULL, pointer(Qv))]
        regs = [PB2ObjReg(length(reg_rows[i]), pointer(reg_rows[i]), pointer(reg_R[i]), C_NULL, dt_power, 0, C_NULL)
                for i in eachindex(reg_rows)]
        GC.@preserve terms regs begin
            desc = PB2ObjDesc(traj.N, traj.dim, first(traj.components[traj.timestep]) - 1, 1, length(regs),
                              pointer(terms), isempty(regs) ? C_NULL : pointer(regs), device)
            h = Ref{Ptr{Cvoid}}(C_NULL)
            check(ccall((:pb2_obj_create, LIB), Cint, (Ref{PB2ObjDesc}, Ref{Ptr{Cvoid}}), desc, h))
        end
    end
    J = B200Objective(h[], traj.dim * traj.N)
    finalizer(J -> ccall((:pb2_obj_destroy, LIB), Cvoid, (Ptr{Cvoid},), J.handle), J)
    return J
end

function DirectTrajOpt.objective_value(J::B200Objective, traj::NamedTrajectory)
    v = Ref{Float64}(0.0)
    check(ccall((:pb2_obj_value_gradient, LIB), Cint, (Ptr{Cvoid}, Ptr{Float64}, Ref{Float64}, Ptr{Float64}, Cint),
                J.handle, traj.datavec, v, C_NULL, PB2_HOST))
    return v[]
end

function DirectTrajOpt.gradient!(∇::AbstractVector{Float64}, J::B200Objective, traj::NamedTrajectory)
    v = Ref{Float64}(0.0)
    fill!(view(∇, (J.n+1):length(∇)), 0.0)          # global variables: no dependence
    check(ccall((:pb2_obj_value_gradient, LIB), Cint, (Ptr{Cvoid}, Ptr{Float64}, Ref{Float64}, Ptr{Float64}, Cint),
                J.handle, traj.datavec, v, ∇, PB2_HOST))
    return nothing
end

# ---- registry hook: `integrator = "b200_bilinear"` in a ProblemSpec ---------------------------------
# factory signature (qtraj, N; alg) -> integrator            src/specs/materialize.jl:216-222
function __init__()
    Piccolo.Specs.register_integrator!(:b200_bilinear,
        Piccolo.Specs.RegistryEntry(factory = (qtraj, N; alg = nothing) ->
            BilinearIntegrator(qtraj, NamedTrajectory(qtraj, N), Val(:b200))))
end

end # module
