# PiccoloB200.jl -- thin `ccall` shim that plugs libpiccolo_b200.so (include/piccolo_b200.h) into
# Piccolo.jl's existing integrator extension point.
#
# NOT EXECUTED IN THIS REPOSITORY'S CI (field and helper names below were checked against the reference
# source by reading it, e.g. QuantumSystem.H_drives / drive_matrix / compact_lindbladian_generators): the build image has no Julia toolchain and DirectTrajOpt.jl
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
    t_off::Int32                # 0-based row of :t (time-dependent handles only)
    time_dependent::Int32       # != 0: drive j enters as c_j(t_k) u_j, coefficients via pb2_set_time_coefficients
    dense_blocks::Int32         # != 0: d/dx_k as one dense x_dim × x_dim block (SURVEY.md:388)
end

const PB2_OPT_EARLY_Z, PB2_OPT_PIPELINED = Int32(1), Int32(2)

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
    modulations::Vector{Any}      # (c_j, ċ_j) closures of ModulatedDrive (drives.jl:342-388); empty = time-independent
    function B200BilinearIntegrator(kind, G0::Matrix{Float64}, Gj::Vector{Matrix{Float64}},
                                    traj, x_name::Symbol, u_name::Symbol; device = 0, n_states = 1,
                                    modulations = Any[], dense_blocks = false)
        b = size(G0, 1)
        # n_b contiguous state blocks share the generator: d columns of a unitary, or the n_states
        # kets / densities of a multi-state trajectory fused into one integrator (x_name = the first)
        n_b = kind == PB2_UNITARY ? b ÷ 2 : n_states
        Gjflat = isempty(Gj) ? zeros(1) : reduce(vcat, vec.(Gj))
        comps = traj.components
        desc = PB2Desc(kind, b, n_b, length(Gj), traj.N, traj.dim,
                       first(comps[x_name]) - 1, first(comps[traj.timestep]) - 1, first(comps[u_name]) - 1,
                       traj.global_dim, 0, device, 0, pointer(G0), pointer(Gjflat),
                       isempty(modulations) ? 0 : first(comps[:t]) - 1, isempty(modulations) ? 0 : 1,
                       dense_blocks ? 1 : 0)
        h = Ref{Ptr{Cvoid}}(C_NULL)
        GC.@preserve G0 Gjflat check(ccall((:pb2_create, LIB), Cint, (Ref{PB2Desc}, Ref{Ptr{Cvoid}}), desc, h))
        B = new(h[], x_name, u_name, b * n_b,
                ccall((:pb2_dim, LIB), Int64, (Ptr{Cvoid},), h[]),
                ccall((:pb2_nnz_jac, LIB), Int64, (Ptr{Cvoid},), h[]),
                ccall((:pb2_nnz_hess, LIB), Int64, (Ptr{Cvoid},), h[]),
                traj.dim * traj.N + traj.global_dim, collect(Any, modulations))
        finalizer(B -> ccall((:pb2_destroy, LIB), Cvoid, (Ptr{Cvoid},), B.handle), B)
        return B
    end
end

"Promise that lets consecutive *_async callbacks overlap on the GPU (include/piccolo_b200.h, pb2_set_option)."
set_option!(B::B200BilinearIntegrator, opt::Int32, value::Integer = 1) =
    check(ccall((:pb2_set_option, LIB), Cint, (Ptr{Cvoid}, Cint, Int64), B.handle, opt, value))

# Time-dependent (carrier-modulated) drives: the closures c_j(t) stay here; the library gets their values and
# derivatives at the knot times before each evaluation (integrators.jl:38-46, drives.jl:342-388).
function upload_time_coefficients!(B::B200BilinearIntegrator, traj::NamedTrajectory)
    isempty(B.modulations) && return
    t = vec(traj[:t]); m = length(B.modulations)
    c = [B.modulations[j][1](t[k]) for j in 1:m, k in 1:traj.N]                    # m × N, column-major
    ċ = [B.modulations[j][2](t[k]) for j in 1:m, k in 1:traj.N]                    # ModulatedDrive.modulation_deriv
    check(ccall((:pb2_set_time_coefficients, LIB), Cint, (Ptr{Cvoid}, Ptr{Float64}, Ptr{Float64}, Cint),
                B.handle, c, ċ, PB2_HOST))
end

# ---- generator factors: exactly what the reference's closures add up per knot --------------------
# G(u) = G(H_drift) + Σ_j u_j G(H_j).  The systems store the Hamiltonian terms (H_drift, H_drives::Vector of
# AbstractDrive: src/quantum/systems/quantum_systems.jl:60-71, open_quantum_systems.jl:30-41), not generator
# matrices; the isomorphism is Isomorphisms.G (src/quantum/primitives/isomorphisms.jl:350,359) as in the
# closure built at quantum_systems.jl:212-227.
generator_parts(sys::QuantumSystem) = (
    Matrix{Float64}(Piccolo.Isomorphisms.G(sys.H_drift)),
    [Matrix{Float64}(Piccolo.Isomorphisms.G(Piccolo.drive_matrix(base_drive(d)))) for d in sys.H_drives],
)
# Compact Lindbladian: the reference's integrator (integrators.jl:82-95) adds Σ_j rate_j 𝒢c_dissipators[j] to
# the Hamiltonian drift through compact_generator_closure (open_quantum_systems.jl:607-636).  For
# LinearDissipators the rate is a constant, so it folds into the drift factor exactly;
# compact_lindbladian_generators (open_quantum_systems.jl:576-590) is the reference's own helper that does it.
function generator_parts(sys::OpenQuantumSystem)
    𝒢d, 𝒢s = Piccolo.compact_lindbladian_generators(sys)
    return Matrix{Float64}(𝒢d), [Matrix{Float64}(G) for G in 𝒢s]
end

# what the C ABI can express: coefficients u[index] (LinearDrive, drives.jl:52-55) and constant dissipation
# rates (LinearDissipator, dissipators.jl:33-39); anything else keeps the reference's integrator
linear_drives_only(sys) = all(d -> d isa Piccolo.LinearDrive, sys.H_drives)
# ... and carrier-modulated linear drives, ModulatedDrive(LinearDrive(H_j, j), c_j) (drives.jl:342-388): the
# generator stays separable, Ĝ(u, t) = G₀ + Σ_j c_j(t) u_j G_j, which the time-dependent handles evaluate
separable_drive(d) = d isa Piccolo.LinearDrive || (d isa Piccolo.ModulatedDrive && d.base isa Piccolo.LinearDrive)
separable_drives_only(sys::QuantumSystem) = all(separable_drive, sys.H_drives)
base_drive(d) = d isa Piccolo.ModulatedDrive ? d.base : d
# (c_j, ċ_j) per drive; the identity for unmodulated ones.  Empty when nothing is modulated.
function drive_modulations(sys::QuantumSystem)
    any(d -> d isa Piccolo.ModulatedDrive, sys.H_drives) || return Any[]
    return Any[d isa Piccolo.ModulatedDrive ? (d.modulation, d.modulation_deriv) : (t -> 1.0, t -> 0.0) for d in sys.H_drives]
end
linear_drives_only(sys::OpenQuantumSystem) =
    all(d -> d isa Piccolo.LinearDrive, sys.H_drives) &&
    !Piccolo.has_nonlinear_dissipators(getfield(sys, :dissipators))

# ---- constructors mirroring src/control/integrators.jl:35-95 ---------------------------------------
# Anything the C ABI cannot express (nonlinear / time-dependent drives) falls through to the
# reference's own integrator, decided here at construction time.
function Piccolo.BilinearIntegrator(qtraj::UnitaryTrajectory, traj::NamedTrajectory, ::Val{:b200})
    sys = qtraj.system
    separable_drives_only(sys) || return BilinearIntegrator(qtraj, traj.N)
    G0, Gj = generator_parts(sys)
    B200BilinearIntegrator(PB2_UNITARY, G0, Gj, traj, Piccolo.state_name(qtraj), Piccolo.drive_name(qtraj);
                           modulations = drive_modulations(sys))
end
function Piccolo.BilinearIntegrator(qtraj::KetTrajectory, traj::NamedTrajectory, ::Val{:b200})
    sys = qtraj.system
    separable_drives_only(sys) || return BilinearIntegrator(qtraj, traj.N)
    G0, Gj = generator_parts(sys)
    B200BilinearIntegrator(PB2_KET, G0, Gj, traj, Piccolo.state_name(qtraj), Piccolo.drive_name(qtraj);
                           modulations = drive_modulations(sys))
end
function Piccolo.BilinearIntegrator(qtraj::DensityTrajectory, traj::NamedTrajectory, ::Val{:b200})
    sys = qtraj.system
    (sys.time_dependent || !linear_drives_only(sys)) && return BilinearIntegrator(qtraj, traj.N)
    G0, Gj = generator_parts(sys)
    B200BilinearIntegrator(PB2_DENSITY, G0, Gj, traj, Piccolo.state_name(qtraj), Piccolo.drive_name(qtraj))
end

# MultiKetTrajectory / SamplingTrajectory: a vector of integrators, one per state component, all reading
# the same control rows (integrators.jl:102-117, 134-226) ...
function Piccolo.BilinearIntegrator(qtraj::MultiKetTrajectory, traj::NamedTrajectory, ::Val{:b200}; fused = false)
    sys = qtraj.system
    (sys.time_dependent || !linear_drives_only(sys)) && return BilinearIntegrator(qtraj, traj.N)
    G0, Gj = generator_parts(sys)
    names = Piccolo.state_names(qtraj)
    # ... or, since every ket obeys the same generator and the blocks are contiguous in the knot column,
    # ONE integrator that evaluates them all in a single launch (rows knot-major, state-major inside a knot)
    fused && return B200BilinearIntegrator(PB2_KET, G0, Gj, traj, names[1], Piccolo.drive_name(qtraj); n_states = length(names))
    return [B200BilinearIntegrator(PB2_KET, G0, Gj, traj, nm, Piccolo.drive_name(qtraj)) for nm in names]
end
function Piccolo.BilinearIntegrator(qtraj::SamplingTrajectory, traj::NamedTrajectory, ::Val{:b200})
    base = qtraj.base_trajectory
    kind = base isa UnitaryTrajectory ? PB2_UNITARY : (base isa Union{KetTrajectory,MultiKetTrajectory} ? PB2_KET : PB2_DENSITY)
    members = Piccolo.sampling_member_states(qtraj)          # Symbol or Vector{Symbol} per member (member-major)
    out = AbstractIntegrator[]
    for (sys, states) in zip(qtraj.systems, members)
        (sys.time_dependent || !linear_drives_only(sys)) && return BilinearIntegrator(qtraj, traj.N)
        G0, Gj = generator_parts(sys)
        for nm in (states isa AbstractVector ? states : [states])
            push!(out, B200BilinearIntegrator(kind, G0, Gj, traj, nm, Piccolo.drive_name(qtraj)))
        end
    end
    return out
end

# All member integrators of an ensemble in ONE launch (pb2_batch_*): same constructor shape as above, one object.
mutable struct B200IntegratorBatch
    handle::Ptr{Cvoid}
    names::Vector{Symbol}
    dim::Int; nnz_jac::Int; nnz_hess::Int        # per member
end
function B200IntegratorBatch(kind, gens::Vector, traj::NamedTrajectory, names::Vector{Symbol}, u_name::Symbol; device = 0)
    comps = traj.components
    keep = Any[]
    descs = map(zip(gens, names)) do ((G0, Gj), nm)
        Gjflat = isempty(Gj) ? zeros(1) : reduce(vcat, vec.(Gj)); push!(keep, G0, Gjflat)
        b = size(G0, 1)
        PB2Desc(kind, b, kind == PB2_UNITARY ? b ÷ 2 : 1, length(Gj), traj.N, traj.dim, first(comps[nm]) - 1,
                first(comps[traj.timestep]) - 1, first(comps[u_name]) - 1, traj.global_dim, 0, device, 0,
                pointer(G0), pointer(Gjflat), 0, 0, 0)
    end
    h = Ref{Ptr{Cvoid}}(C_NULL)
    GC.@preserve keep check(ccall((:pb2_batch_create, LIB), Cint, (Ptr{PB2Desc}, Cint, Ref{Ptr{Cvoid}}), descs, length(descs), h))
    B = B200IntegratorBatch(h[], names, ccall((:pb2_batch_dim, LIB), Int64, (Ptr{Cvoid},), h[]),
                            ccall((:pb2_batch_nnz_jac, LIB), Int64, (Ptr{Cvoid},), h[]),
                            ccall((:pb2_batch_nnz_hess, LIB), Int64, (Ptr{Cvoid},), h[]))
    finalizer(B -> ccall((:pb2_batch_destroy, LIB), Cvoid, (Ptr{Cvoid},), B.handle), B)
    return B
end
"δ (dim × members) and Jacobian values (nnz_jac × members), member i = what member i's own integrator returns"
function residual_jacobian!(δ::Matrix{Float64}, vals::Matrix{Float64}, B::B200IntegratorBatch, traj::NamedTrajectory)
    check(ccall((:pb2_batch_residual_jacobian, LIB), Cint, (Ptr{Cvoid}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Cint),
                B.handle, traj.datavec, δ, vals, PB2_HOST))
end

# rollout!(qtraj, pulse) for the trajectory's own zero-order-hold controls + rollout_divergence
# (rollouts_extensions.jl:46-92, problems.jl:186-208, 336-356): states x_dim × N, out = (ε, ‖Δx_N‖, ‖x_N‖)
function rollout(B::B200BilinearIntegrator, traj::NamedTrajectory; x0 = nothing)
    states = Matrix{Float64}(undef, B.x_dim, traj.N); out = zeros(3)
    check(ccall((:pb2_rollout, LIB), Cint, (Ptr{Cvoid}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Cint),
                B.handle, traj.datavec, x0 === nothing ? C_NULL : x0, states, out, PB2_HOST))
    return states, out[1]
end

# ---- the AbstractIntegrator interface --------------------------------------------------------------
function DirectTrajOpt.evaluate!(δ::AbstractVector{Float64}, B::B200BilinearIntegrator, traj::NamedTrajectory)
    upload_time_coefficients!(B, traj)
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
    upload_time_coefficients!(B, traj)
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
    timesteps_all_equal::Int32      # != 0: K-1 rows Δt_{k+1} - Δt_k (TimeStepsAllEqualConstraint, _problem_templates.jl:175-180)
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
                          pad(first(comps[x]) - 1), pad(first(comps[ẋ]) - 1), pad(length(comps[x])), device, 0)
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

# Host-side description of a term / regularizer (kept alive by the objective; the library copies them
# at pb2_obj_create).  Vectors are empty when absent.
struct ObjTerm
    flags::Int32
    rows::Vector{Int32}                 # 0-based rows of the knot column
    a_re::Vector{Float64}; a_im::Vector{Float64}; a_sq::Vector{Float64}; a_lin::Vector{Float64}
    scale::Float64
    times::Vector{Int32}                # 0-based knots; empty = terminal knot
    Q::Vector{Float64}
end
struct ObjReg
    rows::Vector{Int32}
    R::Vector{Float64}
    baseline::Matrix{Float64}           # rows × N, or 0 × 0
    dt_power::Int32
    times::Vector{Int32}                # empty = every knot
end

mutable struct B200Objective <: DirectTrajOpt.Objectives.AbstractObjective
    terms::Vector{ObjTerm}
    regs::Vector{ObjReg}
    layout::NTuple{4,Int}               # (N, dim, 0-based Δt row, device)
    handle::Ptr{Cvoid}                  # built on first use, so that `J1 + J2 + …` costs nothing
end
B200Objective(terms, regs, traj::NamedTrajectory; device = 0) =
    B200Objective(terms, regs, (traj.N, traj.dim, first(traj.components[traj.timestep]) - 1, device), C_NULL)

function Base.:+(a::B200Objective, b::B200Objective)
    a.layout == b.layout || error("objectives were built on different trajectory layouts")
    return B200Objective(vcat(a.terms, b.terms), vcat(a.regs, b.regs), a.layout, C_NULL)
end

ptr_or_null(v::Array) = isempty(v) ? Ptr{eltype(v)}(C_NULL) : pointer(v)

function handle!(J::B200Objective)
    J.handle == C_NULL || return J.handle
    N, dim, dt_off, device = J.layout
    GC.@preserve J begin
        terms = [PB2ObjTerm(t.flags, length(t.rows), ptr_or_null(t.rows), ptr_or_null(t.a_re), ptr_or_null(t.a_im),
                            ptr_or_null(t.a_sq), ptr_or_null(t.a_lin), t.scale, length(t.times), ptr_or_null(t.times),
                            pointer(t.Q)) for t in J.terms]
        regs = [PB2ObjReg(length(r.rows), ptr_or_null(r.rows), pointer(r.R), ptr_or_null(r.baseline), r.dt_power,
                          length(r.times), ptr_or_null(r.times)) for r in J.regs]
        GC.@preserve terms regs begin
            desc = PB2ObjDesc(N, dim, dt_off, length(terms), length(regs), ptr_or_null(terms), ptr_or_null(regs), device)
            h = Ref{Ptr{Cvoid}}(C_NULL)
            check(ccall((:pb2_obj_create, LIB), Cint, (Ref{PB2ObjDesc}, Ref{Ptr{Cvoid}}), desc, h))
            J.handle = h[]
        end
    end
    finalizer(J -> ccall((:pb2_obj_destroy, LIB), Cvoid, (Ptr{Cvoid},), J.handle), J)
    return J.handle
end

rows0(traj, name) = Int32.(collect(traj.components[name]) .- 1)

# ⟨goal|ψ⟩ for ψ̃ = [Re ψ; Im ψ]:  Re = a_re·ψ̃, Im = a_im·ψ̃
overlap_coeffs(g::AbstractVector{<:Complex}) = (vcat(real(g), imag(g)), vcat(-imag(g), real(g)))
const NOVEC, NOTIMES = Float64[], Int32[]
const ONE_MINUS = Int32(1)

"KetInfidelityObjective(ψ_goal, name, traj; Q)   src/control/objectives.jl:56-64"
function B200KetInfidelityObjective(ψ_goal::AbstractVector{<:Complex}, name::Symbol, traj::NamedTrajectory; Q = 100.0, device = 0)
    a_re, a_im = overlap_coeffs(ComplexF64.(ψ_goal))
    return B200Objective([ObjTerm(ONE_MINUS, rows0(traj, name), a_re, a_im, NOVEC, NOVEC, 1.0, NOTIMES, [Float64(Q)])],
                         ObjReg[], traj; device)
end

"CoherentKetInfidelityObjective(goals, names, traj; Q, weights)   objectives.jl:181-216 (weights normalised as :136-144)"
function B200CoherentKetInfidelityObjective(ψ_goals, names::Vector{Symbol}, traj::NamedTrajectory;
                                            Q = 100.0, weights = nothing, device = 0)
    n = length(ψ_goals)
    ws = Piccolo.QuantumObjectives.coherent_fidelity_weights(weights, n)
    w = isnothing(ws) ? fill(1 / n, n) : ws
    parts = [overlap_coeffs(ComplexF64.(g)) for g in ψ_goals]
    rows = reduce(vcat, [rows0(traj, nm) for nm in names])
    a_re = reduce(vcat, [w[i] .* parts[i][1] for i = 1:n])
    a_im = reduce(vcat, [w[i] .* parts[i][2] for i = 1:n])
    return B200Objective([ObjTerm(ONE_MINUS, rows, a_re, a_im, NOVEC, NOVEC, 1.0, NOTIMES, [Float64(Q)])], ObjReg[], traj; device)
end

"UnitaryInfidelityObjective(U_goal, name, traj; Q)   objectives.jl:330-337, 347-356"
function B200UnitaryInfidelityObjective(U_goal::AbstractMatrix{<:Complex}, name::Symbol, traj::NamedTrajectory;
                                        Q = 100.0, device = 0)
    n = size(U_goal, 1)
    parts = [overlap_coeffs(ComplexF64.(U_goal[:, c])) for c = 1:n]
    a_re = reduce(vcat, first.(parts)); a_im = reduce(vcat, last.(parts))
    return B200Objective([ObjTerm(ONE_MINUS, rows0(traj, name), a_re, a_im, NOVEC, NOVEC, 1 / n^2, NOTIMES, [Float64(Q)])],
                         ObjReg[], traj; device)
end

"UnitaryInfidelityObjective(op::EmbeddedOperator, …)   objectives.jl:339-345 (unitary subspace goal)"
function B200UnitaryInfidelityObjective(op::EmbeddedOperator, name::Symbol, traj::NamedTrajectory; Q = 100.0, device = 0)
    Ug = ComplexF64.(unembed(op)); sub = op.subspace; n = length(sub)
    Nl = isqrt(length(traj.components[name]) ÷ 2)
    a_re = zeros(2Nl^2); a_im = zeros(2Nl^2); a_sq = zeros(2Nl^2)
    for (p, c) in enumerate(sub), (q, i) in enumerate(sub)
        g = Ug[q, p]; re = 2Nl * (c - 1) + i; im = re + Nl
        a_re[re], a_re[im] = real(g), imag(g)
        a_im[re], a_im[im] = -imag(g), real(g)
        a_sq[re] = a_sq[im] = 1.0
    end
    return B200Objective([ObjTerm(ONE_MINUS, rows0(traj, name), a_re, a_im, a_sq, NOVEC, 1 / (n * (n + 1)), NOTIMES, [Float64(Q)])],
                         ObjReg[], traj; device)
end

# Re tr(ρ W) = a·x for ρ = compact_iso_to_density(x)   (isomorphisms.jl:176-191 ordering)
function compact_trace_coeffs(W::AbstractMatrix{<:Complex})
    n = size(W, 1); a = Float64[]
    for k = 1:n, j = 1:k
        push!(a, j == k ? real(W[k, k]) : real(W[k, j] + W[j, k]))
    end
    for k = 2:n, j = 1:(k-1)
        push!(a, imag(W[j, k] - W[k, j]))
    end
    return a
end

"DensityMatrixInfidelityObjective(name, ρ_goal, traj; Q)   objectives.jl:387-411"
B200DensityMatrixInfidelityObjective(name::Symbol, ρ_goal::AbstractMatrix{<:Complex}, traj::NamedTrajectory; Q = 100.0, device = 0) =
    B200Objective([ObjTerm(ONE_MINUS, rows0(traj, name), NOVEC, NOVEC, NOVEC, compact_trace_coeffs(ρ_goal), 1.0, NOTIMES, [Float64(Q)])],
                  ObjReg[], traj; device)

"DensityMatrixPureStateInfidelityObjective(name, ψ_goal, traj; Q)   objectives.jl:412-429"
B200DensityMatrixPureStateInfidelityObjective(name::Symbol, ψ::AbstractVector{<:Complex}, traj::NamedTrajectory; Q = 100.0, device = 0) =
    B200DensityMatrixInfidelityObjective(name, ψ * ψ', traj; Q, device)

"LeakageObjective(indices, name, traj; times, Qs)   objectives.jl:464-474"
function B200LeakageObjective(indices::AbstractVector{Int}, name::Symbol, traj::NamedTrajectory;
                              times = 1:traj.N, Qs = fill(1.0, length(times)), device = 0)
    rows = rows0(traj, name)[indices]
    return B200Objective([ObjTerm(Int32(0), rows, NOVEC, NOVEC, fill(1 / length(indices), length(rows)), NOVEC, 1.0,
                                  Int32.(collect(times) .- 1), Float64.(Qs))], ObjReg[], traj; device)
end

"QuadraticRegularizer(name, traj, R; baseline, times)  (DirectTrajOpt; Δt power: see include/piccolo_b200.h)"
function B200QuadraticRegularizer(name::Symbol, traj::NamedTrajectory, R; baseline = zeros(0, 0), times = Int[],
                                  dt_power = 0, device = 0)
    rows = rows0(traj, name)
    Rv = R isa Number ? fill(Float64(R), length(rows)) : Float64.(R)
    return B200Objective(ObjTerm[], [ObjReg(rows, Rv, Matrix{Float64}(baseline), Int32(dt_power), Int32.(collect(times) .- 1))],
                         traj; device)
end

function DirectTrajOpt.objective_value(J::B200Objective, traj::NamedTrajectory)
    v = Ref{Float64}(0.0)
    check(ccall((:pb2_obj_value_gradient, LIB), Cint, (Ptr{Cvoid}, Ptr{Float64}, Ref{Float64}, Ptr{Float64}, Cint),
                handle!(J), traj.datavec, v, C_NULL, PB2_HOST))
    return v[]
end

function DirectTrajOpt.gradient!(∇::AbstractVector{Float64}, J::B200Objective, traj::NamedTrajectory)
    v = Ref{Float64}(0.0)
    fill!(view(∇, (J.layout[1]*J.layout[2]+1):length(∇)), 0.0)          # global variables: no dependence
    check(ccall((:pb2_obj_value_gradient, LIB), Cint, (Ptr{Cvoid}, Ptr{Float64}, Ref{Float64}, Ptr{Float64}, Cint),
                handle!(J), traj.datavec, v, ∇, PB2_HOST))
    return nothing
end

# objective block of eval_h: σ ∇²J as COO (upper triangle, duplicates sum) -- pb2_obj_hessian
function DirectTrajOpt.hessian_structure(J::B200Objective)
    n = ccall((:pb2_obj_nnz_hess, LIB), Int64, (Ptr{Cvoid},), handle!(J))
    rows = Vector{Int64}(undef, n); cols = similar(rows)
    check(ccall((:pb2_obj_structure_hess, LIB), Cint, (Ptr{Cvoid}, Ptr{Int64}, Ptr{Int64}), J.handle, rows, cols))
    return collect(zip(rows, cols))
end
function hessian_values!(vals::AbstractVector{Float64}, J::B200Objective, traj::NamedTrajectory, σ::Float64)
    check(ccall((:pb2_obj_hessian, LIB), Cint, (Ptr{Cvoid}, Ptr{Float64}, Cdouble, Ptr{Float64}, Cint),
                handle!(J), traj.datavec, σ, vals, PB2_HOST))
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
