"""piccolo.jl_b200 -- B200-native evaluator for Piccolo.jl's direct-collocation knot path.

Host-side mirror (Python, because no Julia toolchain exists in this image) of the reference's
integrator plug-in interface for this path only:

    reference (Julia)                                          here
    ---------------------------------------------------------  ---------------------------------
    BilinearIntegrator(qtraj, N)   src/control/integrators.jl:35-95   BilinearIntegrator(qtraj, N)
    evaluate!(delta, B, traj)      integrators.jl:311                 evaluate_(delta, B, traj)
    eval_jacobian(B, traj)         integrators.jl:780-782             eval_jacobian(B, traj)
    jacobian_structure / hessian_structure   test/aqua.jl:6-9         same names
    test_integrator(B, traj; atol)  integrators.jl:359                test_integrator(B, traj, atol=)
    B.dim, B.x_dim, B.x_name       integrators.jl:307-309,552         same fields
    *InfidelityObjective, LeakageObjective, QuadraticRegularizer      objectives.py (src/control/objectives.jl)

All arithmetic happens in libpiccolo_b200.so (hand-written sm_100a CUDA behind the C ABI of
include/piccolo_b200.h).  There is no CPU fallback: importing works anywhere, but constructing an
integrator without the built library or without a CUDA device raises.
"""
from .capi import PB2Error, lib_path, load_library  # noqa: F401
from .generators import (  # noqa: F401
    G, ad_vec, compact_generator_parts, compact_lindbladian_parts, iso, iso_D,
    density_lift_matrix, density_projection_matrix,
)
from .integrators import (  # noqa: F401
    B200BilinearIntegrator, B200IntegratorBatch, B200KnotLinearConstraints, BilinearIntegrator, DensityTrajectory, KetTrajectory,
    MultiDensityTrajectory, MultiKetTrajectory, NamedTrajectory, OpenQuantumSystem, QuantumSystem, SamplingTrajectory, UnitaryTrajectory,
    eval_jacobian, evaluate_, hessian_of_lagrangian, hessian_structure, jacobian_structure, rollout_divergence,
    test_integrator,
)
from .objectives import (  # noqa: F401
    B200Objective, CoherentKetInfidelityObjective, DensityMatrixInfidelityObjective,
    DensityMatrixPureStateInfidelityObjective, KetInfidelityObjective, LeakageObjective, QuadraticRegularizer,
    UnitaryInfidelityObjective, gradient_, objective_value,
)
from .sharding import ShardedBilinearIntegrator, knot_partition  # noqa: F401

__all__ = [n for n in dir() if not n.startswith("_")]
