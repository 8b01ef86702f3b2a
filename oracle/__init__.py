"""CPU oracle for the direct-collocation knot path.  TEST INFRASTRUCTURE ONLY.

This package is a CPU restatement (NumPy/SciPy + a C++ port under ``oracle/c``)
of the arithmetic the reference evaluates per knot:

    delta_k = x_{k+1} - exp(dt_k * G(u_k)) x_k          (docs/src/concepts/index.md:21,62)

with G(u) built from the reference's isomorphisms
(src/quantum/primitives/isomorphisms.jl) and system closures
(src/quantum/systems/quantum_systems.jl:212-227, open_quantum_systems.jl:541-636).

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs may import it, and only as the checker.  The shipped
product (``piccolo.jl_b200``) never imports anything from here.

Parity status: PINNED for the residual (four converged reference trajectories
under tests/golden/ satisfy delta ~ 0 with this restatement, plus the
reference's isomorphism known-answer tests).  Jacobian/Hessian VALUES are pinned
only as analytic derivatives of that pinned residual (cross-checked against
finite differences and a second, independent algorithm in oracle/c).  The COO
index ORDER is defined by DirectTrajOpt.jl, whose source is not vendored in the
reference: index order is "parity unpinned" and documented as ours.
"""
