"""Rebuild the KnotProblem for each committed golden trajectory (oracle side)."""
import json
import os

import numpy as np

from oracle import knot as KN
from oracle import systems as S

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
META = json.load(open(os.path.join(GOLDEN, "golden_meta.json")))
TIGHT = {  # max |delta| the reference's converged solution leaves, with headroom
    "two_qubit_zoh": 1e-10,
    "trajectories_density": 1e-12,
    "systems_cat_density": 1e-12,
    "trajectories_ket": 1e-8,
}


def load(name):
    if name == "trajectories_multi":
        raise ValueError("multi-state fixture: use load_multi()")
    Z = np.load(os.path.join(GOLDEN, name + ".npz"))["Z"]
    Z = np.asfortranarray(Z)
    K = META[name]["K"]
    if name == "two_qubit_zoh":
        s = S.MultiTransmonSystem([4.0, 4.1], [0.2, 0.2], [[0, 0.1], [0.1, 0]],
                                  levels_per_transmon=2, drive_bounds=0.1)
        G0, Gj = s.G_parts()
        p = KN.make_problem("unitary", G0, Gj, K)
    elif name == "trajectories_density":
        s = S.OpenQuantumSystem(S.PAULI_Z, [S.PAULI_X, S.PAULI_Y], [1, 1],
                                [np.array([[0.1, 0], [0, 0]])])
        G0, Gj = S.compact_generator_parts(s)
        p = KN.make_problem("density", G0, Gj, K)
    elif name == "systems_cat_density":
        s = S.CatSystem(cat_levels=3, buffer_levels=2)
        G0, Gj = S.compact_generator_parts(s)
        p = KN.make_problem("density", G0, Gj, K)
    else:
        s = S.QuantumSystem(S.PAULI_Z, [S.PAULI_X, S.PAULI_Y], [1, 1])
        G0, Gj = s.G_parts()
        p = KN.make_problem("ket" if name == "trajectories_ket" else "unitary", G0, Gj, K)
    assert p.D == Z.shape[0] and p.K == Z.shape[1]
    return p, Z


def load_multi():
    """The reference's MultiKetTrajectory solution: (list of per-state KnotProblems, Z).  Every state
    block has its own integrator; they share the Δt / u rows (integrators.jl:102-117)."""
    import dataclasses
    Z = np.asfortranarray(np.load(os.path.join(GOLDEN, "trajectories_multi.npz"))["Z"])
    meta = META["trajectories_multi"]
    s = S.QuantumSystem(S.PAULI_Z, [S.PAULI_X, S.PAULI_Y], [1, 1])
    G0, Gj = s.G_parts()
    base = KN.make_problem("ket", G0, Gj, meta["K"])
    n_s, b = meta["n_states"], base.b
    probs = [dataclasses.replace(base, D=Z.shape[0], x_off=i * b, dt_off=n_s * b, u_off=n_s * b + 2)
             for i in range(n_s)]
    return probs, Z
