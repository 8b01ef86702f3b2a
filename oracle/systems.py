"""System builders and generator extraction (oracle restatement; test infrastructure).

Follows the reference's constructors only as far as needed to produce the dense
generator factors (G_drift, G_drives[j]) the knot path consumes:

  QuantumSystem(H_drift, H_drives, bounds)      src/quantum/systems/quantum_systems.jl:192-249
      G(u) = G(H_drift) + sum_j u_j G(H_j)      :212-227
  CompositeQuantumSystem                        src/quantum/systems/composite_quantum_systems.jl:92-154
  lift_operator                                 src/quantum/operators/lifted_operators.jl:22-31
  annihilate                                    src/quantum/object_utils.jl:154
  TransmonSystem / TransmonDipoleCoupling / MultiTransmonSystem
                                                src/quantum/templates/transmons/transmon_system.jl:34-96,139-171,199-263
  CatSystem                                     src/quantum/templates/cats/cat_system.jl:54-125
  compact_lindbladian_parts                     src/quantum/systems/open_quantum_systems.jl:541-562
  compact_generator_closure (LinearDrive / LinearDissipator only)   :607-636
"""
from dataclasses import dataclass, field
from typing import List

import numpy as np

from . import isomorphisms as iso

PAULI_X = np.array([[0, 1], [1, 0]], dtype=complex)
PAULI_Y = np.array([[0, -1j], [1j, 0]], dtype=complex)
PAULI_Z = np.array([[1, 0], [0, -1]], dtype=complex)
GATE_H = np.array([[1, 1], [1, -1]], dtype=complex) / np.sqrt(2)
GATE_CX = np.array(
    [[1, 0, 0, 0], [0, 1, 0, 0], [0, 0, 0, 1], [0, 0, 1, 0]], dtype=complex
)


def annihilate(levels):
    return np.diag(np.sqrt(np.arange(1, levels)), k=1).astype(complex)


def lift_operator(op, i, subsystem_levels):
    """i is 0-based here; subsystem i is the i-th Kronecker factor from the left."""
    mats = [np.eye(l, dtype=complex) for l in subsystem_levels]
    mats[i] = np.asarray(op, dtype=complex)
    out = mats[0]
    for M in mats[1:]:
        out = np.kron(out, M)
    return out


@dataclass
class QuantumSystem:
    H_drift: np.ndarray
    H_drives: List[np.ndarray]
    drive_bounds: List[float]
    dissipation_operators: List[np.ndarray] = field(default_factory=list)
    dissipation_rates: List[float] = field(default_factory=list)

    @property
    def levels(self):
        return self.H_drift.shape[0]

    @property
    def n_drives(self):
        return len(self.H_drives)

    @property
    def is_open(self):
        return len(self.dissipation_operators) > 0

    def H(self, u):
        out = self.H_drift.astype(complex).copy()
        for j, Hj in enumerate(self.H_drives):
            out = out + u[j] * Hj
        return out

    # closed-system generator factors, 2d x 2d real
    def G_parts(self):
        return iso.G(self.H_drift), [iso.G(Hj) for Hj in self.H_drives]


def OpenQuantumSystem(H_drift, H_drives, drive_bounds, dissipation_operators=(), rates=None):
    Ls = [np.asarray(L, dtype=complex) for L in dissipation_operators]
    if rates is None:
        rates = [1.0] * len(Ls)  # LinearDissipator default rate (dissipators.jl)
    return QuantumSystem(
        np.asarray(H_drift, dtype=complex),
        [np.asarray(H, dtype=complex) for H in H_drives],
        list(drive_bounds),
        Ls,
        list(rates),
    )


def compact_lindbladian_parts(sys):
    """(Gc_drift_ham, [Gc_drive_i], [Gc_dissipator_j]), each P * M * L, d^2 x d^2 real."""
    n = sys.levels
    P = iso.density_projection_matrix(n)
    L = iso.density_lift_matrix(n)
    drift = P @ iso.G(iso.ad_vec(sys.H_drift)) @ L
    drives = [P @ iso.G(iso.ad_vec(Hj)) @ L for Hj in sys.H_drives]
    diss = [P @ iso.iso_D(Lj) @ L for Lj in sys.dissipation_operators]
    return drift, drives, diss


def compact_generator_parts(sys):
    """Fold constant-rate (LinearDissipator) terms into the drift: (G0, [Gj])."""
    drift, drives, diss = compact_lindbladian_parts(sys)
    G0 = drift.copy()
    for rate, Dj in zip(sys.dissipation_rates, diss):
        G0 = G0 + rate * Dj
    return G0, drives


# ----------------------------------------------------------------------------- #
# templates
# ----------------------------------------------------------------------------- #

def TransmonSystem(omega=4.0, delta=0.2, levels=3, frame_omega=None, drives=True,
                   drive_bounds=(1.0, 1.0)):
    """Rotating-frame duffing transmon, multiplied by 2 pi (transmon_system.jl:34-96)."""
    if frame_omega is None:
        frame_omega = omega
    a = annihilate(levels)
    ad = a.conj().T
    H_drift = (omega - frame_omega) * ad @ a - delta / 2 * ad @ ad @ a @ a
    H_drives = [a + ad, 1j * (a - ad)] if drives else []
    return QuantumSystem(2 * np.pi * H_drift, [2 * np.pi * H for H in H_drives],
                         list(drive_bounds) if drives else [])


def MultiTransmonSystem(omegas, deltas, gs, levels_per_transmon=3, drive_bounds=1.0,
                        subsystem_drive_indices=None):
    """transmon_system.jl:199-263 (rotating frame) + composite_quantum_systems.jl:92-154."""
    n = len(omegas)
    if subsystem_drive_indices is None:
        subsystem_drive_indices = list(range(n))
    lv = [levels_per_transmon] * n
    subs = [
        TransmonSystem(omega=w, delta=dl, levels=levels_per_transmon,
                       drives=(i in subsystem_drive_indices),
                       drive_bounds=(drive_bounds, drive_bounds))
        for i, (w, dl) in enumerate(zip(omegas, deltas))
    ]
    d = int(np.prod(lv))
    H_drift = np.zeros((d, d), dtype=complex)
    gs = np.asarray(gs, dtype=float)
    for i in range(n - 1):
        for j in range(i + 1, n):
            ai = lift_operator(annihilate(lv[i]), i, lv)
            aj = lift_operator(annihilate(lv[j]), j, lv)
            H_drift = H_drift + 2 * np.pi * gs[i, j] * (ai @ aj.conj().T + ai.conj().T @ aj)
    H_drives, bounds = [], []
    for i, s in enumerate(subs):
        H_drift = H_drift + lift_operator(s.H_drift, i, lv)
        for Hd in s.H_drives:
            H_drives.append(lift_operator(Hd, i, lv))
        bounds += s.drive_bounds
    return QuantumSystem(H_drift, H_drives, bounds)


def CatSystem(g2=0.36, chi_aa=-7e-3, chi_bb=-32.0, chi_ab=0.79, kappa_a=53e-3, kappa_b=13.0,
              cat_levels=13, buffer_levels=3, prefactor=1.0, drive_bounds=(1.0, 1.0)):
    """cat_system.jl:54-125."""
    g2, chi_aa, chi_bb, chi_ab = (prefactor * v for v in (g2, chi_aa, chi_bb, chi_ab))
    kappa_a, kappa_b = prefactor * kappa_a, prefactor * kappa_b
    a = np.kron(annihilate(cat_levels), np.eye(buffer_levels))
    b = np.kron(np.eye(cat_levels), annihilate(buffer_levels))
    ad, bd = a.conj().T, b.conj().T
    H_drift = (-chi_aa / 2 * ad @ ad @ a @ a - chi_bb / 2 * bd @ bd @ b @ b
               - chi_ab * ad @ a @ bd @ b + g2 * ad @ ad @ b + np.conj(g2) * a @ a @ bd)
    H_drives = [b + bd, ad @ a]
    Ls = [np.sqrt(kappa_a) * a, np.sqrt(kappa_b) * b]
    return OpenQuantumSystem(2 * np.pi * H_drift, [2 * np.pi * H for H in H_drives],
                             list(drive_bounds),
                             [np.sqrt(2 * np.pi) * L for L in Ls])


def coherent_ket(alpha, levels):
    from math import factorial
    return np.array([np.exp(-0.5 * abs(alpha) ** 2) * alpha ** n / np.sqrt(factorial(n))
                     for n in range(levels)], dtype=complex)
