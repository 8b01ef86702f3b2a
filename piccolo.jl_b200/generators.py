"""Generator factors the knot kernels consume: dense real G_drift, G_drives[j].

Host-side setup code (runs once per problem, not per callback).  Same definitions as the
reference's /root/reference/src/quantum/primitives/isomorphisms.jl -- iso :350, G :359,
ad_vec :384-387, iso_D :394-396, density_lift_matrix :236-277, density_projection_matrix
:292-321 -- and open_quantum_systems.jl:541-562 (compact_lindbladian_parts), written with
index arithmetic instead of the reference's push!-loops.
"""
import numpy as np


def iso(H):
    """[[Re H, -Im H], [Im H, Re H]]"""
    H = np.asarray(H, dtype=complex)
    return np.block([[H.real, -H.imag], [H.imag, H.real]])


def G(H):
    """iso(-iH) = [[Im H, Re H], [-Re H, Im H]]"""
    H = np.asarray(H, dtype=complex)
    return np.block([[H.imag, H.real], [-H.real, H.imag]])


def ad_vec(H, anti=False):
    H = np.asarray(H, dtype=complex)
    Id = np.eye(H.shape[0])
    return np.kron(Id, H) + (1.0 if anti else -1.0) * np.kron(H.T, Id)


def iso_D(L):
    L = np.asarray(L, dtype=complex)
    return iso(np.kron(L.conj(), L) - 0.5 * ad_vec(L.conj().T @ L, anti=True))


def _compact_index(n):
    """(j, k) pairs of the compact density iso: Re upper triangle then Im strict upper, col-major."""
    re = [(j, k) for k in range(n) for j in range(k + 1)]
    im = [(j, k) for k in range(1, n) for j in range(k)]
    return re, im


def density_lift_matrix(n):
    re, im = _compact_index(n)
    L = np.zeros((2 * n * n, n * n))
    for c, (j, k) in enumerate(re):
        L[k * n + j, c] = 1.0
        L[j * n + k, c] = 1.0
    for c, (j, k) in enumerate(im, start=len(re)):
        L[n * n + k * n + j, c] = 1.0
        L[n * n + j * n + k, c] = -1.0
    return L


def density_projection_matrix(n):
    re, im = _compact_index(n)
    P = np.zeros((n * n, 2 * n * n))
    for r, (j, k) in enumerate(re):
        P[r, k * n + j] = 1.0
    for r, (j, k) in enumerate(im, start=len(re)):
        P[r, n * n + k * n + j] = 1.0
    return P


def compact_lindbladian_parts(H_drift, H_drives, dissipation_operators):
    n = np.asarray(H_drift).shape[0]
    P, L = density_projection_matrix(n), density_lift_matrix(n)
    drift = P @ G(ad_vec(H_drift)) @ L
    drives = [P @ G(ad_vec(Hj)) @ L for Hj in H_drives]
    diss = [P @ iso_D(Lj) @ L for Lj in dissipation_operators]
    return drift, drives, diss


def compact_generator_parts(H_drift, H_drives, dissipation_operators, rates=None):
    """(G0, [Gj]) with constant-rate dissipators folded into the drift
    (compact_generator_closure, open_quantum_systems.jl:607-636, LinearDrive/LinearDissipator case)."""
    drift, drives, diss = compact_lindbladian_parts(H_drift, H_drives, dissipation_operators)
    rates = [1.0] * len(diss) if rates is None else list(rates)
    for r, Dj in zip(rates, diss):
        drift = drift + r * Dj
    return drift, drives
