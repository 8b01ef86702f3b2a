"""Real isomorphisms of quantum objects (oracle restatement; test infrastructure).

Follows /root/reference/src/quantum/primitives/isomorphisms.jl:
  ket_to_iso :55, iso_to_ket :62, operator_to_iso_vec :110-118,
  iso_vec_to_operator :73-81, density_to_iso_vec :153, density_to_compact_iso :176-191,
  density_lift_matrix :236-277, density_projection_matrix :292-321,
  iso :350, G :359, H :368-373, ad_vec :384-387, iso_D :394-396.
"""
import numpy as np


def ket_to_iso(psi):
    psi = np.asarray(psi, dtype=complex)
    return np.concatenate([psi.real, psi.imag])


def iso_to_ket(x):
    x = np.asarray(x, dtype=float)
    n = x.size // 2
    return x[:n] + 1j * x[n:]


def operator_to_iso_vec(U):
    """Stack, for each column i of U, [Re U[:,i]; Im U[:,i]]."""
    U = np.asarray(U, dtype=complex)
    n = U.shape[0]
    out = np.empty(2 * n * U.shape[1])
    for i in range(U.shape[1]):
        out[i * 2 * n : i * 2 * n + n] = U[:, i].real
        out[i * 2 * n + n : (i + 1) * 2 * n] = U[:, i].imag
    return out


def iso_vec_to_operator(x):
    x = np.asarray(x, dtype=float)
    n = int(round(np.sqrt(x.size // 2)))
    U = np.empty((n, n), dtype=complex)
    for i in range(n):
        U[:, i] = x[i * 2 * n : i * 2 * n + n] + 1j * x[i * 2 * n + n : (i + 1) * 2 * n]
    return U


def density_to_iso_vec(rho):
    rho = np.asarray(rho, dtype=complex)
    return ket_to_iso(rho.reshape(-1, order="F"))


def density_to_compact_iso(rho):
    rho = np.asarray(rho, dtype=complex)
    n = rho.shape[0]
    x = np.empty(n * n)
    idx = 0
    for k in range(n):
        for j in range(k + 1):
            x[idx] = rho[j, k].real
            idx += 1
    for k in range(1, n):
        for j in range(k):
            x[idx] = rho[j, k].imag
            idx += 1
    return x


def compact_iso_to_density(x):
    x = np.asarray(x, dtype=float)
    n = int(round(np.sqrt(x.size)))
    rho = np.zeros((n, n), dtype=complex)
    idx = 0
    for k in range(n):
        for j in range(k + 1):
            rho[j, k] = x[idx]
            if j != k:
                rho[k, j] = x[idx]
            idx += 1
    for k in range(1, n):
        for j in range(k):
            rho[j, k] += 1j * x[idx]
            rho[k, j] -= 1j * x[idx]
            idx += 1
    return rho


def density_lift_matrix(n):
    n2 = n * n
    L = np.zeros((2 * n2, n2))
    col = 0
    for k in range(n):
        for j in range(k + 1):
            L[k * n + j, col] = 1.0
            if j != k:
                L[j * n + k, col] = 1.0
            col += 1
    for k in range(1, n):
        for j in range(k):
            L[n2 + k * n + j, col] = 1.0
            L[n2 + j * n + k, col] = -1.0
            col += 1
    return L


def density_projection_matrix(n):
    n2 = n * n
    P = np.zeros((n2, 2 * n2))
    row = 0
    for k in range(n):
        for j in range(k + 1):
            P[row, k * n + j] = 1.0
            row += 1
    for k in range(1, n):
        for j in range(k):
            P[row, n2 + k * n + j] = 1.0
            row += 1
    return P


_IM2 = np.array([[0.0, -1.0], [1.0, 0.0]])


def iso(H):
    H = np.asarray(H, dtype=complex)
    return np.kron(np.eye(2), H.real) + np.kron(_IM2, H.imag)


def G(H):
    """iso(-iH) = [[Im H, Re H], [-Re H, Im H]]."""
    return iso(-1j * np.asarray(H, dtype=complex))


def H_of_G(Gm):
    Gm = np.asarray(Gm, dtype=float)
    d = Gm.shape[0] // 2
    return -Gm[d:, :d] + 1j * Gm[:d, :d]


def ad_vec(H, anti=False):
    """I (x) H - (-1)^anti * H^T (x) I   (conj(H)' in Julia is the plain transpose)."""
    H = np.asarray(H, dtype=complex)
    Id = np.eye(H.shape[0])
    sign = -1.0 if anti else 1.0
    return np.kron(Id, H) - sign * np.kron(H.T, Id)


def iso_D(L):
    L = np.asarray(L, dtype=complex)
    LdL = L.conj().T @ L
    return iso(np.kron(L.conj(), L) - 0.5 * ad_vec(LdL, anti=True))
