"""Host-side (NumPy) grid and pseudo-inverse formulae used at setup time.
Only the pieces on the time-step path are provided:
gauss_lobatto (pypde/bases/dmsuite.py:209-214) and pseudoinverse_spectral
(:327-379).  Values are bit-identical to the reference's: same closed forms,
evaluated entry by entry in Python floats."""
import numpy as np
import scipy.sparse as sp


def gauss_lobatto(n):
    """Chebyshev-Gauss-Lobatto points, ascending from -1 to 1 (n + 1 points)."""
    k = np.linspace(n, 0, n + 1)
    if n > 1:
        return np.sin(np.pi * (n - 2 * k) / (2 * n))
    return 0


def pseudoinverse_diagonals(N, deriv=2):
    """Diagonals of the pseudo-inverse of the spectral derivative matrix D^deriv
    (Sahuck Oh's preconditioner).  Returns (offsets, [diag arrays])."""
    if deriv == 2:
        lo = np.array([1 / (4 * i * (i - 1)) for i in range(2, N)], dtype=float)
        lo[0] *= 2
        mid = np.zeros(N)
        mid[2:-2] = np.array([-1 / (2 * (i ** 2 - 1)) for i in range(2, N - 2)])
        up = np.zeros(N - 2)
        up[2:-2] = np.array([1 / (4 * i * (i + 1)) for i in range(2, N - 4)])
        return (-2, 0, 2), [lo, mid, up]
    if deriv == 1:
        lo = np.array([1 / (2 * i) for i in range(1, N)], dtype=float)
        lo[0] *= 2
        up = np.zeros(N - 1)
        up[1:-1] = np.array([-1 / (2 * i) for i in range(1, N - 2)])
        return (-1, 1), [lo, up]
    raise ValueError("pseudoinverse_spectral does only support deriv==1 or 2")


def pseudoinverse_sparse(N, deriv=2):
    offs, dg = pseudoinverse_diagonals(N, deriv)
    return sp.diags(dg, list(offs), shape=(N, N), format="csr")


def pseudoinverse_spectral(N, deriv=2):
    """Dense N x N array, same signature as the reference."""
    return pseudoinverse_sparse(N, deriv).toarray()
