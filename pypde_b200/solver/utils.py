"""Host-side setup helper (pypde/solver/utils.py:13-29)."""
import numpy as np


def eigdecomp(A):
    """Eigen-decomposition sorted by descending eigenvalue: returns w, Q, Q^-1."""
    w, Q = np.linalg.eig(A)
    order = np.argsort(w)[::-1]
    w, Q = w[order], Q[:, order]
    return w, Q, np.linalg.inv(Q)
