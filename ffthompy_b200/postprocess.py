"""Assembly of the homogenised matrix from the minimisers — drop-in for the device-relevant part
of ffthompy/postprocess.py (assembly_matrix, add_macro2minimizer).  `postprocess(pb, ...)` itself
is the reference's driver glue around Material and stays with the caller."""
import itertools

import numpy as np


def assembly_matrix(Afun, solutions):
    """A_H[i, j] = <A e_i, e_j> (postprocess.py:53-70); GaNi solutions are spectrally interpolated to
    the grid of the coefficients first."""
    dim = len(solutions)
    if not np.allclose(Afun.N, solutions[0].N):
        Nbar = Afun.N
        sol = []
        for ii in np.arange(dim):
            sol.append(solutions[ii].project(Nbar))
    else:
        sol = solutions

    AH = np.zeros([dim, dim])
    Asol = [Afun(s) for s in sol]  # each A e_i once (the reference recomputes it dim times)
    for ii, jj in itertools.product(list(range(dim)), repeat=2):
        AH[ii, jj] = Asol[ii] * sol[jj]
    return AH


def add_macro2minimizer(X, E):
    """postprocess.py:73-86"""
    if np.allclose(X.mean(), E):
        return X
    elif np.allclose(X.mean(), np.zeros_like(E)):
        EN = X.zeros_like(name='EN')
        EN.set_mean(E)
        return X + EN
    else:
        raise ValueError("Field is neither zero-mean nor E-mean.")
