"""The caller of the hot path, restated over the package API: what ffthompy/applications.py:60-90
and postprocess.py:41-46 do with Tensor/DFT/Operator/linear_solver (the reference's Material and
Problem classes cannot travel to the GPU box, their outputs come from tests/golden/)."""
import numpy as np


def build_operator(A_val, G, N):
    from ffthompy_b200.tensors import Tensor, DFT, Operator
    A = Tensor(name='A', val=np.array(A_val), order=2, N=N, multype=21)
    FN = DFT(name='FN', inverse=False, N=N)
    FiN = DFT(name='FiN', inverse=True, N=N)
    GN = Operator(name='G1', mat=[[FiN, G, FN]])
    return A, Operator(name='FiGFA', mat=[[GN, A]])


def solve_loads(A_val, G, N, tol, maxiter=1e3, solver='CG', par=None, callback_factory=None):
    """unit loads -> minimisers (incl. the macroscopic part), infos; applications.py:60-83"""
    from ffthompy_b200.tensors import Tensor
    from ffthompy_b200.general.solver import linear_solver
    from ffthompy_b200.postprocess import add_macro2minimizer
    A, Afun = build_operator(A_val, G, N)
    D = A.shape[0]
    sols, infos = [], []
    for iL in range(D):
        E = np.zeros(D)
        E[iL] = 1
        EN = Tensor(name='EN', N=N, shape=(D,), Fourier=False)
        EN.set_mean(E)
        x0 = EN.zeros_like(name='x0')
        B = Afun(-EN)
        p = {'tol': tol, 'maxiter': maxiter}
        if par:
            p.update(par)
        cb = callback_factory(Afun, B) if callback_factory else None
        X, info = linear_solver(solver=solver, Afun=Afun, B=B, x0=x0, par=p, callback=cb)
        if cb is not None:
            info['cb'] = cb
        sols.append(add_macro2minimizer(X, E))
        infos.append(info)
    return A, Afun, sols, infos


def green_for(physics, kind, N, Y, primaldual):
    """applications.py:24-37,104-122: the projection used by the solve (lazy closed form)"""
    import ffthompy_b200.projections as proj
    N = np.array(N, dtype=int)
    if physics == 'scalar':
        _, G1, G2 = proj.scalar(N, Y, NyqNul=True, tensor=True)
    else:
        _, G1h, G1s, G2h, G2s = proj.elasticity(N, Y, NyqNul=True, tensor=True)
        G1, G2 = G1h+G1s, G2h+G2s
    Nbar = N if kind == 'GaNi' else 2*N-1
    if kind == 'Ga':
        G1, G2 = G1.enlarge(Nbar), G2.enlarge(Nbar)
    return (G1 if primaldual == 'primal' else G2), tuple(int(n) for n in Nbar)
