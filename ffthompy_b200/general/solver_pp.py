"""Solver callbacks — drop-in for ffthompy/general/solver_pp.py (CallBack)."""
import numpy as np

from ..tensors import Tensor


class CallBack():
    """Records the true residual norm ||B - A(x)|| at every call (general/solver_pp.py:6-34).
    Note the extra operator application per iteration this implies."""

    def __init__(self, **kwargs):
        self.__dict__.update(kwargs)
        self.iter = -1
        self.res_norm = []
        self.energy_norm = []

    def __call__(self, x):
        self.iter += 1
        if isinstance(x, np.ndarray):
            X = Tensor(val=np.reshape(x, self.B._vshape()), order=self.B.order, N=self.B.N, Y=self.B.Y)
        else:
            X = x
        res = self.B - self.A(X)
        self.res_norm.append(res.norm())
        return

    def __repr__(self):
        try:
            ss = ''
            ss += '    iterations : %d\n' % self.iter
            ss += '    res_norm : %g' % self.res_norm[-1]
            ss += '\n'
        except Exception:
            ss = 'the results are not initialized yet'
        return ss


class CallBack_GA():
    """Detailed callback (general/solver_pp.py:37-75; legacy in the reference — it is only reachable
    through pb.solver['callback'] == 'detailed').  Kept importable and functional over the Tensor
    algebra: residual norm, energy bound and non-conformity per iteration."""

    def __init__(self, **kwargs):
        self.__dict__.update(kwargs)
        self.iter = -1
        self.res_norm = []
        self.bound = []
        self.nonconformity = []

    def __call__(self, x):
        self.iter += 1
        X = x
        E2N = getattr(self, 'E2N', getattr(self, 'EN', None))
        Aex = getattr(self, 'Aex', getattr(self, 'A_Ga', None))
        if np.linalg.norm(X.mean() - E2N.mean()) < 1e-8:
            res = self.A(X)
            eN = X
        else:
            res = self.B-self.A(X)
            eN = X + E2N
        self.res_norm.append(res.norm())
        GeN = self.GN(eN) + E2N
        GeN_E = GeN + E2N
        self.bound.append(Aex(GeN_E)*GeN_E)
        self.nonconformity.append((GeN-eN).norm())
        return

    def __repr__(self):
        try:
            ss = ''
            ss += '    iterations    : %d\n' % self.iter
            ss += '    res_norm      : %g\n' % self.res_norm[-1]
            ss += '    bound         : %g\n' % self.bound[-1]
            ss += '    nonconformity : %g' % self.bound[-1]
        except Exception:
            ss = 'no output'
        return ss
