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
