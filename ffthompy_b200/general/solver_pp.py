"""Solver callbacks with the interface of ffthompy/general/solver_pp.py.

`CallBack(A=Afun, B=B)` is what ffthompy/applications.py:71-72 attaches to every solve: called with the iterate after
each iteration (and once with x0 by `linear_solver`), it appends the TRUE residual norm ||B - A(x)|| to `res_norm`.
Here the operator application inside it is the fused device pipeline, and the CG loop that calls it stays on the
device path too (general/solver.py: one fh_cg_steps call per iteration)."""
import numpy as np

from ..tensors import Tensor


def _as_tensor(x, like):
    if isinstance(x, np.ndarray):    # SciPy solvers hand over flat host vectors
        return Tensor(val=np.reshape(x, like._vshape()), order=like.order, N=like.N, Y=like.Y)
    return x


class CallBack(object):
    """true-residual history (general/solver_pp.py:6-34): attributes A (operator), B (right-hand side)"""

    def __init__(self, **kwargs):
        vars(self).update(kwargs)
        self.iter = -1
        self.res_norm = []
        self.energy_norm = []

    def __call__(self, x):
        self.iter += 1
        r = self.B-self.A(_as_tensor(x, self.B))
        self.res_norm.append(r.norm())

    def __repr__(self):
        if not self.res_norm:
            return 'the results are not initialized yet'
        return '    iterations : %d\n    res_norm : %g\n' % (self.iter, self.res_norm[-1])


class CallBack_GA(object):
    """per-iteration residual, energy bound and non-conformity (general/solver_pp.py:37-75; reachable in the
    reference only through pb.solver['callback'] == 'detailed').  Attributes: A, B, GN (projection operator),
    E2N / EN (macroscopic field), Aex / A_Ga (exactly integrated coefficients)."""

    def __init__(self, **kwargs):
        vars(self).update(kwargs)
        self.iter = -1
        self.res_norm = []
        self.bound = []
        self.nonconformity = []

    def _attr(self, *names):
        for nm in names:
            if hasattr(self, nm):
                return getattr(self, nm)
        raise AttributeError('CallBack_GA needs one of %s' % (names,))

    def __call__(self, x):
        self.iter += 1
        macro = self._attr('E2N', 'EN')
        A_exact = self._attr('Aex', 'A_Ga')
        has_macro = np.linalg.norm(x.mean()-macro.mean()) < 1e-8   # iterate already carries the macroscopic part
        r = self.A(x) if has_macro else self.B-self.A(x)
        e = x if has_macro else x+macro
        self.res_norm.append(r.norm())
        conforming = self.GN(e)+macro
        shifted = conforming+macro
        self.bound.append(A_exact(shifted)*shifted)
        self.nonconformity.append((conforming-e).norm())

    def __repr__(self):
        if not self.res_norm:
            return 'no output'
        return ('    iterations    : %d\n    res_norm      : %g\n    bound         : %g\n    nonconformity : %g'
                % (self.iter, self.res_norm[-1], self.bound[-1], self.nonconformity[-1]))
