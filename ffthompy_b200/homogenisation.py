"""Displacement-based (potential) Fourier-Galerkin homogenisation of scalar problems — drop-in for the full-tensor
solvers of ffthompy/tensorsLowRank/homogenisation.py:13-128 (homog_Ga_full, homog_Ga_full_potential,
homog_GaNi_full_potential) and their preconditioner (:296-300).

The unknown is the Fourier-space potential u^ on the solve grid N (one complex component instead of d real ones on
the doubled grid), the operator of one CG iteration is

    u^  ->  P u^  ->  grad  ->  enlarge to Nbar  ->  iF  ->  A(x) .  ->  F  ->  project to N  ->  -div  ->  P

with P = 1/|2 pi xi| the diagonal preconditioner and the reference's 'r'-weighted Fourier scalar product
(tensors/objects.py:618-636) in the CG.  This is the only place where enlarge / decrease are per-iteration work
(SURVEY 3.5): every arrow is one device kernel of this package (fh_grad, fh_spec_remap, fh_irfftn, fh_mul21,
fh_rfftn, fh_div, fh_hadamard), nothing touches the host between them."""
import numpy as np

from .general.base import Timer
from .general.solver import linear_solver
from .tensors import DFT, Operator, Tensor, grad, div, grad_tensor
from .trigpol import mean_index
from .tensors import projection as proj   # the variant WITHOUT Nyquist zeroing, as imported by the reference module


class Struct(dict):
    """attribute-style parameter bag (ffthompy.Struct): pars.solver, pars.Y, ..."""
    __getattr__ = dict.__getitem__
    __setattr__ = dict.__setitem__


def _solve_grid(Nbar):
    return np.array((np.array(Nbar)+1)//2, dtype=int)


def _unit_load(N, dim):
    E = np.zeros(dim)
    E[0] = 1
    EN = Tensor(name='EN', N=N, shape=(dim,), Fourier=False)
    EN.set_mean(E)
    return EN


def get_preconditioner(N, pars):
    """P(xi) = 1/|2 pi xi|, P(0) = 1: real order-0 multiplier in Fourier space (homogenisation.py:296-300)"""
    hGrad = grad_tensor(N, pars.Y)
    k2 = np.einsum('i...,i...', hGrad.val, np.conj(hGrad.val)).real
    k2[mean_index(N)] = 1.
    return Tensor(name='P', val=1./k2**0.5, order=0, N=N, Fourier=True, multype=00)


def homog_Ga_full(Aga, pars):
    """gradient-field formulation on the doubled grid (homogenisation.py:13-39).  The projection comes from
    tensors/projection.py (no Nyquist zeroing): on odd solve grids the operator is the fused G.A pipeline with the
    device CG; on even ones the enlarged multiplier carries split Nyquist planes and is applied materialised."""
    Nbar = Aga.N
    N = _solve_grid(Nbar)
    dim = len(Nbar)
    _, Ghat, _ = proj.scalar(N, np.ones(dim))
    G1N = Operator(name='G1', mat=[[DFT(name='FiN', inverse=True, N=Nbar), Ghat.enlarge(Nbar),
                                    DFT(name='FN', inverse=False, N=Nbar)]])
    PAfun = Operator(name='FiGFA', mat=[[G1N, Aga]])
    EN = _unit_load(Nbar, dim)
    watch = Timer(name='CG (gradient field)')
    X, info = linear_solver(solver='CG', Afun=PAfun, B=PAfun(-EN), x0=Tensor(N=Nbar, shape=(dim,), Fourier=False),
                            par=pars.solver, callback=None)
    watch.measure()
    e = X+EN
    return Struct(AH=Aga(e)*e, X=X, info=info, time=watch.vals[0][0], pars=pars)


def _potential_solve(apply_flux_div, rhs, N, pars):
    """preconditioned CG on the Fourier-space potential: returns (F u, info, seconds)"""
    P = get_preconditioner(N, pars)
    operator = lambda Fx: P*apply_flux_div(P*Fx)   # noqa: E731
    watch = Timer(name='CG (potential)')
    iPU, info = linear_solver(solver='CG', Afun=operator, B=P*rhs, x0=Tensor(N=N, shape=(), Fourier=True),
                              par=pars.solver, callback=None)
    watch.measure()
    print('iterations of CG={}'.format(info['kit']))
    print('norm of residuum={}'.format(info['norm_res']))
    return P*iPU, info, watch.vals[0][0]


def homog_Ga_full_potential(Aga, pars):
    """exact-integration (Ga) coefficients on Nbar = 2N-1, potential on N (homogenisation.py:41-84)"""
    Nbar = Aga.N
    N = _solve_grid(Nbar)
    dim = len(Nbar)
    F2 = DFT(name='FN', inverse=False, N=Nbar)
    iF2 = DFT(name='FiN', inverse=True, N=Nbar)
    EN = _unit_load(Nbar, dim)

    def flux_divergence(X):
        assert(X.Fourier)
        return -div(F2(Aga*iF2(grad(X).enlarge(Nbar))).project(N))

    Fu, info, seconds = _potential_solve(flux_divergence, div(F2(Aga(EN)).decrease(N)), N, pars)
    e = iF2(grad(Fu).project(Nbar))
    return Struct(AH=Aga(e+EN)*(e+EN), e=e, Fu=Fu, info=info, time=seconds)


def homog_GaNi_full_potential(Agani, Aga, pars):
    """numerical-integration (GaNi) coefficients on N; the minimiser is evaluated with the Ga coefficients on the
    doubled grid when `Aga` is given (homogenisation.py:86-128)"""
    N = Agani.N
    dim = len(N)
    F = DFT(name='FN', inverse=False, N=N)
    iF = DFT(name='FiN', inverse=True, N=N)
    EN = _unit_load(N, dim)

    def flux_divergence(X):
        assert(X.Fourier)
        return -div(F(Agani*iF(grad(X))))

    Fu, info, seconds = _potential_solve(flux_divergence, div(F(Agani(EN))), N, pars)
    if Aga is None:
        print('!!!!! homogenised properties are GaNi only !!!!!')
        e = iF(grad(Fu))+EN
        AH = Agani(e)*e
    else:
        Nbar = 2*np.array(N)-1
        e = DFT(name='FiN', inverse=True, N=Nbar)(grad(Fu).project(Nbar))+EN.project(Nbar)
        AH = Aga(e)*e
    return Struct(AH=AH, Fu=Fu, info=info, time=seconds, pars=pars)
