"""Krylov solvers behind `linear_solver` — drop-in for ffthompy/general/solver.py.

CG and Richardson on the fused G·A operator run as device loops (csrc/fh_fused.cu: fh_cg,
fh_richardson) — no host round trip other than the residual norm that decides termination.
Any other operator / callback / custom scalar product goes through the same algorithms
written over the Tensor algebra (each step a device kernel), exactly as in the reference.
"""
import numpy as np

from .base import Timer
from ..tensors import Tensor, Operator


def _is_vectri(B):
    try:
        from ..matvecs import VecTri
        return isinstance(B, VecTri)
    except Exception:
        return False


def linear_solver(Afun, B, ATfun=None, x0=None, par=None, solver=None, callback=None):
    """Wrapper for the linear solvers suited to FFT-based homogenisation (general/solver.py:8-60)."""
    tim = Timer('Solving linsys by %s' % solver)
    if x0 is None:
        x0 = B.zeros_like()

    if callback is not None:
        callback(x0)

    if solver.lower() in ['cg']:  # conjugate gradients
        x, info = CG(Afun, B, x0=x0, par=par, callback=callback)
    elif solver.lower() in ['bicg']:  # biconjugate gradients
        x, info = BiCG(Afun, ATfun, B, x0=x0, par=par, callback=callback)
    elif solver.lower() in ['iterative', 'richardson']:  # iterative solver
        x, info = richardson(Afun, B, x0, par=par, callback=callback)
    elif solver.lower() in ['chebyshev', 'cheby']:  # iterative solver
        x, info = cheby2TERM(A=Afun, B=B, x0=x0, par=par, callback=callback)
    elif solver.split('_')[0].lower() in ['scipy']:  # solvers in scipy (host-side bridge)
        x, info = _scipy_bridge(Afun, ATfun, B, x0, par, solver, callback)
    else:
        msg = "This kind (%s) of linear solver is not implemented" % solver
        raise NotImplementedError(msg)

    tim.measure(print_time=False)
    info.update({'time': tim.vals})
    return x, info


def _scipy_bridge(Afun, ATfun, B, x0, par, solver, callback):
    """general/solver.py:28-53: SciPy iterates on host vectors; every matvec is a device call."""
    import scipy.sparse.linalg as spslin
    x0vec = x0.ravel() if isinstance(x0, np.ndarray) else np.asarray(x0.vec()).ravel()
    Afun.define_operand(B)
    if solver in ['scipy.sparse.linalg.cg', 'scipy_cg']:
        Afunvec = spslin.LinearOperator(Afun.matshape, matvec=lambda v: np.asarray(Afun.matvec(v)).ravel(),
                                        dtype=np.float64)
        xcol, info = spslin.cg(Afunvec, np.asarray(B.vec()).ravel(), x0=x0vec, rtol=par['tol'],
                               maxiter=int(par['maxiter']), M=None, callback=callback)
    elif solver in ['scipy.sparse.linalg.bicg', 'scipy_bicg']:
        ATfun.define_operand(B)
        Afunvec = spslin.LinearOperator(Afun.matshape, matvec=lambda v: np.asarray(Afun.matvec(v)).ravel(),
                                        rmatvec=lambda v: np.asarray(ATfun.matvec(v)).ravel(), dtype=np.float64)
        xcol, info = spslin.bicg(Afunvec, np.asarray(B.vec()).ravel(), x0=x0vec, rtol=par['tol'],
                                 maxiter=int(par['maxiter']), M=None, callback=callback)
    else:
        raise NotImplementedError("This kind (%s) of linear solver is not implemented" % solver)
    x = B.empty_like(name='x')
    x.val = np.reshape(xcol, B._vshape())
    return x, {'info': info}


def _fused_for(Afun, B, x0, par, callback):
    if callback is not None or (par is not None and 'scal' in par):
        return None
    if not (isinstance(Afun, Operator) and isinstance(B, Tensor) and isinstance(x0, Tensor)):
        return None
    f = Afun.fused()
    if f is None or not (f.accepts(B) and f.accepts(x0)):
        return None
    return f


def richardson(Afun, B, x0, par=None, callback=None):
    """general/solver.py:63-77"""
    omega = 1./par['alpha']
    res = {'norm_res': 1e15,
           'kit': 0}
    f = _fused_for(Afun, B, x0, par, callback)
    if f is not None:
        xd, kit, nres = f.richardson(B._dev(), x0._dev(), par['alpha'], par['tol'], int(par['maxiter']))
        res['kit'], res['norm_res'] = kit, nres
        return x0.copy(val=xd), res
    x = x0
    norm = get_norm(B, par)
    while (res['norm_res'] > par['tol'] and res['kit'] < par['maxiter']):
        res['kit'] += 1
        residuum = B-Afun(x)
        x = x + omega*residuum
        res['norm_res'] = norm(residuum)
        if callback is not None:
            callback(x)
    return x, res


def CG(Afun, B, x0, par=None, callback=None):
    """Conjugate gradients (general/solver.py:80-139): absolute tolerance on
    sqrt(<r,r>) with the mean-normalised scalar product of the operands."""
    if par is None:
        par = dict()
    if 'tol' not in list(par.keys()):
        par['tol'] = 1e-6
    if 'maxiter' not in list(par.keys()):
        par['maxiter'] = int(1e3)

    f = _fused_for(Afun, B, x0, par, callback)
    if f is not None:
        xd, kit, nres, hist = f.cg(B._dev(), x0._dev(), par['tol'], int(par['maxiter']))
        res = {'kit': kit, 'norm_res': nres if kit > 0 else 0, 'norm_res_log': hist}
        return x0.copy(val=xd), res

    scal = get_scal(B, par)

    res = dict()
    xCG = x0
    Ax = Afun(x0)
    R = B - Ax
    P = R
    rr = scal(R, R)
    res['kit'] = 0
    res['norm_res'] = np.double(rr)**0.5  # /np.norm(E_N)
    norm_res_log = []
    norm_res_log.append(res['norm_res'])
    while (res['norm_res'] > par['tol']) and (res['kit'] < par['maxiter']):
        res['kit'] += 1  # number of iterations
        AP = Afun(P)
        alp = float(rr/scal(P, AP))
        xCG = xCG + alp*P
        R = R - alp*AP
        rrnext = scal(R, R)
        bet = rrnext/rr
        rr = rrnext
        P = R + bet*P
        res['norm_res'] = np.double(rr)**0.5
        norm_res_log.append(res['norm_res'])
        if callback is not None:
            callback(xCG)
    if res['kit'] == 0:
        res['norm_res'] = 0
    res['norm_res_log'] = np.array(norm_res_log)
    return xCG, res


def BiCG(Afun, ATfun, B, x0, par=None, callback=None):
    """BiConjugate gradients (general/solver.py:142-204), over the operand algebra."""
    if par is None:
        par = dict()
    if 'tol' not in par:
        par['tol'] = 1e-6
    if 'maxiter' not in par:
        par['maxiter'] = 1e3
    scal = get_scal(B, par)

    res = dict()
    xBiCG = x0
    Ax = Afun(x0)
    R = B - Ax
    Rs = R
    rr = float(scal(R, Rs))
    P = R
    Ps = Rs
    res['kit'] = 0
    res['norm_res'] = rr**0.5  # /np.norm(E_N)
    while (res['norm_res'] > par['tol']) and (res['kit'] < par['maxiter']):
        res['kit'] += 1  # number of iterations
        AP = Afun(P)
        alp = rr/float(scal(AP, Ps))
        xBiCG = xBiCG + alp*P
        R = R - alp*AP
        Rs = Rs - alp*ATfun(Ps)
        rrnext = float(scal(R, Rs))
        bet = rrnext/rr
        rr = rrnext
        P = R + bet*P
        Ps = Rs + bet*Ps
        res['norm_res'] = rr**0.5
        if callback is not None:
            callback(xBiCG)
    if res['kit'] == 0:
        res['norm_res'] = 0
    return xBiCG, res


def cheby2TERM(A, B, x0, M=None, par=None, callback=None):
    """Chebyshev two-term iteration (general/solver.py:206-285)."""
    if par is None:
        par = dict()
    if 'tol' not in par:
        par['tol'] = 1e-06
    if 'maxit' not in par:
        par['maxit'] = 1e7
    if 'eigrange' not in par:
        raise NotImplementedError("It is necessary to calculate eigenvalues.")
    else:
        Egv = par['eigrange']

    res = dict()
    res['kit'] = 0
    bnrm2 = (B*B)**0.5
    Ib = 1.0/bnrm2
    if bnrm2 == 0:
        bnrm2 = 1.0
    x = x0
    r = B - A(x)
    r0 = np.double(r*r)**0.5
    res['norm_res'] = Ib*r0  # For Normal Residue
    if res['norm_res'] < par['tol']:  # if errnorm is less than tol
        return x, res

    d = (Egv[1]+Egv[0])/2.0  # np.mean(par['eigrange'])
    c = (Egv[1]-Egv[0])/2.0  # par['eigrange'][1] - d
    v = 0*x0
    while (res['norm_res'] > par['tol']) and (res['kit'] < par['maxit']):
        res['kit'] += 1
        x_prev = x
        if res['kit'] == 1:
            p = 0
            w = 1/d
        elif res['kit'] == 2:
            p = -(1/2)*(c/d)*(c/d)
            w = 1/(d-c*c/2/d)
        else:
            p = -(c*c/4)*w*w
            w = 1/(d-c*c*w/4)
        v = r - p*v
        x = x_prev + w*v
        r = B - A(x)

        res['norm_res'] = (1.0/r0)*r.norm()

        if callback is not None:
            callback(x)

    if par['tol'] < res['norm_res']:  # if tolerance is less than error norm
        print("Chebyshev solver does not converges!")
    else:
        print("Chebyshev solver converges.")

    if res['kit'] == 0:
        res['norm_res'] = 0
    return x, res


def get_scal(B, par):
    "defines scalar multiplication depending on vectors (general/solver.py:287-298)"
    if 'scal' in par:
        scal = par['scal']
    else:
        if isinstance(B, np.matrix) or _is_vectri(B):
            scal = lambda X, Y: float(X.T*Y)  # noqa: E731
        elif isinstance(B, Tensor):
            scal = lambda X, Y: X*Y  # noqa: E731
        else:
            scal = lambda X, Y: np.sum(X*Y.conj()).real  # noqa: E731
    return scal


def get_norm(B, par):
    scal = get_scal(B, par)
    norm = lambda X: scal(X, X)**0.5  # noqa: E731
    return norm
