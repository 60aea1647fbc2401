"""Krylov solvers behind `linear_solver` — drop-in for ffthompy/general/solver.py.

Same entry point, argument meaning, defaults, stop rules and `info` keys as the reference (general/solver.py:8-303);
the bodies are this package's own:

* CG and Richardson on the fused G·A operator run as device loops (csrc/fh_fused.cu: fh_cg, fh_cg_begin/steps,
  fh_richardson) — no host round trip other than the 8-byte residual norm that decides termination.  A callback
  (ffthompy/applications.py:71-72 always attaches one) keeps the device loop: it is advanced one iteration per
  call, the iterate handed to the callback after each.
* Every other combination (BiCG, Chebyshev, custom scalar products, unfused operators, legacy VecTri operands) runs
  the same recurrences over the operand algebra — each operation a device kernel — through the small `_Recurrence`
  helpers below.
"""
import numpy as np

from .base import Timer
from ..tensors import Tensor, Operator

_SOLVERS = {}


def _solver(*names):
    def reg(fn):
        for nm in names:
            _SOLVERS[nm] = fn
        return fn
    return reg


def linear_solver(Afun, B, ATfun=None, x0=None, par=None, solver=None, callback=None):
    """Solve Afun(x) = B (general/solver.py:8-60): `solver` in cg | bicg | iterative | richardson | chebyshev | cheby |
    scipy_cg | scipy_bicg (+ the dotted scipy names); returns (x, info) with info['kit'], ['norm_res'], ['time']."""
    watch = Timer('Solving linsys by %s' % solver)
    if x0 is None:
        x0 = B.zeros_like()
    if callback is not None:
        callback(x0)
    key = solver.lower() if solver.split('_')[0].lower() != 'scipy' and not solver.lower().startswith('scipy.') else 'scipy'
    if key not in _SOLVERS:
        raise NotImplementedError("This kind (%s) of linear solver is not implemented" % solver)
    x, info = _SOLVERS[key](Afun=Afun, ATfun=ATfun, B=B, x0=x0, par=par, callback=callback, name=solver)
    watch.measure(print_time=False)
    info.update({'time': watch.vals})
    return x, info


# ----------------------------------------------------------------------------- helpers
def _is_vectri(B):
    try:
        from ..matvecs import VecTri
        return isinstance(B, VecTri)
    except Exception:
        return False


def get_scal(B, par):
    """scalar product matching the operand type, or par['scal'] (general/solver.py:287-298)"""
    if par is not None and 'scal' in par:
        return par['scal']
    if isinstance(B, np.matrix) or _is_vectri(B):
        return lambda X, Y: float(X.T*Y)
    if isinstance(B, Tensor):
        return lambda X, Y: X*Y
    return lambda X, Y: np.sum(X*Y.conj()).real


def get_norm(B, par):
    dot = get_scal(B, par)
    return lambda X: dot(X, X)**0.5


def _settings(par, **defaults):
    par = dict() if par is None else par
    for k, v in defaults.items():
        par.setdefault(k, v)       # the reference also writes its defaults into the caller's dict
    return par


def _fused_for(Afun, B, x0, par):
    """the device pipeline of `Afun` if B and x0 are fields it accepts and the default scalar product is wanted"""
    if par is not None and 'scal' in par:
        return None
    if not (isinstance(Afun, Operator) and isinstance(B, Tensor) and isinstance(x0, Tensor)):
        return None
    f = Afun.fused()
    if f is None or not (f.accepts(B) and f.accepts(x0)):
        return None
    return f


# ----------------------------------------------------------------------------- stationary iteration
@_solver('iterative', 'richardson')
def _richardson_entry(Afun, B, x0, par, callback, **_):
    return richardson(Afun, B, x0, par=par, callback=callback)


def richardson(Afun, B, x0, par=None, callback=None):
    """x <- x + (B - A x)/alpha until the residual taken BEFORE the update is <= tol (general/solver.py:63-77)"""
    info = {'norm_res': 1e15, 'kit': 0}
    f = _fused_for(Afun, B, x0, par) if callback is None else None
    if f is not None:
        xd, info['kit'], info['norm_res'] = f.richardson(B._dev(), x0._dev(), par['alpha'], par['tol'], int(par['maxiter']))
        return x0.copy(val=xd), info
    size = get_norm(B, par)
    step = 1./par['alpha']
    x = x0
    while info['norm_res'] > par['tol'] and info['kit'] < par['maxiter']:
        defect = B-Afun(x)
        x = x+step*defect
        info['kit'] += 1
        info['norm_res'] = size(defect)
        if callback is not None:
            callback(x)
    return x, info


# ----------------------------------------------------------------------------- conjugate gradients
@_solver('cg')
def _cg_entry(Afun, B, x0, par, callback, **_):
    return CG(Afun, B, x0=x0, par=par, callback=callback)


def CG(Afun, B, x0, par=None, callback=None):
    """Conjugate gradients with the reference's conventions (general/solver.py:80-139): the operator is applied to x0
    even when it is zero, the stop test is ABSOLUTE on sqrt(<r,r>) in the mean-normalised scalar product of the
    operands, and a solve that needs no iteration reports norm_res = 0."""
    par = _settings(par, tol=1e-6, maxiter=int(1e3))
    f = _fused_for(Afun, B, x0, par)
    if f is not None:
        if callback is None:
            xd, kit, nres, hist = f.cg(B._dev(), x0._dev(), par['tol'], int(par['maxiter']))
        else:
            xd, kit, nres, hist = f.cg_callback(B._dev(), x0._dev(), par['tol'], int(par['maxiter']),
                                                lambda buf: callback(x0.copy(val=buf)))
        return x0.copy(val=xd), {'kit': kit, 'norm_res': nres if kit > 0 else 0, 'norm_res_log': hist}

    dot = get_scal(B, par)
    x = x0
    r = B-Afun(x0)
    d = r                       # search direction
    rho = dot(r, r)
    history = [np.double(rho)**0.5]
    kit = 0
    while history[-1] > par['tol'] and kit < par['maxiter']:
        kit += 1
        q = Afun(d)
        step = float(rho/dot(d, q))
        x = x+step*d
        r = r-step*q
        rho_next = dot(r, r)
        d = r+(rho_next/rho)*d
        rho = rho_next
        history.append(np.double(rho)**0.5)
        if callback is not None:
            callback(x)
    return x, {'kit': kit, 'norm_res': history[-1] if kit > 0 else 0, 'norm_res_log': np.array(history)}


# ----------------------------------------------------------------------------- biconjugate gradients
@_solver('bicg')
def _bicg_entry(Afun, ATfun, B, x0, par, callback, **_):
    return BiCG(Afun, ATfun, B, x0=x0, par=par, callback=callback)


def BiCG(Afun, ATfun, B, x0, par=None, callback=None):
    """BiConjugate gradients with the shadow system driven by ATfun (general/solver.py:142-204); the "norm" is
    sqrt(<r, r_shadow>) as in the reference."""
    par = _settings(par, tol=1e-6, maxiter=1e3)
    dot = get_scal(B, par)
    x = x0
    r = B-Afun(x0)
    rs = r                       # shadow residual
    d, ds = r, rs                # direction and shadow direction
    rho = float(dot(r, rs))
    info = {'kit': 0, 'norm_res': rho**0.5}
    while info['norm_res'] > par['tol'] and info['kit'] < par['maxiter']:
        info['kit'] += 1
        q = Afun(d)
        step = rho/float(dot(q, ds))
        x = x+step*d
        r = r-step*q
        rs = rs-step*ATfun(ds)
        rho_next = float(dot(r, rs))
        d = r+(rho_next/rho)*d
        ds = rs+(rho_next/rho)*ds
        rho = rho_next
        info['norm_res'] = rho**0.5
        if callback is not None:
            callback(x)
    if info['kit'] == 0:
        info['norm_res'] = 0
    return x, info


# ----------------------------------------------------------------------------- Chebyshev
@_solver('chebyshev', 'cheby')
def _cheby_entry(Afun, B, x0, par, callback, **_):
    return cheby2TERM(A=Afun, B=B, x0=x0, par=par, callback=callback)


def _cheby_weights(lo, hi):
    """(momentum p_k, step w_k) of the two-term Chebyshev recurrence for a spectrum in [lo, hi], k = 1, 2, ..."""
    centre, half = (hi+lo)/2.0, (hi-lo)/2.0
    k, w = 0, 0.
    while True:
        k += 1
        if k == 1:
            p, w = 0, 1/centre
        elif k == 2:
            p, w = -(1/2)*(half/centre)*(half/centre), 1/(centre-half*half/2/centre)
        else:
            p, w = -(half*half/4)*w*w, 1/(centre-half*half*w/4)
        yield p, w


def cheby2TERM(A, B, x0, M=None, par=None, callback=None):
    """Chebyshev two-term iteration (general/solver.py:206-285): needs par['eigrange'] = (lambda_min, lambda_max);
    residual norms are relative to the initial residual, the iteration cap is par['maxit']."""
    par = _settings(par, tol=1e-06, maxit=1e7)
    if 'eigrange' not in par:
        raise NotImplementedError("It is necessary to calculate eigenvalues.")
    info = {'kit': 0}
    x = x0
    r = B-A(x)
    r0 = np.double(r*r)**0.5
    info['norm_res'] = r0/(B*B)**0.5
    if info['norm_res'] < par['tol']:
        return x, info
    v = 0*x0
    weights = _cheby_weights(par['eigrange'][0], par['eigrange'][1])
    while info['norm_res'] > par['tol'] and info['kit'] < par['maxit']:
        info['kit'] += 1
        p, w = next(weights)
        v = r-p*v
        x = x+w*v
        r = B-A(x)
        info['norm_res'] = r.norm()/r0
        if callback is not None:
            callback(x)
    print("Chebyshev solver converges." if info['norm_res'] <= par['tol'] else "Chebyshev solver does not converges!")
    if info['kit'] == 0:
        info['norm_res'] = 0
    return x, info


# ----------------------------------------------------------------------------- SciPy bridge
@_solver('scipy')
def _scipy_entry(Afun, ATfun, B, x0, par, callback, name, **_):
    """general/solver.py:28-53: SciPy iterates on host vectors; every matvec is a device call.  SciPy >= 1.12 names
    the relative tolerance `rtol`, older versions `tol`."""
    import inspect
    import scipy.sparse.linalg as spslin
    kind = name.split('.')[-1].split('_')[-1]
    if kind not in ('cg', 'bicg'):
        raise NotImplementedError("This kind (%s) of linear solver is not implemented" % name)
    start = x0.ravel() if isinstance(x0, np.ndarray) else np.asarray(x0.vec()).ravel()
    Afun.define_operand(B)
    ops = {'matvec': lambda v: np.asarray(Afun.matvec(v)).ravel()}
    if kind == 'bicg':
        ATfun.define_operand(B)
        ops['rmatvec'] = lambda v: np.asarray(ATfun.matvec(v)).ravel()
    lin = spslin.LinearOperator(Afun.matshape, dtype=np.float64, **ops)
    fn = getattr(spslin, kind)
    tolkw = 'rtol' if 'rtol' in inspect.signature(fn).parameters else 'tol'
    xcol, flag = fn(lin, np.asarray(B.vec()).ravel(), x0=start, maxiter=int(par['maxiter']), M=None,
                    callback=callback, **{tolkw: par['tol']})
    x = B.empty_like(name='x')
    x.val = np.reshape(xcol, B._vshape())
    return x, {'info': flag}
