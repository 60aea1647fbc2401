"""Recognition of the solve-loop operator and its fused device pipeline.

The reference builds (applications.py:33-58, tutorials/02_homogenisation.py:170-236)
    GN   = Operator([[FiN, G^, FN]])         projection: inverse DFT · multiplier · DFT
    Afun = Operator([[GN, A]])               y = GN(A(x))
and iterates it in linear_solver.  When `G^` is a closed-form GreenTensor, `A` a real (D, D)
coefficient tensor and the transforms are the 'r' form on the same grid, the whole product
runs as the fixed kernel pipeline of csrc/fh_fused.cu (fh_ga_*), and CG / Richardson run as
device loops (fh_cg / fh_richardson).
"""
import ctypes as C

import numpy as np

from . import _lib as L
from . import device as dev


class FusedGA(object):
    def __init__(self, A, G, N):
        self.N = tuple(int(n) for n in N)
        self.D = int(A.shape[0])
        self.A_dev = A._dev()           # keeps the coefficient buffer alive
        # the operator snapshots A (phase table / symmetry flag, fh_ga_create): key the cache on the buffer
        # identity AND on the Tensor's in-place mutation counter (add_mean / set_mean write into the same buffer)
        self.A_id = (id(self.A_dev), getattr(A, '_version', 0))
        self.green_key = _green_key(G)
        self.plan = dev.plan(self.N)
        lib = dev.lib()
        nwork = int(lib.fh_ga_work_doubles(self.plan, self.D))
        self.work = dev.empty((nwork,))
        self.handle = C.c_void_p()
        g = G.descriptor()
        L.check(lib.fh_ga_create(C.byref(self.handle), self.plan, self.D, dev.ptr(self.A_dev), 0, C.byref(g),
                                 dev.ptr(self.work)))
        self.nreal = int(np.prod(self.N))

    def __del__(self):
        try:
            if getattr(self, 'handle', None):
                L.load().fh_ga_destroy(self.handle)
        except Exception:
            pass

    def config(self):
        """which kernel family serves each axis and how S1 reads the coefficients (fh_ga_config)"""
        fl, pitch, mt = C.c_int(), C.c_int(), C.c_int()
        L.check(dev.lib().fh_ga_config(self.handle, C.byref(fl), C.byref(pitch), C.byref(mt)))
        f = fl.value
        # 'odd': compile-time two-pass kernels of csrc/fh_odd.cu (255 = 15 x 17); 'rt': run-time-length in-place kernels
        fam = lambda fast, rt, odd: ('odd' if f & odd else 'pow2' if f & fast else  # noqa: E731
                                     ('rt' if f & rt else 'generic'))
        return {'last': fam(1, 1 << 16, 1 << 20), 'mid1': fam(2, 1 << 17, 1 << 21), 'mid0': fam(4, 1 << 18, 1 << 22),
                'coefficients': ('full', 'symmetric', 'phase')[(f >> 4) & 3], 'nphase': (f >> 8) & 0xff,
                'pitch': pitch.value}

    def accepts(self, x):
        return ((not x.Fourier) and x.order == 1 and tuple(x.shape) == (self.D,) and tuple(x.N) == self.N
                and not x._is_complex())

    def apply(self, x_dev):
        y = dev.empty((self.D,)+self.N)
        L.check(dev.lib().fh_ga_apply(self.handle, dev.ptr(x_dev), dev.ptr(y)))
        return y

    def cg(self, B_dev, x0_dev, tol, maxiter):
        """general/solver.py:80-139 as a device loop.  Returns x (device), kit, norm_res, history."""
        from . import ops
        x = ops.clone(x0_dev)
        vecs = dev.empty((3*self.D*self.nreal,))
        kit = C.c_int64()
        nres = C.c_double()
        cap = int(min(max(maxiter, 0), 100000))+1
        hist = (C.c_double*cap)()
        L.check(dev.lib().fh_cg(self.handle, dev.ptr(B_dev), dev.ptr(x), float(tol), int(maxiter), dev.ptr(vecs),
                                C.byref(kit), C.byref(nres), hist, cap))
        k = int(kit.value)
        return x, k, float(nres.value), np.array(hist[:min(k+1, cap)])

    def cg_callback(self, B_dev, x0_dev, tol, maxiter, on_iterate):
        """the same device loop advanced one iteration per call (fh_cg_begin / fh_cg_steps), `on_iterate(x)` after
        each with a private copy of the iterate (general/solver.py:134-135: callback(xCG)).  The operator stays
        fused; only the deferred x update is flushed every iteration instead of once at the end."""
        from . import ops
        lib = dev.lib()
        x = ops.clone(x0_dev)
        vecs = dev.empty((3*self.D*self.nreal,))
        nres, done = C.c_double(), C.c_int64()
        L.check(lib.fh_cg_begin(self.handle, dev.ptr(B_dev), dev.ptr(x), dev.ptr(vecs), C.byref(nres)))
        hist = [float(nres.value)]
        kit = 0
        while hist[-1] > tol and kit < maxiter:
            L.check(lib.fh_cg_steps(self.handle, dev.ptr(x), dev.ptr(vecs), float(tol), 1, C.byref(done),
                                    C.byref(nres), None))
            if done.value != 1:
                break
            kit += 1
            hist.append(float(nres.value))
            on_iterate(ops.clone(x))
        return x, kit, hist[-1], np.array(hist)

    def richardson(self, B_dev, x0_dev, alpha, tol, maxiter):
        """general/solver.py:63-77 as a device loop."""
        from . import ops
        x = ops.clone(x0_dev)
        vecs = dev.empty((2*self.D*self.nreal,))
        kit = C.c_int64()
        nres = C.c_double()
        L.check(dev.lib().fh_richardson(self.handle, dev.ptr(B_dev), dev.ptr(x), float(alpha), float(tol),
                                        int(maxiter), dev.ptr(vecs), C.byref(kit), C.byref(nres)))
        return x, int(kit.value), float(nres.value)


def _green_key(G):
    g = G.green
    return (g['kind'], tuple(g['band']), tuple(sorted(g['coef'].items())), tuple(G.N), tuple(G.Y), G.fft_form)


def match(op):
    """(A, G, N) if `op` is Operator([[Operator([[FiN, G, FN]]), A]]) in fusable form, else None."""
    from .tensors.objects import Tensor
    from .tensors.operators import DFT, Operator
    from .projections import GreenTensor
    try:
        if len(op.mat_rev) != 1 or len(op.mat_rev[0]) != 2:
            return None
        A, GN = op.mat_rev[0]
        if not (isinstance(A, Tensor) and isinstance(GN, Operator)):
            return None
        if len(GN.mat_rev) != 1 or len(GN.mat_rev[0]) != 3:
            return None
        FN, G, FiN = GN.mat_rev[0]
        if not (isinstance(FN, DFT) and isinstance(FiN, DFT) and isinstance(G, GreenTensor)):
            return None
        if FN.inverse or not FiN.inverse or FN.fft_form != 'r' or FiN.fft_form != 'r':
            return None
        if not G.lazy or G.fft_form != 'r' or G.multype not in (21, '21'):
            return None
        N = tuple(int(n) for n in G.N)
        if tuple(int(n) for n in FN.N) != N or tuple(int(n) for n in FiN.N) != N or len(N) not in (2, 3):
            return None
        if isinstance(A, GreenTensor) or A.Fourier or A.order != 2 or A.multype not in (21, '21'):
            return None
        if A._is_complex() or tuple(A.N) != N or A.shape[0] != A.shape[1] or A.shape[0] != G.shape[0]:
            return None
        if A.shape[0] not in (2, 3, 6):
            return None
        return A, G, N
    except AttributeError:
        return None


def get_fused(op, cached):
    m = match(op)
    if m is None:
        return None
    A, G, N = m
    if cached is not None and cached.A_id == (id(A._dev()), getattr(A, '_version', 0)) and cached.green_key == _green_key(G):
        return cached
    return FusedGA(A, G, N)
