"""development probe: Fourier-space ops on NOT Hermitian-consistent 'r' spectra of even grids, package vs the
unmodified reference (baseline/_ref)"""
import os, sys
import numpy as np
HERE = os.path.dirname(os.path.abspath(__file__)); ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'oracle'))
REF = os.path.join(ROOT, 'baseline', '_ref')
os.environ['FFTHOMPY_REFERENCE'] = REF
import _refshim; _refshim.REFERENCE = REF; _refshim.install()
import ffthompy.tensors as RT
from ffthompy.tensors import Tensor as RTensor, DFT as RDFT
import ffthompy_b200.tensors as MT
from ffthompy_b200.tensors import Tensor as MTensor, DFT as MDFT
from ffthompy_b200 import device; device.init(0)
rng = np.random.default_rng(0)
for N in [(16, 16), (4, 6), (6, 4), (5, 4), (4, 5), (4, 4, 6), (6, 5, 4)]:
    N = np.array(N); dim = len(N); Nbar = 2*N-1
    nh = N[-1]//2+1
    def mk(shape, M=N):
        mh = M[-1]//2+1
        return rng.standard_normal(shape+tuple(M[:-1])+(mh,))+1j*rng.standard_normal(shape+tuple(M[:-1])+(mh,))
    s0 = mk(()); s1 = mk((dim,)); b1 = mk((dim,), Nbar)
    def both(shape_val, order, M, fn):
        r = fn(RTensor(name='x', val=shape_val.copy(), order=order, N=np.array(M), Fourier=True, fft_form='r'), RT)
        m = fn(MTensor(name='x', val=shape_val.copy(), order=order, N=np.array(M), Fourier=True, fft_form='r'), MT)
        rv = r.val if hasattr(r, 'val') else np.array(r)
        mv = m.val if hasattr(m, 'val') else np.array(m)
        return np.abs(rv-mv).max()/max(1e-300, np.abs(rv).max())
    out = []
    out.append(('grad', both(s0, 0, N, lambda X, T: T.grad(X))))
    out.append(('div', both(s1, 1, N, lambda X, T: T.div(X))))
    out.append(('enlarge0', both(s0, 0, N, lambda X, T: X.enlarge(Nbar))))
    out.append(('enlarge1', both(s1, 1, N, lambda X, T: X.enlarge(Nbar))))
    out.append(('decrease1', both(b1, 1, Nbar, lambda X, T: X.decrease(N))))
    out.append(('project1', both(b1, 1, Nbar, lambda X, T: X.project(N))))
    out.append(('ifft', both(b1, 1, Nbar, lambda X, T: T.DFT(inverse=True, N=Nbar)(X))))
    out.append(('ifftN', both(s1, 1, N, lambda X, T: T.DFT(inverse=True, N=N)(X))))
    out.append(('dot', both(s1, 1, N, lambda X, T: np.array([X*X]))))
    out.append(('norm', both(s0, 0, N, lambda X, T: np.array([X.norm()]))))
    out.append(('mean', both(s1, 1, N, lambda X, T: np.array(X.mean()))))
    out.append(('gradenl', both(s0, 0, N, lambda X, T: T.grad(X).enlarge(Nbar))))
    print(tuple(N), ' '.join('%s %.1e' % kv for kv in out), flush=True)
