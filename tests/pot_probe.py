import os, sys, types, contextlib, io
import numpy as np
HERE = os.path.dirname(os.path.abspath(__file__)); ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'oracle'))
REF = os.path.join(ROOT, 'baseline', '_ref')
os.environ['FFTHOMPY_REFERENCE'] = REF
import _refshim; _refshim.REFERENCE = REF; _refshim.install()
for m in ('tt', 'tt.core', 'tt.core.vector'):
    sys.modules.setdefault(m, types.ModuleType(m))
sys.modules['tt.core.vector'].vector = type('vector', (), {})
from ffthompy import Struct as RStruct
import ffthompy.tensors as RT
import ffthompy.tensorsLowRank.homogenisation as RH
import ffthompy_b200.tensors as MT
import ffthompy_b200.homogenisation as MH
from ffthompy_b200 import device; device.init(0)
g = np.load(os.path.join(HERE, 'golden', 'potential.npz'))
dim, n = 2, 16
tag = 'pot_d%d_n%d' % (dim, n)
N = n*np.ones(dim, dtype=int); Nbar = 2*N-1
res = {}
for name, T, H, S in (('ref', RT, RH, RStruct), ('mine', MT, MH, MH.Struct)):
    Aga = T.Tensor(name='Aga', val=g[tag+'_Aga'].copy(), order=2, N=Nbar, multype=21)
    pars = S(dim=dim, N=N, Y=np.ones(dim), solver=dict(tol=1e-8, maxiter=200))
    F2 = T.DFT(name='FN', inverse=False, N=Nbar); iF2 = T.DFT(name='FiN', inverse=True, N=Nbar)
    P = H.get_preconditioner(N, pars)
    E = np.zeros(dim); E[0] = 1
    EN = T.Tensor(name='EN', N=Nbar, shape=(dim,), Fourier=False); EN.set_mean(E)
    def DF(X):
        FAX = F2(Aga*iF2(T.grad(X).enlarge(Nbar)))
        FAX = FAX.project(N)
        return -T.div(FAX)
    B = T.div(F2(Aga(EN)).decrease(N))
    PB = P*B
    rng = np.random.default_rng(0)
    nh = n//2+1
    xv = rng.standard_normal((n, nh))+1j*rng.standard_normal((n, nh))
    X = T.Tensor(name='x', val=xv.copy(), order=0, N=N, Fourier=True)
    res[name] = dict(P=np.array(P.val), B=np.array(B.val), PB=np.array(PB.val), op=np.array((P*DF(P*X)).val),
                     op_PB=np.array((P*DF(P*PB)).val), dotPB=PB*PB, PX=np.array((P*X).val),
                     g=np.array(T.grad(P*X).val), ge=np.array(T.grad(P*X).enlarge(Nbar).val))
for k in res['ref']:
    a, b = res['ref'][k], res['mine'][k]
    print(k, np.shape(a), np.shape(b), '%.3e' % (np.abs(np.asarray(a)-np.asarray(b)).max()/max(1e-300, np.abs(np.asarray(a)).max())))
