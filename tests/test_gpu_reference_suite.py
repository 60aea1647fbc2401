"""The reference's own unit-test properties, exercised on the B200 objects (same API calls, same
tolerances): ffthompy/tensors/unittest_operators.py, tensors/unittest_tensors.py,
general/unittest_solver.py and matvecs/unittest_matvec.py."""
import itertools
import warnings

import numpy as np
import pytest
from numpy.linalg import norm

pytestmark = pytest.mark.gpu
fft_forms = [0, 'r', 'c']


@pytest.mark.parametrize('dim,fft_form', list(itertools.product([2, 3], fft_forms)))
def test_operators(dim, fft_form):
    """unittest_operators.py:25-98"""
    from ffthompy_b200.tensors import Tensor, DFT, grad, div, symgrad, potential, grad_div_tensor
    N = 5*np.ones(dim, dtype=int)
    F = DFT(N=N, inverse=False, fft_form=fft_form)
    iF = DFT(N=N, inverse=True, fft_form=fft_form)
    assert 'DFT' in repr(F)
    u = Tensor(name='u', shape=(), N=N, Fourier=False, fft_form=fft_form).randomize()
    Fu = F(u)
    u2 = iF(Fu)
    assert (u == u2)[1] < 1e-13, 'Fourier transform'
    for fft_formc in [f for f in fft_forms if f != fft_form]:
        FuC = Fu.set_fft_form(fft_formc, copy=True)
        Fu2 = FuC.set_fft_form(fft_form, copy=True)
        assert abs(Fu.norm()-FuC.norm()) < 1e-13
        assert norm(Fu.mean()-FuC.mean()) < 1e-13
        assert (Fu == Fu2)[1] < 1e-13
    # scalar problem
    u = Tensor(name='u', shape=(1,), N=N, Fourier=False, fft_form=fft_form).randomize()
    u.val -= np.mean(u.val)
    Fu = F(u)
    Fu2 = potential(grad(Fu))
    assert (Fu == Fu2)[1] < 1e-13
    u2 = potential(grad(u))
    assert (u == u2)[1] < 1e-13
    hG, hD = grad_div_tensor(N, fft_form=fft_form)
    assert (hD(hG(Fu)) == div(grad(Fu)))[1] < 1e-13
    # vectorial problem
    u = Tensor(name='u', shape=(dim,), N=N, Fourier=False, fft_form=fft_form)
    u.randomize()
    u.add_mean(-u.mean())
    Fu = F(u)
    Fu2 = potential(grad(Fu))
    assert (Fu == Fu2)[1] < 1e-13
    u2 = potential(grad(u))
    assert (u == u2)[1] < 1e-13
    # vectorial problem - symmetric gradient
    Fu2 = potential(symgrad(Fu), small_strain=True)
    assert (Fu == Fu2)[1] < 1e-13
    u2 = potential(symgrad(u), small_strain=True)
    assert (u == u2)[1] < 1e-13
    # matrix version of DFT
    u = Tensor(name='u', shape=(1,), N=N, Fourier=False, fft_form='c').randomize()
    F = DFT(N=N, inverse=False, fft_form='c')
    Fu = F(u)
    dft = F.matrix(shape=u.shape)
    Fu2 = np.asarray(dft.dot(u.val.ravel())).ravel()
    assert norm(Fu.val.ravel()-Fu2) < 1e-13


@pytest.mark.parametrize('fft_form', fft_forms)
def test_compatibility(fft_form):
    """unittest_operators.py:100-173"""
    from ffthompy_b200.tensors import Tensor, DFT, grad, symgrad, potential, Operator
    from ffthompy_b200.tensors.projection import scalar, elasticity_small_strain, elasticity_large_deformation
    dim = 3
    N = 5*np.ones(dim, dtype=int)
    F = DFT(inverse=False, N=N, fft_form=fft_form)
    iF = DFT(inverse=True, N=N, fft_form=fft_form)
    # scalar problem
    _, G1l, G2l = scalar(N, Y=np.ones(dim), fft_form=fft_form)
    P1 = Operator(name='P1', mat=[[iF, G1l, F]])
    P2 = Operator(name='P2', mat=[[iF, G2l, F]])
    u = Tensor(name='u', shape=(1,), N=N, Fourier=False, fft_form=fft_form)
    u.randomize()
    grad_u = grad(u)
    assert (P1(grad_u)-grad_u).norm() < 1e-13
    assert P2(grad_u).norm() < 1e-13
    e = P1(Tensor(name='u', shape=(dim,), N=N, Fourier=False, fft_form=fft_form).randomize())
    e2 = grad(potential(e))
    assert (e-e2).norm() < 1e-13
    # vectorial problem
    hG = elasticity_large_deformation(N=N, Y=np.ones(dim), fft_form=fft_form)
    P1 = Operator(name='P', mat=[[iF, hG, F]])
    u = Tensor(name='u', shape=(dim,), N=N, Fourier=False, fft_form=fft_form)
    u.randomize()
    grad_u = grad(u)
    assert (P1(grad_u)-grad_u).norm() < 1e-13
    e = Tensor(name='F', shape=(dim, dim), N=N, Fourier=False, fft_form=fft_form)
    e = P1(e.randomize())
    e2 = grad(potential(e))
    assert (e-e2).norm() < 1e-13
    # transpose
    P1TT = P1.transpose().transpose()
    assert (P1(grad_u) == P1TT(grad_u))[0]
    assert (hG == (hG.transpose_left().transpose_left()))[0]
    assert (hG == (hG.transpose_right().transpose_right()))[0]
    # vectorial problem - symmetric gradient
    hG = elasticity_small_strain(N=N, Y=np.ones(dim), fft_form=fft_form)
    P1 = Operator(name='P', mat=[[iF, hG, F]])
    u = Tensor(name='u', shape=(dim,), N=N, Fourier=False, fft_form=fft_form)
    u.randomize()
    grad_u = symgrad(u)
    assert (P1(grad_u)-grad_u).norm() < 1e-13
    e = Tensor(name='strain', shape=(dim, dim), N=N, Fourier=False, fft_form=fft_form)
    e = P1(e.randomize())
    e2 = symgrad(potential(e, small_strain=True))
    assert (e-e2).norm() < 1e-13
    # means
    Fu = F(u)
    E = np.random.random(u.shape)
    u.set_mean(E)
    assert norm(u.mean()-E) < 1e-13
    Fu.set_mean(E)
    assert norm(Fu.mean()-E) < 1e-13
    assert 'Operator' in repr(P1) and 'Tensor' in repr(u)


@pytest.mark.parametrize('dim,fft_form', list(itertools.product([2, 3], fft_forms)))
def test_projections(dim, fft_form):
    """unittest_operators.py:175-203: idempotency and mutual orthogonality of all projection pieces"""
    import ffthompy_b200.projections as proj
    N = dim*(5,)
    Y = np.ones(dim)
    for projections in (proj.scalar(N, Y, tensor=True, fft_form=fft_form),
                        proj.elasticity(N, Y, tensor=True, fft_form=fft_form)):
        for P, Q in itertools.product(projections, repeat=2):
            if (P == Q)[0]:
                assert (P*P-P).norm() < 1e-13  # idempotent
            else:
                assert (P*Q).norm() < 1e-13  # orthogonality


@pytest.mark.parametrize('dim,n,fft_form', list(itertools.product([2, 3], [4, 5], ['r', 0, 'c'])))
def test_tensors_even_odd(dim, n, fft_form):
    """unittest_tensors.py:11-47: project(2N) preserves mean/norm and the values on coincident nodes"""
    from ffthompy_b200.tensors import Tensor
    N = dim*(n,)
    M = tuple(2*np.array(N))
    u = Tensor(name='test', shape=(), N=N, Fourier=False, fft_form=fft_form)
    u.randomize()
    Fu = u.fourier(copy=True)
    FuM = Fu.project(M)
    uM = FuM.fourier()
    if n % 2 == 0:
        assert u.norm() >= FuM.norm()-1e-14
        assert np.all(u.norm(componentwise=True) >= FuM.norm(componentwise=True)-1e-14)
        assert u.norm() >= uM.norm()-1e-14
    else:
        assert abs(u.norm()-FuM.norm()) < 1e-7
        assert np.all(np.abs(u.norm(componentwise=True)-FuM.norm(componentwise=True)) < 1e-7)
        assert abs(u.norm()-uM.norm()) < 1e-7
    assert abs(u.mean()-FuM.mean()) < 1e-7
    assert abs(u.mean()-uM.mean()) < 1e-7
    slc = tuple(u.order*[slice(None), ]+[slice(0, M[i], 2) for i in range(dim)])
    assert np.linalg.norm(u.val-uM.val[slc]) < 1e-7
    assert np.linalg.norm(u.val-uM.val[slc]) < 1e-13  # in fact exact to rounding


def test_solver_projections_agree():
    """unittest_solver.py:21-33"""
    from ffthompy_b200.projections import scalar
    from ffthompy_b200.tensors.projection import scalar as scalar_tensor
    N = 5*np.ones(2, dtype=int)
    hG0N, hG1N, hG2N = scalar(N, Y=np.ones(2))
    hG0Nt, hG1Nt, hG2Nt = scalar_tensor(N, Y=np.ones(2))
    assert norm(hG0N.val-hG0Nt.val) < 1e-13
    assert norm(hG1N.val-hG1Nt.val) < 1e-13
    assert norm(hG2N.val-hG2Nt.val) < 1e-13


def test_solvers_cross_agreement():
    """unittest_solver.py:35-75: CG, scipy_cg, richardson, chebyshev agree to 1e-8"""
    from ffthompy_b200.tensors import Tensor, DFT, Operator
    from ffthompy_b200.tensors.projection import scalar as scalar_tensor
    from ffthompy_b200.general.solver import linear_solver
    dim, n = 2, 5
    N = n*np.ones(dim, dtype=int)
    _, hG1Nt, _ = scalar_tensor(N, Y=np.ones(dim))
    FN = DFT(name='FN', inverse=False, N=N)
    FiN = DFT(name='FiN', inverse=True, N=N)
    G1N = Operator(name='G1', mat=[[FiN, hG1Nt, FN]])
    A = Tensor(name='A', val=np.einsum('ij,...->ij...', np.eye(dim), 1.+10.*np.random.random(tuple(N))),
               order=2, N=N, multype=21)
    E = np.zeros((dim,)+dim*(n,))
    E[0] = 1.  # set macroscopic loading
    E = Tensor(name='E', val=E, order=1, N=N)
    GAfun = Operator(name='GA', mat=[[G1N, A]])
    GAfun.define_operand(E)
    B = GAfun(-E)
    x0 = E.copy(name='x0')
    x0.val[:] = 0
    par = {'tol': 1e-10, 'maxiter': int(1e3), 'alpha': 0.5*(1.+10.), 'eigrange': [1., 10.]}
    X, _ = linear_solver(Afun=GAfun, B=B, x0=x0, par=par, solver='CG')
    for solver in ['CG', 'scipy_cg', 'richardson', 'chebyshev']:
        x, _ = linear_solver(Afun=GAfun, B=B, x0=x0, par=par, solver=solver)
        assert norm(X.val-x.val) < 1e-8, solver
    with pytest.raises(NotImplementedError):
        linear_solver(Afun=GAfun, B=B, x0=x0, par=par, solver='gmres')


def test_legacy_matvec():
    """matvecs/unittest_matvec.py:14-40 and the legacy solve of SURVEY D.10"""
    from ffthompy_b200.matvecs import DFT, VecTri, Matrix, LinOper
    from ffthompy_b200.general.solver import linear_solver
    import ffthom_oracle as O
    with warnings.catch_warnings():
        warnings.simplefilter('ignore')
        for dim in [2, 3]:
            for n in [4, 5]:
                N = n*np.ones(dim, dtype=int)
                ur = VecTri(name='rand', dim=2, N=N, valtype='rand')
                FN = DFT(name='FN', inverse=False, N=N, d=dim)
                FiN = DFT(name='FiN', inverse=True, N=N, d=dim)
                Fur = FN(ur)
                assert np.linalg.norm(Fur.vec()-FN.matrix().dot(ur.vec())) < 1e-13
                assert np.linalg.norm(ur.vec()-FiN.matrix().dot(Fur.vec())) < 1e-13
                assert np.abs(Fur.val-O.cfftnc(ur.val, tuple(N))).max() < 1e-14
        for dim in [2, 3]:
            N = 5*np.ones(dim, dtype=int)
            uN = VecTri(name='rand', dim=dim, N=N, valtype='rand')
            for i in range(2):
                assert (uN == uN.project(2*N-i).project(N)) < 1e-13
        # legacy solve loop: centred projection Matrix x centred DFT, odd N
        N = np.array([5, 5])
        _, G1, _ = O.proj_scalar(N, np.ones(2), fft_form='c')
        hG1 = Matrix(name='hG1', val=G1.astype(complex), Fourier=True)
        rng = np.random.default_rng(2)
        Aval = np.einsum('ij,...->ij...', np.eye(2), 1+10*(rng.random((5, 5)) < 0.4))
        A = Matrix(name='A', val=Aval, Fourier=False)
        FN, FiN = DFT(name='FN', inverse=False, N=N), DFT(name='FiN', inverse=True, N=N)
        GA = LinOper(name='GA', mat=[[FiN, hG1, FN, A]])
        E = VecTri(macroval=np.array([1., 0.]), N=N)
        B = GA(-E)
        X, info = linear_solver(solver='CG', Afun=GA, B=B, x0=VecTri(N=N), par={'tol': 1e-8, 'maxiter': 100})
        # same problem in the centred oracle
        Afun = lambda x: O.icfftnc(np.einsum('ij...,j...->i...', G1, O.cfftnc(np.einsum('ij...,j...->i...', Aval, x), N)), N)  # noqa: E731
        Ev = np.zeros((2, 5, 5))
        Ev[0] = 1
        xo, io = O.cg(Afun, Afun(-Ev), np.zeros_like(Ev), 1e-8, 100, N)
        assert info['kit'] == io['kit']
        assert np.abs(X.val-xo).max() < 1e-10


def test_block_vectors_and_operators():
    """MultiVector / MultiOper / ScipyOper (matvecs/objects.py:938-1165): block algebra against NumPy on the host values"""
    from ffthompy_b200.matvecs import VecTri, Matrix, MultiVector, MultiOper, ScipyOper, Scalar
    with warnings.catch_warnings():
        warnings.simplefilter('ignore')
        N = np.array([5, 4])
        rng = np.random.default_rng(7)
        u = [VecTri(name='u%d' % i, val=rng.standard_normal((2, 5, 4))) for i in range(2)]
        v = [VecTri(name='v%d' % i, val=rng.standard_normal((2, 5, 4))) for i in range(2)]
        U, V = MultiVector(val=u), MultiVector(val=v)
        assert U.dim == 2 and U.size == 2*2*5*4 and U.ltype == ['VecTri', 'VecTri']
        # scalar product = sum of the block scalar products (each 1/prod(N)-weighted, as VecTri*VecTri)
        ref = sum(np.sum(a.val*b.val) for a, b in zip(u, v))/np.prod(N)
        assert abs(U*V-ref) < 1e-13
        W = U+V
        assert all(np.abs(W[m].val-(u[m].val+v[m].val)).max() < 1e-15 for m in range(2))
        W = U-V
        assert all(np.abs(W[m].val-(u[m].val-v[m].val)).max() < 1e-15 for m in range(2))
        W = Scalar(val=2.5)*U
        assert all(np.abs(W[m].val-2.5*u[m].val).max() < 1e-15 for m in range(2))
        assert np.abs(np.asarray(U.vec()).ravel()-np.hstack([a.val.ravel() for a in u])).max() == 0
        # block operator of pointwise matrices
        Mv = [[rng.standard_normal((2, 2, 5, 4)) for _ in range(2)] for _ in range(2)]
        Op = MultiOper(val=[[Matrix(name='M%d%d' % (m, n), val=Mv[m][n]) for n in range(2)] for m in range(2)])
        Y = Op(U)
        for m in range(2):
            ref = sum(np.einsum('ij...,j...->i...', Mv[m][n], u[n].val) for n in range(2))
            assert np.abs(Y[m].val-ref).max() < 1e-13
        Yt = Op.transpose()(U)
        for m in range(2):
            ref = sum(np.einsum('ji...,j...->i...', Mv[n][m], u[n].val) for n in range(2))
            assert np.abs(Yt[m].val-ref).max() < 1e-13
        # flat-vector bridge
        S = ScipyOper(A=Op, X=U, AT=Op.transpose())
        assert S.shape == (U.size, U.size)
        x = rng.standard_normal(U.size)
        y = np.asarray(S.matvec(x)).ravel()
        yref = np.asarray(Op(S.revec(x)).vec()).ravel()
        assert np.abs(y-yref).max() < 1e-14
        # <A x, z> == <x, A^T z>
        z = rng.standard_normal(U.size)
        assert abs(np.dot(y, z)-np.dot(x, np.asarray(S.rmatvec(z)).ravel())) < 1e-11
