"""GPU parity of the compile-time odd-length kernel family (csrc/fh_odd.cu, 255 = 15 x 17: the exact-integration grid of
BASELINE config 2): operator application and complete device CG against the CPU oracle, against the run-time-length
family it replaces (FH_ODD=0), on grids whose row count is odd (partial last CTA of S1 / S5) and with 255 on two and on
all three axes.  The single-axis cases (255 on each axis in turn, three coefficient modes, both physics) live in
tests/test_gpu_bench_sizes.py::test_benchmarked_axis_lengths_against_the_oracle.

Tolerances (fp64): operator 1e-12 relative (max norm) against the oracle, 1e-13 between the two device families, CG
iteration counts EQUAL."""
import numpy as np
import pytest

import ffthom_oracle as O
import harness

pytestmark = pytest.mark.gpu


@pytest.fixture(scope='module', autouse=True)
def _device():
    from ffthompy_b200 import device
    device.init(0)
    before = device.launch_count()
    yield
    assert device.launch_count() > before, 'no kernel of libffthom_b200.so was launched'


def _coefficients(physics, N, mode, rng):
    d = len(N)
    if physics == 'elasticity':
        D = d*(d+1)//2
        Cm, Ci = O.elastic_mandel(1, 1), O.elastic_mandel(10, 5)
    else:
        D = d
        Cm, Ci = np.eye(d), 11.*np.eye(d)
    ph = rng.random(N) < 0.3
    A = np.einsum('ij,...->ij...', Cm, 1.-ph)+np.einsum('ij,...->ij...', Ci, 1.*ph)
    if mode != 'phase':
        M = 0.05*rng.standard_normal((D, D)+N)
        A = A+np.einsum('ik...,jk...->ij...', M, M)
        A = 0.5*(A+np.einsum('ij...->ji...', A))
    return D, A


def _oracle_green(physics, N):
    d = len(N)
    if physics == 'elasticity':
        Go = O.proj_elasticity(N, np.ones(d))
        return Go[1]+Go[2]
    return O.proj_scalar(N, np.ones(d))[1]


@pytest.mark.parametrize('mode', ['phase', 'symmetric'])
@pytest.mark.parametrize('physics', ['scalar', 'elasticity'])
@pytest.mark.parametrize('N', [(3, 5, 255), (7, 255, 255), (255, 3, 255), (255, 255, 9)])
def test_odd_family_operator_and_cg(N, physics, mode, monkeypatch):
    from ffthompy_b200.tensors import Tensor
    from ffthompy_b200.general.solver import linear_solver
    Na = np.array(N)
    rng = np.random.default_rng(sum(N)+len(mode)+len(physics))
    D, Aval = _coefficients(physics, N, mode, rng)
    G = harness.green_for(physics, 'GaNi', N, np.ones(3), 'primal')[0]
    A, Afun = harness.build_operator(Aval, G, Na)
    cfg = Afun.fused().config()
    assert cfg['coefficients'] == mode, cfg
    for ax, key in enumerate(('mid0', 'mid1', 'last')):
        if N[ax] == 255:
            assert cfg[key] == 'odd', cfg
    Afo = O.GA(Aval, _oracle_green(physics, N), N)
    u = rng.standard_normal((D,)+N)
    ref = Afo(u)
    got = Afun(Tensor(name='u', val=u, order=1, N=Na)).val
    assert np.abs(got-ref).max() < 1e-12*np.abs(ref).max()
    # the family it replaces, same operator
    monkeypatch.setenv('FH_ODD', '0')
    A2, Afun2 = harness.build_operator(Aval, G, Na)
    assert 'odd' not in [Afun2.fused().config()[k] for k in ('mid0', 'mid1', 'last')]
    got_rt = Afun2(Tensor(name='u', val=u, order=1, N=Na)).val
    monkeypatch.delenv('FH_ODD')
    assert np.abs(got-got_rt).max() < 1e-13*np.abs(ref).max()
    # complete CG: same iteration count and solution as the oracle
    E = np.zeros((D,)+N)
    E[0] = 1.
    xo, io = O.cg(Afo, Afo(-E), np.zeros_like(E), 1e-6, 1000, N)
    EN = Tensor(name='EN', N=Na, shape=(D,), Fourier=False)
    EN.set_mean(np.eye(D)[0])
    X, info = linear_solver(solver='CG', Afun=Afun, B=Afun(-EN), x0=EN.zeros_like(),
                            par={'tol': 1e-6, 'maxiter': 1000}, callback=None)
    assert info['kit'] == io['kit']
    assert np.abs(X.val-xo).max() < 1e-9


def test_odd_family_cube_properties():
    """255^3 scalar (the config-2 grid itself; too large for the oracle in a test): the operator is a projection composed
    with A, so with A = I it is idempotent and annihilates constants; S1..S5 of the odd family against the run-time family."""
    import os
    from ffthompy_b200.tensors import Tensor
    N = (255, 255, 255)
    Na = np.array(N)
    rng = np.random.default_rng(255)
    Aval = np.einsum('ij,...->ij...', np.eye(3), np.ones(N))
    G = harness.green_for('scalar', 'GaNi', N, np.ones(3), 'primal')[0]
    A, Afun = harness.build_operator(Aval, G, Na)
    assert [Afun.fused().config()[k] for k in ('mid0', 'mid1', 'last')] == ['odd']*3
    u = Tensor(name='u', val=rng.standard_normal((3,)+N), order=1, N=Na)
    Pu = Afun(u)
    PPu = Afun(Pu)
    scale = np.abs(Pu.val).max()
    assert np.abs(PPu.val-Pu.val).max() < 1e-12*scale
    assert abs(Pu.val.mean()) < 1e-12*scale
    os.environ['FH_ODD'] = '0'
    try:
        A2, Afun2 = harness.build_operator(Aval, G, Na)
        assert 'odd' not in [Afun2.fused().config()[k] for k in ('mid0', 'mid1', 'last')]
        Pu_rt = Afun2(u)
    finally:
        del os.environ['FH_ODD']
    assert np.abs(Pu_rt.val-Pu.val).max() < 1e-12*scale


def test_large_download_matches_the_plain_copy():
    """device.download of a large result (pinned staging ring, parallel first touch: csrc/fh_host.cu) is bit-identical
    to torch's copy, for sizes that are not multiples of the slot size and for complex data"""
    import torch
    from ffthompy_b200 import device as dev
    for shape, cplx in (((3, 1 << 20), False), ((5, 1234567), False), ((2, 777777), True), ((1, (8 << 20)//8+3), False)):
        t = torch.randn(shape, dtype=torch.complex128 if cplx else torch.float64, device=dev.device())
        got = dev.download(t)
        ref = t.cpu().numpy()
        assert got.dtype == ref.dtype and got.shape == ref.shape
        assert np.array_equal(got, ref)
