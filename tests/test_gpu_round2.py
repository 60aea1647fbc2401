"""Round-2 behaviours: callbacks keep the device CG loop, BiCG, the fused-operator cache follows in-place changes of
the coefficient tensor, the one-pass A_H assembly, and the opt-in 8-column axis-0 kernel."""
import os
import subprocess
import sys

import numpy as np
import pytest

import ffthom_oracle as O
import harness

pytestmark = pytest.mark.gpu


@pytest.fixture(scope='module', autouse=True)
def _device():
    from ffthompy_b200 import device
    device.init(0)


def _problem(N, seed=0):
    d = len(N)
    rng = np.random.default_rng(seed)
    phase = (rng.random(N) < 0.3).astype(float)
    Aval = np.einsum('ij,...->ij...', np.eye(d), 1+10*phase)
    G, _ = harness.green_for('scalar', 'GaNi', N, np.ones(d), 'primal')
    return Aval, G


def test_callback_runs_on_the_device_loop():
    """ffthompy/applications.py:71-81 always attaches a CallBack: the solve must stay on fh_cg_begin / fh_cg_steps
    (one fused operator application per iteration + one inside the callback), with the same iterates as callback=None"""
    from ffthompy_b200 import device
    from ffthompy_b200.tensors import Tensor
    from ffthompy_b200.general.solver import linear_solver
    from ffthompy_b200.general.solver_pp import CallBack
    N = (24, 20, 16)
    Aval, G = _problem(N)
    A, Afun = harness.build_operator(Aval, G, np.array(N))
    EN = Tensor(name='EN', N=np.array(N), shape=(3,), Fourier=False)
    EN.set_mean(np.array([1., 0, 0]))
    B = Afun(-EN)
    X0, i0 = linear_solver(solver='CG', Afun=Afun, B=B, x0=EN.zeros_like(), par={'tol': 1e-8, 'maxiter': 1e3}, callback=None)
    cb = CallBack(A=Afun, B=B)
    n0 = device.launch_count()
    X1, i1 = linear_solver(solver='CG', Afun=Afun, B=B, x0=EN.zeros_like(), par={'tol': 1e-8, 'maxiter': 1e3}, callback=cb)
    launches = device.launch_count()-n0
    assert i1['kit'] == i0['kit'] and len(cb.res_norm) == i1['kit']+1 and cb.iter == i1['kit']
    assert np.array_equal(X1.val, X0.val)                       # same kernels, same order: bit-identical iterates
    assert np.allclose(i1['norm_res_log'], i0['norm_res_log'], rtol=0, atol=0)
    # the recurrence residual and the callback's true residual agree while far from rounding level
    assert np.allclose(cb.res_norm[1:4], i1['norm_res_log'][1:4], rtol=1e-8)
    # device loop: ~9 launches per iteration + ~8 per callback (operator, subtraction, norm); the generic Tensor-algebra
    # loop needs > 30 per iteration
    assert launches < 24*(i1['kit']+2), launches


def test_bicg_on_a_symmetric_operator_equals_cg():
    from ffthompy_b200.tensors import Tensor
    from ffthompy_b200.general.solver import linear_solver
    N = (15, 12)
    Aval, G = _problem(N, seed=3)
    A, Afun = harness.build_operator(Aval, G, np.array(N))
    EN = Tensor(name='EN', N=np.array(N), shape=(2,), Fourier=False)
    EN.set_mean(np.array([0., 1.]))
    B = Afun(-EN)
    par = {'tol': 1e-9, 'maxiter': 200}
    Xc, ic = linear_solver(solver='CG', Afun=Afun, B=B, x0=EN.zeros_like(), par=dict(par), callback=None)
    Xb, ib = linear_solver(solver='BiCG', Afun=Afun, ATfun=Afun, B=B, x0=EN.zeros_like(), par=dict(par), callback=None)
    assert ib['kit'] == ic['kit']
    assert np.abs(Xb.val-Xc.val).max() < 1e-10
    Go = O.proj_scalar(N, np.ones(2))[1]
    Afo = O.GA(Aval, Go, N)
    E = np.zeros((2,)+N)
    E[1] = 1.
    xo, io = O.cg(Afo, Afo(-E), np.zeros_like(E), 1e-9, 200, N)
    assert io['kit'] == ib['kit'] and np.abs(Xb.val-xo).max() < 1e-10


def test_fused_cache_follows_in_place_changes_of_the_coefficients():
    """A.add_mean / set_mean write into the device buffer the fused operator has analysed (phase table)"""
    from ffthompy_b200.tensors import Tensor
    N = (16, 16, 16)
    Aval, G = _problem(N)
    A, Afun = harness.build_operator(Aval, G, np.array(N))
    rng = np.random.default_rng(1)
    u = Tensor(name='u', val=rng.standard_normal((3,)+N), order=1, N=np.array(N))
    y0 = Afun(u).val.copy()
    f0 = Afun.fused()
    assert f0.config()['coefficients'] == 'phase'
    shift = 0.5*np.eye(3)
    A.add_mean(shift)
    y1 = Afun(u).val
    assert Afun.fused() is not f0
    Go = O.proj_scalar(N, np.ones(3))[1]
    ref = O.GA(Aval+shift.reshape(3, 3, 1, 1, 1), Go, N)(u.val)
    assert np.abs(y1-ref).max() < 1e-12*np.abs(ref).max()
    assert np.abs(y1-y0).max() > 1e-3


@pytest.mark.parametrize('N,D', [((9, 8), 2), ((6, 5, 7), 3), ((8, 6, 4), 6)])
def test_one_pass_assembly_equals_pairwise(N, D):
    from ffthompy_b200.tensors import Tensor
    from ffthompy_b200.postprocess import assembly_matrix, _one_pass
    rng = np.random.default_rng(2)
    M = rng.standard_normal((D, D)+N)
    Aval = np.einsum('ij...,kj...->ik...', M, M)+np.eye(D).reshape((D, D)+(1,)*len(N))
    A = Tensor(name='A', val=Aval, order=2, N=np.array(N), multype=21)
    sols = [Tensor(name='e%d' % i, val=rng.standard_normal((D,)+N), order=1, N=np.array(N)) for i in range(D)]
    AH = assembly_matrix(A, sols)
    assert _one_pass(A, sols) is not None
    ref = np.array([[np.sum(np.einsum('ij...,j...->i...', Aval, sols[i].val)*sols[j].val)/np.prod(N) for j in range(D)]
                    for i in range(D)])
    assert np.abs(AH-ref).max() <= 1e-13*np.abs(ref).max()


@pytest.mark.parametrize('which', ['1', '2'])
def test_mid2_kernel_is_opt_in_and_correct(which):
    """FH_MID2=1 selects the 8-column axis-0 + Green kernel (csrc/fh_mid2.cuh), FH_MID2=2 the row-group-per-warp kernel
    (csrc/fh_mid3.cuh, N0 = 256); kept behind the switch with their parity check (measurements: DESIGN.md section 4)"""
    code = r'''
import numpy as np
import ffthom_oracle as O, harness
from ffthompy_b200 import device
from ffthompy_b200.tensors import Tensor
device.init(0)
for N, phys in [((256, 8, 16), 'elasticity'), ((128, 16, 16), 'elasticity'), ((256, 16, 8), 'scalar'),
                ((256, 5, 20), 'elasticity')]:
    d = 3
    D = 6 if phys == 'elasticity' else 3
    G = harness.green_for(phys, 'GaNi', N, np.ones(3), 'primal')[0]
    Go = O.proj_elasticity(N, np.ones(3)) if D == 6 else O.proj_scalar(N, np.ones(3))
    Go = Go[1]+Go[2] if D == 6 else Go[1]
    rng = np.random.default_rng(5)
    M = 0.3*rng.standard_normal((D, D)+N)
    Aval = np.einsum('ij...,kj...->ik...', M, M)+np.eye(D).reshape((D, D, 1, 1, 1))
    A, Afun = harness.build_operator(Aval, G, np.array(N))
    u = rng.standard_normal((D,)+N)
    ref = O.GA(Aval, Go, N)(u)
    got = Afun(Tensor(name='u', val=u, order=1, N=np.array(N))).val
    err = np.abs(got-ref).max()/np.abs(ref).max()
    assert err < 1e-12, (N, err)
print('MID2 OK')
'''
    here = os.path.dirname(os.path.abspath(harness.__file__))
    root = os.path.dirname(here)
    env = dict(os.environ, FH_MID2=which, PYTHONPATH=os.pathsep.join([root, os.path.join(root, 'oracle'), here]))
    r = subprocess.run([sys.executable, '-c', code], capture_output=True, text=True, env=env, timeout=600)
    assert r.returncode == 0 and 'MID2 OK' in r.stdout, r.stdout[-1500:]+r.stderr[-1500:]
