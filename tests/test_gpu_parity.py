"""GPU parity tests: the CUDA path (through the package API -> ctypes -> C ABI) against the golden
fixtures of the unmodified reference and against the CPU oracle on the same seeded inputs.

Tolerances (fp64): fields/spectra 1e-12..1e-13 relative, homogenised matrices 1e-10 relative,
CG / Richardson iteration counts EQUAL (BASELINE.json north_star)."""
import numpy as np
import pytest

import ffthom_oracle as O
from conftest import Golden, example_tags
import harness

pytestmark = pytest.mark.gpu


@pytest.fixture(scope='module', autouse=True)
def _device():
    from ffthompy_b200 import device
    device.init(0)
    before = device.launch_count()
    yield
    assert device.launch_count() > before, 'no kernel of libffthom_b200.so was launched'


def _parse(key):
    kind, Ns, Ys, f, nyq = key.split('_')[:5]
    N = tuple(int(v) for v in Ns[1:].split('x'))
    Y = tuple(float(v) for v in Ys[1:].split('x'))
    form = f[1:]
    form = 0 if form == '0' else form
    return kind, N, Y, form, bool(int(nyq[3:]))


def _proj_keys():
    g = Golden()['projections']
    return sorted(set('_'.join(k.split('_')[:5]) for k in g.files if k.startswith(('scalar_N', 'elastic_N'))))


@pytest.mark.parametrize('key', _proj_keys())
def test_lazy_projections_materialise_to_reference_arrays(golden, key):
    """T2: `.val` of the closed-form Green tensors == reference arrays (abs 5e-15), all fft forms"""
    import ffthompy_b200.projections as proj
    g = golden['projections']
    kind, N, Y, form, nyq = _parse(key)
    if kind == 'scalar':
        got = dict(zip(('G0', 'G1', 'G2'), proj.scalar(np.array(N), np.array(Y), NyqNul=nyq, fft_form=form)))
    else:
        got = dict(zip(('G0', 'G1h', 'G1s', 'G2h', 'G2s'),
                       proj.elasticity(np.array(N), np.array(Y), NyqNul=nyq, fft_form=form)))
    for name, G in got.items():
        assert G.lazy
        ref = g[key+'_'+name]
        val = G.val
        assert val.shape == ref.shape and val.dtype == np.float64
        assert np.abs(val-ref).max() < 5e-15, (key, name)


def test_enlarged_multipliers_carry_the_reference_scale(golden):
    import ffthompy_b200.projections as proj
    g = golden['projections']
    for N in [(5, 5), (4, 4), (5, 5, 5), (4, 4, 4)]:
        Nbar = tuple(2*np.array(N)-1)
        tag = 'x'.join(map(str, N))
        _, G1, G2 = proj.scalar(np.array(N), np.ones(len(N)))
        assert np.abs(G1.enlarge(Nbar).val-g['enl_scalar_N%s_G1' % tag]).max() < 1e-14
        assert np.abs(G2.enlarge(Nbar).val-g['enl_scalar_N%s_G2' % tag]).max() < 1e-14
        Ge = proj.elasticity(np.array(N), np.ones(len(N)))
        assert np.abs((Ge[1]+Ge[2]).enlarge(Nbar).val-g['enl_elastic_N%s_G1' % tag]).max() < 1e-14
        # the same through the materialised path (generic spectrum re-mapping kernel)
        _, G1m, _ = proj.scalar(np.array(N), np.ones(len(N)))
        G1m.val  # drops laziness
        assert not G1m.lazy
        assert np.abs(G1m.enlarge(Nbar).val-g['enl_scalar_N%s_G1' % tag]).max() < 1e-13


def test_fourth_order_green_tensors(golden):
    from ffthompy_b200.tensors.projection import elasticity_small_strain, elasticity_large_deformation, scalar
    g = golden['projections']
    for form in ('r', 0, 'c'):
        N, Y = np.array([5, 4, 3]), np.array([1., 2., .5])
        assert np.abs(elasticity_small_strain(N, Y, fft_form=form).val-g['g4_small_f%s' % form]).max() < 2e-15
        assert np.abs(elasticity_large_deformation(N, Y, fft_form=form).val-g['g4_large_f%s' % form]).max() < 2e-15
        for name, G in zip(('G0', 'G1', 'G2'), scalar(np.array([5, 4]), np.array([1., 2.]), fft_form=form)):
            assert np.abs(G.val-g['g4_scalar_f%s_%s' % (form, name)]).max() < 2e-15


TENSOR_GRIDS = [(4, 4), (5, 5), (5, 4), (4, 4, 4), (5, 5, 5), (5, 4, 6), (11, 12)]


@pytest.mark.parametrize('N', TENSOR_GRIDS)
@pytest.mark.parametrize('form', ['r', 0, 'c'])
def test_tensor_algebra(golden, N, form):
    """T1/T5: DFT conventions, norms, means, fft-form changes, enlarge/decrease/project, grad/div/potential"""
    from ffthompy_b200.tensors import Tensor, DFT, grad, div, potential
    g = golden['tensors']
    tag = 'N%s_f%s' % ('x'.join(map(str, N)), form)
    u = Tensor(name='u', val=g['u_'+tag].copy(), order=1, N=N, Fourier=False, fft_form=form)
    Fu = u.fourier(copy=True)
    assert Fu.Fourier and not u.Fourier
    assert np.abs(Fu.val-g['Fu_'+tag]).max() < 1e-13
    Fu2 = DFT(N=N, inverse=False, fft_form=form)(u)
    assert np.abs(Fu2.val-g['Fu_'+tag]).max() < 1e-13
    assert abs(u.norm()-g['norm_u_'+tag]) < 1e-13
    assert abs(Fu.norm()-g['norm_Fu_'+tag]) < 1e-13
    assert np.abs(Fu.mean()-g['mean_Fu_'+tag]).max() < 1e-13
    assert np.abs(u.mean()-g['u_'+tag].mean(axis=tuple(range(1, 1+len(N))))).max() < 1e-14
    assert np.abs(Fu.fourier(copy=True).val-g['iFu_'+tag]).max() < 1e-13
    assert np.abs(DFT(N=N, inverse=True, fft_form=form)(Fu).val-g['iFu_'+tag]).max() < 1e-13
    for f2 in ('r', 0, 'c'):
        if f2 != form:
            conv = Fu.set_fft_form(f2, copy=True)
            assert conv.fft_form == f2
            assert np.abs(conv.val-g['Fu_%s_to%s' % (tag, f2)]).max() < 1e-13
            assert abs(conv.norm()-Fu.norm()) < 1e-13
    M = tuple(2*np.array(N))
    assert np.abs(Fu.copy().enlarge(M).val-g['enl2N_'+tag]).max() < 1e-12
    M2 = tuple(2*np.array(N)-1)
    assert np.abs(Fu.copy().enlarge(M2).val-g['enl2Nm1_'+tag]).max() < 1e-12
    assert np.abs(u.project(M).val-g['proj2N_real_'+tag]).max() < 1e-13
    Md = tuple(int(m) for m in g['decM_'+tag])
    assert np.abs(Fu.copy().decrease(Md).val-g['dec_'+tag]).max() < 1e-12
    assert np.abs(grad(Fu).val-g['grad_'+tag]).max() < 1e-12
    s = Tensor(name='s', val=g['s_'+tag].copy(), order=1, N=N, Fourier=False, fft_form=form)
    assert np.abs(grad(s).val-g['grads_'+tag]).max() < 1e-12
    if len(N) == 2:
        assert np.abs(div(Fu).val-g['div_'+tag]).max() < 1e-12
        assert np.abs(potential(Fu).val-g['pot_'+tag]).max() < 1e-13


def test_enlarge_leaves_operand_in_c_form():
    """reference quirk reproduced (tensors/objects.py:438,479)"""
    from ffthompy_b200.tensors import Tensor
    u = Tensor(name='u', val=np.random.default_rng(0).random((1, 4, 6)), order=1, N=(4, 6))
    Fu = u.fourier(copy=True)
    ref = O.set_fft_form(O.fftn(u.val, (4, 6), 'r'), (4, 6), 'r', 'c')
    Fu.enlarge((8, 12))
    assert Fu.fft_form == 'c' and np.abs(Fu.val-ref).max() < 1e-14


def test_pointwise_inverse_and_contractions(golden):
    from ffthompy_b200.tensors import Tensor
    from ffthompy_b200 import trigpol
    g = golden['tensors']
    A = Tensor(name='A', val=g['inv_A'].copy(), order=2, N=(4, 5), multype=21)
    assert np.abs(A.inv().val-g['inv_Ainv']).max() < 1e-13
    assert np.abs(trigpol.get_inverse(g['inv_A'])-g['inv_Ainv']).max() < 1e-13
    rng = np.random.default_rng(3)
    x = Tensor(name='x', val=rng.random((3, 4, 5)), order=1, N=(4, 5))
    assert np.abs(A(x).val-np.einsum('ij...,j...->i...', g['inv_A'], x.val)).max() < 1e-13
    AA = A*A
    assert AA.order == 2 and np.abs(AA.val-np.einsum('ij...,jk...->ik...', g['inv_A'], g['inv_A'])).max() < 1e-12
    assert np.abs(A.transpose().val-np.einsum('ij...->ji...', g['inv_A'])).max() == 0
    y = Tensor(name='y', val=rng.random((3, 4, 5)), order=1, N=(4, 5))
    assert abs(x*y-np.sum(x.val*y.val)/20) < 1e-14
    assert np.abs((2.5*x-y+0.5).val-(2.5*x.val-y.val+0.5)).max() < 1e-14
    c4 = Tensor(name='c4', val=rng.random((2, 2, 2, 2, 4, 5)), order=4, N=(4, 5), multype=42)
    e = Tensor(name='e', val=rng.random((2, 2, 4, 5)), order=2, N=(4, 5))
    assert np.abs(c4(e).val-np.einsum('ijkl...,kl...->ij...', c4.val, e.val)).max() < 1e-13
    assert np.abs(trigpol.enlarge(np.arange(12.).reshape(3, 4), (5, 7))-O.trigpol_enlarge(np.arange(12.).reshape(3, 4), (5, 7))).max() == 0
    big = rng.random((7, 8))
    assert np.abs(trigpol.decrease(big, (4, 5))-O.trigpol_decrease(big, (4, 5))).max() == 0


@pytest.mark.parametrize('tag', example_tags())
def test_reference_golden_examples(golden, tag):
    """T4/T6: the reference's 12 golden example problems (run_unittests.py:29-67) through the package:
    A_H for every post-processing variant 1e-10 relative (the reference's own bar is 1e-9 absolute),
    CG iteration counts equal, first minimiser 1e-10."""
    from ffthompy_b200.tensors import Tensor
    from ffthompy_b200.postprocess import assembly_matrix
    g = golden['examples']
    meta = [m for m in golden.example_meta() if m[0] == tag][0]
    _, physics, kind, N, Y, pds, tol, maxiter, _ = meta
    for pd in pds:
        G, Nbar = harness.green_for(physics, kind, N, Y, pd)
        A, Afun, sols, infos = harness.solve_loads(g['%s_%s_A' % (tag, pd)], G, Nbar, tol, maxiter)
        assert Afun.fused() is not None, 'the solve-loop operator was not fused'
        assert [i['kit'] for i in infos] == list(g['%s_%s_kit' % (tag, pd)])
        nr = np.array([i['norm_res'] for i in infos])
        # final residuals: relative where they are above rounding level of the first residual
        r0 = np.array([g['%s_%s_cbres%d' % (tag, pd, iL)][0] for iL in range(len(infos))])
        assert np.all(nr <= tol)
        assert np.allclose(nr, g['%s_%s_normres' % (tag, pd)], rtol=1e-4, atol=1e-13*r0.max())
        assert np.abs(sols[0].val-g['%s_%s_sol0' % (tag, pd)]).max() < 1e-10
        for key in [k for k in g.files if k.startswith('%s_%s_pp_' % (tag, pd)) and k.endswith('_AH')]:
            App_val = g[key[:-3]+'_A']
            App = Tensor(name='A_pp', val=App_val.copy(), order=2, N=App_val.shape[2:], multype=21)
            AH = assembly_matrix(App, sols)
            if pd == 'dual':
                AH = np.linalg.inv(AH)
            ref = g[key]
            assert np.abs(AH-ref).max() <= 1e-10*np.abs(ref).max(), key


def test_callback_residual_history_matches_reference(golden):
    """res_* of the golden pickles: CallBack.res_norm per iteration (general/solver_pp.py:13-24) —
    pins the Ga scale factor and the generic (unfused, callback) CG path."""
    from ffthompy_b200.general.solver_pp import CallBack
    g = golden['examples']
    tag, pd = 'scalar_3d_prob2', 'primal'
    meta = [m for m in golden.example_meta() if m[0] == tag][0]
    _, physics, kind, N, Y, pds, tol, maxiter, _ = meta
    G, Nbar = harness.green_for(physics, kind, N, Y, pd)
    A, Afun, sols, infos = harness.solve_loads(g['%s_%s_A' % (tag, pd)], G, Nbar, tol, maxiter,
                                               callback_factory=lambda Af, B: CallBack(A=Af, B=B))
    for iL, info in enumerate(infos):
        ref = g['%s_%s_cbres%d' % (tag, pd, iL)]
        got = np.array(info['cb'].res_norm)
        assert len(got) == len(ref) == info['kit']+1
        assert np.allclose(got[:-2], ref[:-2], rtol=1e-8)
        assert abs(got[0]-14.3265) < 1e-3   # first residual of load 0


def test_c3_recipe_small(golden):
    """BASELINE config 3 generator at 8^3 and 16^3 (even grids): primal and dual, kits equal, A_H 1e-10"""
    g = golden['configs']
    Cm, Ci = g['c3_Cm'], g['c3_Ci']
    for n in (8, 16):
        N = (n, n, n)
        for pd in ('primal', 'dual'):
            cm, ci = (Cm, Ci) if pd == 'primal' else (np.linalg.inv(Cm), np.linalg.inv(Ci))
            Aval, _ = O.two_phase(N, 20240901, 0.3, cm, ci)
            G, _ = harness.green_for('elasticity', 'GaNi', N, np.ones(3), pd)
            A, Afun, sols, infos = harness.solve_loads(Aval, G, N, 1e-6)
            AH = np.array([[A(sols[i])*sols[j] for j in range(6)] for i in range(6)])
            if pd == 'dual':
                AH = np.linalg.inv(AH)
            assert [i['kit'] for i in infos] == list(g['c3_n%d_%s_kit' % (n, pd)])
            ref = g['c3_n%d_%s_AH' % (n, pd)]
            assert np.abs(AH-ref).max() <= 1e-10*np.abs(ref).max()


@pytest.mark.parametrize('N', [(15, 15), (16, 16), (9, 9, 9), (12, 10, 8)])
def test_scalar_cg_and_richardson(golden, N):
    g = golden['configs']
    d = len(N)
    tag = 'sc_N%s' % 'x'.join(map(str, N))
    rng = np.random.default_rng(0)
    phase = (rng.random(N) < 0.3).astype(float)
    Aval = np.einsum('ij,...->ij...', np.eye(d), 1+10*phase)
    G, _ = harness.green_for('scalar', 'GaNi', N, np.ones(d), 'primal')
    A, Afun, sols, infos = harness.solve_loads(Aval, G, N, 1e-8)
    assert [i['kit'] for i in infos] == list(g[tag+'_kit'])
    AH = np.array([[A(sols[i])*sols[j] for j in range(d)] for i in range(d)])
    assert np.abs(AH-g[tag+'_AH']).max() <= 1e-10*np.abs(AH).max()
    assert np.abs(sols[0].val-g[tag+'_sol0']).max() < 1e-10
    # Richardson, fused device loop and generic path (custom scalar product forces the generic one)
    for par in ({'alpha': 0.5*(1+11.)}, {'alpha': 0.5*(1+11.), 'scal': lambda X, Y: X*Y}):
        A, Afun, sols, infos = harness.solve_loads(Aval, G, N, 1e-6, solver='richardson', par=par)
        assert [i['kit'] for i in infos] == list(g[tag+'_rich_kit'])
        nr = np.array([i['norm_res'] for i in infos])
        assert np.allclose(nr, g[tag+'_rich_normres'], rtol=1e-6)


def test_generic_cg_path_equals_fused(golden):
    """callback / custom scalar product route CG through the Tensor algebra: same counts, same result"""
    g = golden['configs']
    N = (12, 10, 8)
    rng = np.random.default_rng(0)
    phase = (rng.random(N) < 0.3).astype(float)
    Aval = np.einsum('ij,...->ij...', np.eye(3), 1+10*phase)
    G, _ = harness.green_for('scalar', 'GaNi', N, np.ones(3), 'primal')
    A, Afun, sols_f, infos_f = harness.solve_loads(Aval, G, N, 1e-8)
    A, Afun, sols_g, infos_g = harness.solve_loads(Aval, G, N, 1e-8, par={'scal': lambda X, Y: X*Y})
    assert [i['kit'] for i in infos_f] == [i['kit'] for i in infos_g] == list(g['sc_N12x10x8_kit'])
    assert np.abs(sols_f[0].val-sols_g[0].val).max() < 1e-11
    # materialised multiplier (no fusion, generic Operator evaluation) gives the same operator
    Gm, _ = harness.green_for('scalar', 'GaNi', N, np.ones(3), 'primal')
    Gm.val
    A2, Afun2 = harness.build_operator(Aval, Gm, N)
    assert Afun2.fused() is None
    x = sols_f[1]
    assert (Afun(x)-Afun2(x)).norm() < 1e-12


def test_tutorials(golden):
    """C1a: tutorials/02 (value 3.92394827320454, incl. the Moulinec-Suquet combination a*G1h+b*G1s)
    and tutorials/04 (exact integration on the doubled grid)."""
    import ffthompy_b200.projections as proj
    g = golden['configs']
    N = np.array([5, 5])
    _, G1h, G1s, _, _ = proj.elasticity(N, np.ones(2))
    A, Afun, sols, infos = harness.solve_loads(g['tut02_A'], G1h+G1s, N, 1e-8)
    assert abs(A(sols[0])*sols[0]-3.92394827320454) < 1e-12
    assert [i['kit'] for i in infos] == list(g['tut02_kit'])
    a, b = g['tut02_ms_ab']
    A, Afun, sols, infos = harness.solve_loads(g['tut02_A'], a*G1h+b*G1s, N, 1e-8)
    assert Afun.fused() is not None
    assert [i['kit'] for i in infos] == list(g['tut02_ms_kit'])
    AH = np.array([[A(sols[i])*sols[j] for j in range(3)] for i in range(3)])
    assert np.abs(AH-g['tut02_ms_AH']).max() < 1e-10
    N = np.array([25, 25])
    Nbar = 2*N-1
    _, G1, _ = proj.scalar(N, np.ones(2))
    A, Afun, sols, infos = harness.solve_loads(g['tut04_A'], G1.enlarge(Nbar), Nbar, 1e-8)
    assert abs(A(sols[0])*sols[0]-2.464008025892713) < 1e-11
    assert [i['kit'] for i in infos] == list(g['tut04_kit'])


@pytest.mark.parametrize('n', [31, 32, 45, 61, 64, 127, 512])
def test_full_size_properties(n):
    """size-independent properties at larger grids (odd, prime and power-of-two lengths): the projections
    are idempotent and mutually orthogonal through the fused pipeline, G1 + G2 + G0 = identity, and
    the oracle agrees on one operator application."""
    import ffthompy_b200.projections as proj
    from ffthompy_b200.tensors import Tensor, DFT, Operator
    N = np.array([n, n, n] if n <= 64 else ([n, 5, 7] if n == 127 else [n, 16, 32]))
    G0, G1h, G1s, G2h, G2s = proj.elasticity(N, np.ones(3))
    FN, FiN = DFT(inverse=False, N=N), DFT(inverse=True, N=N)
    P1 = Operator(mat=[[FiN, G1h+G1s, FN]])
    P2 = Operator(mat=[[FiN, G2h+G2s, FN]])
    P0 = Operator(mat=[[FiN, G0, FN]])
    rng = np.random.default_rng(5)
    u = Tensor(name='u', val=rng.standard_normal((6,)+tuple(N)), order=1, N=N)
    p1 = P1(u)
    assert (P1(p1)-p1).norm() < 1e-13*u.norm()
    assert P2(p1).norm() < 1e-13*u.norm()
    assert abs(p1*P2(u)) < 1e-13*(u*u)
    if np.all(N % 2 == 1):
        assert (p1+P2(u)+P0(u)-u).norm() < 1e-13*u.norm()
    if n <= 64:
        I6 = np.einsum('ij,...->ij...', np.eye(6), np.ones(tuple(N)))
        A, Afun = harness.build_operator(I6*2.0, G1h+G1s, N)
        ref = O.GA(I6*2.0, sum(O.proj_elasticity(N, np.ones(3))[1:3]), N)(u.val)
        assert np.abs(Afun(u).val-ref).max() < 1e-12*np.abs(ref).max()


@pytest.mark.parametrize('N', [(512, 8, 16), (8, 512, 16), (8, 16, 512), (512, 512), (8, 16, 1024), (6, 1024)])
@pytest.mark.parametrize('physics', ['elasticity', 'scalar'])
def test_axis_length_512_against_the_oracle(N, physics):
    """the register-resident three-pass kernels (csrc/fh_reg3.cuh: S1, S2/S4, S3, S5 at axis length 512, the
    BASELINE config-4 grid size; S1 / S5 also at 1024, the last axis of config 5), one axis at a time: operator application vs the oracle 1e-12 relative, CG
    iteration counts equal, solution 1e-9"""
    from ffthompy_b200.tensors import Tensor
    from ffthompy_b200.general.solver import linear_solver
    d = len(N)
    Na = np.array(N)
    if physics == 'elasticity':
        D = d*(d+1)//2
        G = harness.green_for('elasticity', 'GaNi', N, np.ones(d), 'primal')[0]
        Go = O.proj_elasticity(N, np.ones(d))
        Go = Go[1]+Go[2]
        Cm, Ci = (O.elastic_mandel(1, 1), O.elastic_mandel(10, 5)) if d == 3 else (np.eye(3)*2., np.eye(3)*9.+1.)
    else:
        D = d
        G = harness.green_for('scalar', 'GaNi', N, np.ones(d), 'primal')[0]
        Go = O.proj_scalar(N, np.ones(d))[1]
        Cm, Ci = np.eye(d), 11.*np.eye(d)
    rng = np.random.default_rng(11)
    ph = rng.random(N) < 0.3
    Aval = np.einsum('ij,...->ij...', Cm, 1.-ph)+np.einsum('ij,...->ij...', Ci, 1.*ph)
    A, Afun = harness.build_operator(Aval, G, Na)
    Afo = O.GA(Aval, Go, N)
    u = rng.standard_normal((D,)+N)
    ref = Afo(u)
    got = Afun(Tensor(name='u', val=u, order=1, N=Na)).val
    assert np.abs(got-ref).max() < 1e-12*np.abs(ref).max()
    E = np.zeros((D,)+N)
    E[0] = 1.
    xo, io = O.cg(Afo, Afo(-E), np.zeros_like(E), 1e-6, 1000, N)
    EN = Tensor(name='EN', N=Na, shape=(D,), Fourier=False)
    EN.set_mean(np.eye(D)[0])
    X, info = linear_solver(solver='CG', Afun=Afun, B=Afun(-EN), x0=EN.zeros_like(), par={'tol': 1e-6, 'maxiter': 1000},
                            callback=None)
    assert info['kit'] == io['kit']
    assert np.abs(X.val-xo).max() < 1e-9


def test_deferred_x_update_is_bitwise_the_plain_update():
    """fh_cg_steps applies x += alpha p inside the next S1 (one field read less per iteration); FH_XDEFER=0
    keeps it in the update kernel.  Same expressions in the same order: solutions must be bit-identical."""
    import os
    import subprocess
    import sys
    code = r'''
import hashlib, numpy as np
import ffthom_oracle as O, harness
from ffthompy_b200 import device
device.init(0)
out = []
for N in [(64, 64, 64), (30, 20, 15), (8, 16, 512)]:
    G = harness.green_for('elasticity', 'GaNi', N, np.ones(3), 'primal')[0]
    Aval, _ = O.two_phase(N, 7, 0.3, O.elastic_mandel(1, 1), O.elastic_mandel(10, 5))
    A, Afun, sols, infos = harness.solve_loads(Aval, G, np.array(N), 1e-7)
    out.append(hashlib.sha256(np.ascontiguousarray(sols[1].val).tobytes()).hexdigest()+':%d' % infos[1]['kit'])
print('HASH', ' '.join(out))
'''
    here = os.path.dirname(os.path.abspath(harness.__file__))
    root = os.path.dirname(here)
    res = {}
    for flag in ('0', '1'):
        env = dict(os.environ, FH_XDEFER=flag, PYTHONPATH=os.pathsep.join([root, os.path.join(root, 'oracle'), here]))
        r = subprocess.run([sys.executable, '-c', code], capture_output=True, text=True, env=env, timeout=600)
        assert r.returncode == 0, r.stdout[-1500:]+r.stderr[-1500:]
        res[flag] = [l for l in r.stdout.splitlines() if l.startswith('HASH')][0]
    assert res['0'] == res['1'], res
