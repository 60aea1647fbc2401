"""The CPU oracle (oracle/ffthom_oracle.py) pinned against fixtures generated from the unmodified
reference (tests/golden/*.npz, oracle/make_golden.py), including the reference's own 12 golden
example problems.  Runs without a GPU."""
import numpy as np
import pytest

import ffthom_oracle as O
from conftest import Golden, example_tags


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
def test_projection_arrays(golden, key):
    """projections.py:9-267 (even-N elasticity per SURVEY App. D.1) — abs 1e-15 on O(1) entries"""
    g = golden['projections']
    kind, N, Y, form, nyq = _parse(key)
    if kind == 'scalar':
        got = dict(zip(('G0', 'G1', 'G2'), O.proj_scalar(N, Y, NyqNul=nyq, fft_form=form)))
    else:
        got = dict(zip(('G0', 'G1h', 'G1s', 'G2h', 'G2s'), O.proj_elasticity(N, Y, NyqNul=nyq, fft_form=form)))
    for name, arr in got.items():
        ref = g[key+'_'+name]
        assert arr.shape == ref.shape
        assert np.abs(arr-ref).max() < 2e-15, (key, name)


def test_enlarged_multipliers(golden):
    """hG.enlarge(Nbar) carries the factor prod(Nbar)/prod(N)"""
    g = golden['projections']
    for N in [(5, 5), (4, 4), (5, 5, 5), (4, 4, 4)]:
        Nbar = tuple(2*np.array(N)-1)
        tag = 'x'.join(map(str, N))
        _, G1, G2 = O.proj_scalar(N, np.ones(len(N)))
        assert np.abs(O.enlarge_multiplier(G1, N, Nbar)-g['enl_scalar_N%s_G1' % tag]).max() < 1e-14
        assert np.abs(O.enlarge_multiplier(G2, N, Nbar)-g['enl_scalar_N%s_G2' % tag]).max() < 1e-14
        Ge = O.proj_elasticity(N, np.ones(len(N)))
        assert np.abs(O.enlarge_multiplier(Ge[1]+Ge[2], N, Nbar)-g['enl_elastic_N%s_G1' % tag]).max() < 1e-14
        c = np.prod(Nbar)/np.prod(N)
        assert abs(np.abs(g['enl_scalar_N%s_G1' % tag]).max()-c) < 1e-12


def test_green4(golden):
    g = golden['projections']
    for form in ('r', 0, 'c'):
        N, Y = (5, 4, 3), (1., 2., .5)
        assert np.abs(O.green4(N, Y, 'small_strain', form)-g['g4_small_f%s' % form]).max() < 1e-15
        assert np.abs(O.green4(N, Y, 'large', form)-g['g4_large_f%s' % form]).max() < 1e-15


TENSOR_GRIDS = [(4, 4), (5, 5), (5, 4), (4, 4, 4), (5, 5, 5), (5, 4, 6), (11, 12)]


@pytest.mark.parametrize('N', TENSOR_GRIDS)
@pytest.mark.parametrize('form', ['r', 0, 'c'])
def test_tensor_algebra(golden, N, form):
    """fft forms, norms, enlarge/decrease/project, grad/div/potential vs the reference Tensor — 1e-13"""
    g = golden['tensors']
    tag = 'N%s_f%s' % ('x'.join(map(str, N)), form)
    u = g['u_'+tag]
    Fu = O.fftn(u, N, form)
    assert np.abs(Fu-g['Fu_'+tag]).max() < 1e-13
    assert abs(O.norm(u, N)-g['norm_u_'+tag]) < 1e-13
    assert abs(O.norm(Fu, N, True, form)-g['norm_Fu_'+tag]) < 1e-13
    assert np.abs(O.ifftn(Fu, N, form)-g['iFu_'+tag]).max() < 1e-13
    for f2 in ('r', 0, 'c'):
        if f2 != form:
            assert np.abs(O.set_fft_form(Fu, N, form, f2)-g['Fu_%s_to%s' % (tag, f2)]).max() < 1e-13
    M = tuple(2*np.array(N))
    assert np.abs(O.enlarge(Fu, N, M, form)-g['enl2N_'+tag]).max() < 1e-12
    M2 = tuple(2*np.array(N)-1)
    assert np.abs(O.enlarge(Fu, N, M2, form)-g['enl2Nm1_'+tag]).max() < 1e-12
    assert np.abs(O.project(u, N, M, False, form)-g['proj2N_real_'+tag]).max() < 1e-13
    Md = tuple(int(m) for m in g['decM_'+tag])
    assert np.abs(O.decrease(Fu, N, Md, form)-g['dec_'+tag]).max() < 1e-12
    Y = np.ones(len(N))
    assert np.abs(O.grad(Fu, N, Y, form)-g['grad_'+tag]).max() < 1e-12
    if len(N) == 2:
        assert np.abs(O.div(Fu, N, Y, form)-g['div_'+tag]).max() < 1e-12
        assert np.abs(O.potential_scalar(Fu, N, Y, form)-g['pot_'+tag][0]).max() < 1e-13


def test_get_inverse(golden):
    g = golden['tensors']
    assert np.abs(O.get_inverse(g['inv_A'])-g['inv_Ainv']).max() == 0.0


def _example_green(physics, kind, N, Y, pd):
    N = np.array(N)
    if physics == 'scalar':
        _, G1, G2 = O.proj_scalar(N, Y)
        G = G1 if pd == 'primal' else G2
    else:
        _, G1h, G1s, G2h, G2s = O.proj_elasticity(N, Y)
        G = G1h+G1s if pd == 'primal' else G2h+G2s
    Nbar = N if kind == 'GaNi' else 2*N-1
    if kind == 'Ga':
        G = O.enlarge_multiplier(G, N, Nbar)
    return G, tuple(int(n) for n in Nbar)


@pytest.mark.parametrize('tag', example_tags())
def test_reference_golden_examples(golden, tag):
    """run_unittests.py:29-67: every example problem of the reference — A_H to 1e-10 relative, CG
    iteration counts equal, final residual norms to 1e-6 relative."""
    g = golden['examples']
    meta = [m for m in golden.example_meta() if m[0] == tag][0]
    _, physics, kind, N, Y, pds, tol, maxiter, _ = meta
    for pd in pds:
        A = g['%s_%s_A' % (tag, pd)]
        G, Nbar = _example_green(physics, kind, N, Y, pd)
        AH, infos, sols = O.homogenize(A, G, Nbar, tol=tol, maxiter=maxiter)
        assert [i['kit'] for i in infos] == list(g['%s_%s_kit' % (tag, pd)])
        nr = np.array([i['norm_res'] for i in infos])
        assert np.allclose(nr, g['%s_%s_normres' % (tag, pd)], rtol=1e-6, atol=1e-300) or np.all(nr < 1e-12)
        assert np.abs(sols[0]-g['%s_%s_sol0' % (tag, pd)]).max() < 1e-10
        for key in [k for k in g.files if k.startswith('%s_%s_pp_' % (tag, pd)) and k.endswith('_AH')]:
            App = g[key[:-3]+'_A']
            Npp = App.shape[2:]
            s = [O.project(x, Nbar, Npp) for x in sols] if tuple(Npp) != tuple(Nbar) else sols
            AHpp = O.assembly_matrix(App, s, Npp)
            if pd == 'dual':
                AHpp = np.linalg.inv(AHpp)
            ref = g[key]
            assert np.abs(AHpp-ref).max() <= 1e-10*np.abs(ref).max(), key


def test_c3_recipe(golden):
    """C3 generator at 8^3 (even grid, App. D.1 projections), primal and dual"""
    g = golden['configs']
    n = 8
    N = (n, n, n)
    Cm, Ci = g['c3_Cm'], g['c3_Ci']
    assert np.abs(O.elastic_mandel(1, 1)-Cm).max() < 1e-15 and np.abs(O.elastic_mandel(10, 5)-Ci).max() < 1e-14
    _, G1h, G1s, G2h, G2s = O.proj_elasticity(N, np.ones(3))
    for pd in ('primal', 'dual'):
        cm, ci = (Cm, Ci) if pd == 'primal' else (np.linalg.inv(Cm), np.linalg.inv(Ci))
        A, _ = O.two_phase(N, 20240901, 0.3, cm, ci)
        AH, infos, _ = O.homogenize(A, G1h+G1s if pd == 'primal' else G2h+G2s, N, tol=1e-6)
        if pd == 'dual':
            AH = np.linalg.inv(AH)
        assert [i['kit'] for i in infos] == list(g['c3_n8_%s_kit' % pd])
        assert np.abs(AH-g['c3_n8_%s_AH' % pd]).max() <= 1e-10*np.abs(AH).max()


@pytest.mark.parametrize('N', [(15, 15), (16, 16), (9, 9, 9), (12, 10, 8)])
def test_scalar_cg_and_richardson(golden, N):
    g = golden['configs']
    d = len(N)
    tag = 'sc_N%s' % 'x'.join(map(str, N))
    rng = np.random.default_rng(0)
    phase = (rng.random(N) < 0.3).astype(float)
    A = np.einsum('ij,...->ij...', np.eye(d), 1+10*phase)
    _, G1, _ = O.proj_scalar(N, np.ones(d))
    AH, infos, sols = O.homogenize(A, G1, N, tol=1e-8)
    assert [i['kit'] for i in infos] == list(g[tag+'_kit'])
    assert np.abs(AH-g[tag+'_AH']).max() <= 1e-10*np.abs(AH).max()
    assert np.abs(sols[0]-g[tag+'_sol0']).max() < 1e-10
    Afun = O.GA(A, G1, N)
    for iL in range(d):
        E = np.zeros((d,)+tuple(N))
        E[iL] = 1
        x, info = O.richardson(Afun, Afun(-E), np.zeros_like(E), alpha=0.5*(1+11.), tol=1e-6, N=N)
        assert info['kit'] == g[tag+'_rich_kit'][iL]
        assert abs(info['norm_res']-g[tag+'_rich_normres'][iL]) <= 1e-6*info['norm_res']


def test_tutorials(golden):
    """tutorials/02 (C1a, incl. the Moulinec-Suquet scaled projection) and tutorials/04 (Ga)"""
    g = golden['configs']
    N = (5, 5)
    _, G1h, G1s, _, _ = O.proj_elasticity(N, np.ones(2))
    AH, infos, _ = O.homogenize(g['tut02_A'], G1h+G1s, N, tol=1e-8)
    assert abs(AH[0, 0]-3.92394827320454) < 1e-12
    assert [i['kit'] for i in infos] == list(g['tut02_kit'])
    a, b = g['tut02_ms_ab']
    AHms, infos, _ = O.homogenize(g['tut02_A'], a*G1h+b*G1s, N, tol=1e-8)
    assert [i['kit'] for i in infos] == list(g['tut02_ms_kit'])
    assert np.abs(AHms-g['tut02_ms_AH']).max() < 1e-10
    N, Nbar = (25, 25), (49, 49)
    _, G1, _ = O.proj_scalar(N, np.ones(2))
    AH, infos, _ = O.homogenize(g['tut04_A'], O.enlarge_multiplier(G1, N, Nbar), Nbar, tol=1e-8)
    assert abs(AH[0, 0]-2.464008025892713) < 1e-11
    assert [i['kit'] for i in infos] == list(g['tut04_kit'])
