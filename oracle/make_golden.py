"""Generate the golden fixtures under tests/golden/ from the UNMODIFIED reference
(/root/reference, imported under oracle/_refshim.py).  Run in the build container:

    python oracle/make_golden.py

The reference tree does not exist on the GPU box, so its outputs travel as these small .npz
files.  Everything here calls the reference's own public API; nothing is re-derived.
"""
import contextlib
import io
import os
import pickle
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import _refshim  # noqa: E402

REF = _refshim.install()
OUT = os.path.join(HERE, '..', 'tests', 'golden')
os.makedirs(OUT, exist_ok=True)

import ffthompy.projections as proj  # noqa: E402
from ffthompy.tensors import Tensor, DFT, Operator, grad, div, potential, symgrad  # noqa: E402
from ffthompy.tensors.projection import (scalar as scalar4, elasticity_small_strain,  # noqa: E402
                                         elasticity_large_deformation)
from ffthompy.trigpol import get_Nodd, get_inverse  # noqa: E402
from ffthompy.general.solver import linear_solver  # noqa: E402
from ffthompy.materials import Material  # noqa: E402
from ffthompy.mechanics.matcoef import ElasticTensor  # noqa: E402
from ffthompy.problem import Problem, import_file  # noqa: E402


def quiet(fn, *a, **k):
    with contextlib.redirect_stdout(io.StringIO()):
        return fn(*a, **k)


def elasticity_anyN(N, Y, fft_form='r'):
    """the even-N definition of proj.elasticity built from reference functions only."""
    N = np.array(N, dtype=int)
    if np.all(N % 2 == 1):
        return proj.elasticity(N, Y, NyqNul=True, tensor=True, fft_form=fft_form)
    Nred = get_Nodd(N)
    Gs = proj.elasticity(Nred, Y, NyqNul=False, tensor=True, fft_form=0)
    Gs = [G.enlarge(N) for G in Gs]
    if fft_form == 'r':
        for G in Gs:
            G.set_fft_form('r')
            G.val = G.val/np.prod(G.N)
    elif fft_form == 'c':
        for G in Gs:
            G.set_fft_form('c')
    return Gs


# ----------------------------------------------------------------------------- 1. projections
def gen_projections():
    out = {}
    cases = []
    for N, Y in [((5, 5), (1., 1.)), ((4, 4), (1., 1.)), ((6, 5), (1., 2.)), ((7, 6), (0.5, 1.)), ((5, 5, 5), (1., 1., 1.)),
                 ((6, 5, 4), (1., 2., .5)), ((4, 4, 6), (1., 1., 1.))]:
        for form in ('r', 0, 'c'):
            for nyq in (True, False):
                key = 'scalar_N%s_Y%s_f%s_nyq%d' % ('x'.join(map(str, N)), 'x'.join(map(str, Y)), form, nyq)
                G = proj.scalar(np.array(N), np.array(Y), NyqNul=nyq, tensor=True, fft_form=form)
                for name, g in zip(('G0', 'G1', 'G2'), G):
                    out[key+'_'+name] = g.val
                cases.append(('scalar', N, Y, form, nyq, key))
            key = 'elastic_N%s_Y%s_f%s_nyq1' % ('x'.join(map(str, N)), 'x'.join(map(str, Y)), form)
            G = elasticity_anyN(N, np.array(Y), fft_form=form)
            for name, g in zip(('G0', 'G1h', 'G1s', 'G2h', 'G2s'), G):
                out[key+'_'+name] = g.val
            cases.append(('elastic', N, Y, form, True, key))
            if all(n % 2 == 1 for n in N):
                key = 'elastic_N%s_Y%s_f%s_nyq0' % ('x'.join(map(str, N)), 'x'.join(map(str, Y)), form)
                G = proj.elasticity(np.array(N), np.array(Y), NyqNul=False, tensor=True, fft_form=form)
                for name, g in zip(('G0', 'G1h', 'G1s', 'G2h', 'G2s'), G):
                    out[key+'_'+name] = g.val
                cases.append(('elastic', N, Y, form, False, key))
    # Ga-enlarged multipliers (scale factor prod(Nbar)/prod(N))
    for N in [(5, 5), (4, 4), (5, 5, 5), (4, 4, 4)]:
        Nbar = tuple(2*np.array(N)-1)
        G = proj.scalar(np.array(N), np.ones(len(N)), NyqNul=True, tensor=True)
        out['enl_scalar_N%s_G1' % 'x'.join(map(str, N))] = G[1].enlarge(Nbar).val
        out['enl_scalar_N%s_G2' % 'x'.join(map(str, N))] = G[2].enlarge(Nbar).val
        Ge = elasticity_anyN(N, np.ones(len(N)))
        out['enl_elastic_N%s_G1' % 'x'.join(map(str, N))] = (Ge[1]+Ge[2]).enlarge(Nbar).val
    # 4th-order tensors
    for form in ('r', 0, 'c'):
        N = np.array([5, 4, 3])
        Y = np.array([1., 2., .5])
        out['g4_small_f%s' % form] = elasticity_small_strain(N, Y, fft_form=form).val
        out['g4_large_f%s' % form] = elasticity_large_deformation(N, Y, fft_form=form).val
        for name, g in zip(('G0', 'G1', 'G2'), scalar4(np.array([5, 4]), np.array([1., 2.]), fft_form=form)):
            out['g4_scalar_f%s_%s' % (form, name)] = g.val
    np.savez_compressed(os.path.join(OUT, 'projections.npz'), **out)
    print('projections.npz: %d arrays' % len(out))


# ----------------------------------------------------------------------------- 2. tensor algebra
def gen_tensors():
    out = {}
    rng = np.random.default_rng(7)
    for N in [(4, 4), (5, 5), (5, 4), (4, 4, 4), (5, 5, 5), (5, 4, 6), (11, 12)]:
        for form in ('r', 0, 'c'):
            tag = 'N%s_f%s' % ('x'.join(map(str, N)), form)
            u = Tensor(name='u', val=rng.random((2,)+N), order=1, N=N, Fourier=False, fft_form=form)
            out['u_'+tag] = u.val.copy()
            Fu = u.fourier(copy=True)
            out['Fu_'+tag] = Fu.val.copy()
            out['norm_u_'+tag] = np.array(u.norm())
            out['norm_Fu_'+tag] = np.array(Fu.norm())
            out['mean_Fu_'+tag] = Fu.mean()
            out['iFu_'+tag] = Fu.fourier(copy=True).val
            for f2 in ('r', 0, 'c'):
                if f2 != form:
                    out['Fu_%s_to%s' % (tag, f2)] = Fu.set_fft_form(f2, copy=True).val
            M = tuple(2*np.array(N))
            out['enl2N_'+tag] = Fu.copy().enlarge(M).val
            M2 = tuple(2*np.array(N)-1)
            out['enl2Nm1_'+tag] = Fu.copy().enlarge(M2).val
            out['proj2N_real_'+tag] = u.project(M).val
            if min(N) >= 4:
                Md = tuple(int(n) - (2 if n > 4 else 1) for n in N)
                out['dec_'+tag] = Fu.copy().decrease(Md).val
                out['decM_'+tag] = np.array(Md)
            # differential operators
            out['grad_'+tag] = grad(Fu).val
            s = Tensor(name='s', val=rng.random((1,)+N), order=1, N=N, Fourier=False, fft_form=form)
            out['s_'+tag] = s.val.copy()
            out['grads_'+tag] = grad(s).val
            if len(N) == u.shape[0]:
                out['div_'+tag] = div(Fu).val
                out['pot_'+tag] = potential(Fu).val
    # pointwise inverse
    A = rng.random((3, 3, 4, 5))
    A = np.einsum('ij...,kj...->ik...', A, A)+3*np.eye(3)[:, :, None, None]
    out['inv_A'] = A
    out['inv_Ainv'] = get_inverse(A)
    np.savez_compressed(os.path.join(OUT, 'tensors.npz'), **out)
    print('tensors.npz: %d arrays' % len(out))


# ----------------------------------------------------------------------------- 3. example problems
def gen_examples():
    """The reference's own golden integration suite (run_unittests.py:29-67): rerun every problem,
    check it against test_results/python3/*, and store the inputs (coefficient tensors) with the
    expected homogenised matrices, CG iteration counts and residual histories."""
    os.chdir(REF)
    out = {}
    meta = []
    for f in ['examples/scalar/scalar_2d.py', 'examples/scalar/scalar_3d.py', 'examples/scalar/from_file.py',
              'examples/elasticity/linelas_3d.py']:
        conf = quiet(import_file, f)
        for cp in conf.problems:
            tag = os.path.basename(f).split('.')[0]+'_'+cp['name']
            prob = quiet(Problem, cp, conf)
            quiet(prob.calculate)
            gold = os.path.join(REF, 'test_results', 'python3', tag)
            sys.path.insert(0, os.path.join(REF, os.path.dirname(f)))
            with open(gold, 'rb') as fh:
                res = pickle.load(fh)
            maxdiff = 0.
            for pd in ('primal', 'dual'):
                key = 'mat_'+pd
                if key in res:
                    for k, v in res[key].items():
                        maxdiff = max(maxdiff, np.abs(prob.output[key][k]-v).max())
            assert maxdiff < 1e-9, (tag, maxdiff)
            N = np.array(prob.solve['N'], dtype=int)
            kind = prob.solve['kind']
            Nbar = N if kind == 'GaNi' else 2*N-1
            mat = Material(prob.material)
            for pd in prob.solve['primaldual']:
                A = quiet(mat.get_A_GaNi, N, pd) if kind == 'GaNi' else quiet(mat.get_A_Ga, Nbar=Nbar, primaldual=pd)
                out['%s_%s_A' % (tag, pd)] = A.val
                out['%s_%s_kit' % (tag, pd)] = np.array([r['info']['kit'] for r in prob.output['res_'+pd]])
                out['%s_%s_normres' % (tag, pd)] = np.array([r['info']['norm_res'] for r in prob.output['res_'+pd]])
                for iL, r in enumerate(prob.output['res_'+pd]):
                    out['%s_%s_cbres%d' % (tag, pd, iL)] = np.array(r['cb'].res_norm)
                out['%s_%s_sol0' % (tag, pd)] = prob.output['sol_'+pd][0].val
                for pp in prob.postprocess:
                    if pp['kind'] in ['GaNi', 'gani']:
                        name = 'AH_GaNi_'+pd
                        App = quiet(mat.get_A_GaNi, N, pd)
                    else:
                        Nbarpp = tuple(2*N-1)
                        if 'order' in pp:
                            if pp['order'] is None:
                                name = 'AH_Ga_'+pd
                                App = quiet(mat.get_A_Ga, Nbar=Nbarpp, primaldual=pd, order=None)
                            else:
                                name = 'AH_Ga_o%s_P%d_%s' % (str(pp['order']), np.mean(pp['P']), pd)
                                App = quiet(mat.get_A_Ga, Nbar=Nbarpp, primaldual=pd, order=pp['order'], P=pp['P'])
                        else:
                            name = 'AH_Ga_'+pd
                            App = A
                    out['%s_%s_pp_%s_A' % (tag, pd, name)] = App.val
                    out['%s_%s_pp_%s_AH' % (tag, pd, name)] = prob.output['mat_'+pd][name]
            meta.append((tag, str(prob.physics), str(kind), tuple(int(n) for n in N),
                         tuple(float(y) for y in prob.material['Y']), tuple(str(p) for p in prob.solve['primaldual']),
                         float(prob.solver['tol']), float(prob.solver['maxiter']), float(maxdiff)))
            print('  %s: reproduces the reference golden to %.1e' % (tag, maxdiff))
    out['meta'] = np.array([repr(m) for m in meta])
    np.savez_compressed(os.path.join(OUT, 'examples.npz'), **out)
    print('examples.npz: %d arrays' % len(out))


# ----------------------------------------------------------------------------- 4. synthetic configs, tutorials
def solve_all(A, G, N, D, tol, solver='CG', par=None):
    Afun = Operator(name='FiGFA', mat=[[Operator(name='G', mat=[[DFT(inverse=True, N=N), G, DFT(inverse=False, N=N)]]),
                                        A]])
    sols, kits, hists = [], [], []
    for iL in range(D):
        EN = Tensor(name='EN', N=N, shape=(D,), Fourier=False)
        EN.set_mean(np.eye(D)[iL])
        p = {'tol': tol, 'maxiter': 1e3}
        if par:
            p.update(par)
        X, info = quiet(linear_solver, solver=solver, Afun=Afun, B=Afun(-EN), x0=EN.zeros_like(), par=p, callback=None)
        sols.append(X+EN)
        kits.append(info['kit'])
        hists.append(info['norm_res'])
    AH = np.array([[A(sols[i])*sols[j] for j in range(D)] for i in range(D)])
    return AH, np.array(kits), np.array(hists), sols


def gen_configs():
    out = {}
    # C3 recipe: random two-phase elasticity, even grids through App. D.1
    Cm = ElasticTensor(bulk=1, mu=1).mandel
    Ci = ElasticTensor(bulk=10, mu=5).mandel
    out['c3_Cm'], out['c3_Ci'] = Cm, Ci
    for n in (8, 16):
        N = np.array([n, n, n])
        rng = np.random.default_rng(20240901)
        phase = (rng.random((n, n, n)) < 0.3).astype(float)
        for pd in ('primal', 'dual'):
            cm, ci = (Cm, Ci) if pd == 'primal' else (np.linalg.inv(Cm), np.linalg.inv(Ci))
            A = Tensor(name='A', val=np.einsum('ij,...->ij...', cm, 1-phase)+np.einsum('ij,...->ij...', ci, phase),
                       order=2, N=N, multype=21)
            G0, G1h, G1s, G2h, G2s = elasticity_anyN(N, np.ones(3))
            G = G1h+G1s if pd == 'primal' else G2h+G2s
            AH, kits, nres, _ = solve_all(A, G, N, 6, 1e-6)
            if pd == 'dual':
                AH = np.linalg.inv(AH)
            out['c3_n%d_%s_AH' % (n, pd)] = AH
            out['c3_n%d_%s_kit' % (n, pd)] = kits
            out['c3_n%d_%s_normres' % (n, pd)] = nres
        print('  C3 n=%d: AH00 primal %.15g dual %.15g' % (n, out['c3_n%d_primal_AH' % n][0, 0],
                                                          out['c3_n%d_dual_AH' % n][0, 0]))
    # scalar GaNi on a random two-phase medium (SURVEY App. C last row), odd and even grids, CG + Richardson
    for N in [(15, 15), (16, 16), (9, 9, 9), (12, 10, 8)]:
        d = len(N)
        rng = np.random.default_rng(0)
        phase = (rng.random(N) < 0.3).astype(float)
        A = Tensor(name='A', val=np.einsum('ij,...->ij...', np.eye(d), 1+10*phase), order=2, N=np.array(N), multype=21)
        _, G1, G2 = proj.scalar(np.array(N), np.ones(d), NyqNul=True, tensor=True)
        tag = 'sc_N%s' % 'x'.join(map(str, N))
        AH, kits, nres, sols = solve_all(A, G1, np.array(N), d, 1e-8)
        out[tag+'_AH'], out[tag+'_kit'], out[tag+'_normres'] = AH, kits, nres
        out[tag+'_sol0'] = sols[0].val
        AHr, kitr, nresr, _ = solve_all(A, G1, np.array(N), d, 1e-6, solver='richardson', par={'alpha': 0.5*(1+11.)})
        out[tag+'_rich_AH'], out[tag+'_rich_kit'], out[tag+'_rich_normres'] = AHr, kitr, nresr
    # C1a: tutorial 02 verbatim inputs (2-D plane-strain elasticity, N=5x5, GaNi)
    dim = 2
    N = 5*np.ones(dim, dtype=np.int32)
    K, Gm = np.array([1, 10.]), np.array([1, 5.])
    mM = ElasticTensor(bulk=K[0], mu=Gm[0], plane='strain')
    mI = ElasticTensor(bulk=K[1], mu=Gm[1], plane='strain')
    pbmat = {'Y': np.ones(dim), 'inclusions': ['cube', 'otherwise'], 'positions': [np.zeros(dim), ''],
             'params': [0.6*np.ones(dim), ''], 'vals': [mI.mandel, mM.mandel]}
    A = quiet(Material(pbmat).get_A_GaNi, N, 'primal')
    _, hG1h, hG1s, hG2h, hG2s = proj.elasticity(N, np.ones(dim), NyqNul=True, tensor=True)
    D = 3
    AH, kits, nres, sols = solve_all(A, hG1h+hG1s, N, D, 1e-8)
    out['tut02_A'] = A.val
    out['tut02_AH'] = AH
    out['tut02_kit'] = kits
    # Moulinec-Suquet scaled projection a*G1h + b*G1s (tutorials/02_homogenisation.py:225-231)
    a = 1/(K.mean()+4./3*Gm.mean())
    b = 1./(2*Gm.mean())
    AHms, kitms, _, _ = solve_all(A, a*hG1h+b*hG1s, N, D, 1e-8)
    out['tut02_ms_ab'] = np.array([a, b])
    out['tut02_ms_AH'] = AHms
    out['tut02_ms_kit'] = kitms
    print('  tutorial 02: AH00 = %.15g (kit %s), MS kit %s' % (AH[0, 0], kits, kitms))
    # tutorial 04: exact integration on the doubled grid (Ga), 2-D scalar N=25 -> Nbar=49
    dim = 2
    N = 25*np.ones(dim, dtype=np.int32)
    P = 5*np.ones(dim, dtype=np.int32)
    pbmat = {'Y': np.ones(dim), 'inclusions': ['square', 'otherwise'], 'positions': [np.zeros(dim), ''],
             'params': [0.6*np.ones(dim), ''], 'vals': [11*np.eye(dim), np.eye(dim)], 'order': 1, 'P': P}
    Nbar = 2*N-1
    A = quiet(Material(pbmat).get_A_Ga, Nbar=Nbar, primaldual='primal')
    _, hG1N, _ = proj.scalar(N, np.ones(dim), NyqNul=True, tensor=True)
    hG1N = hG1N.enlarge(Nbar)
    AH, kits, nres, sols = solve_all(A, hG1N, Nbar, dim, 1e-8)
    out['tut04_A'] = A.val
    out['tut04_AH'] = AH
    out['tut04_kit'] = kits
    out['tut04_normres'] = nres
    print('  tutorial 04: AH00 = %.15g (kit %s)' % (AH[0, 0], kits))
    np.savez_compressed(os.path.join(OUT, 'configs.npz'), **out)
    print('configs.npz: %d arrays' % len(out))


# ----------------------------------------------------------------------------- 5. round 2: BASELINE configs at CPU-feasible sizes
def gen_round2(big=True):
    """C1b (2-D scalar 31x31 GaNi + Ga 61x61 through the reference's Problem driver), a C2-shaped problem (3-D scalar
    'cube' 0.7, kind 'Ga', order None, N=16->31 and 32->63, primal + dual bounds) and C3 (random two-phase
    elasticity) at 64^3 (primal + dual) and 128^3 (primal) -- SURVEY App. C recipes.  Written to round2.npz so the
    round-1 fixture files stay bit-identical."""
    out = {}
    # ---- C1b
    for kind in ('GaNi', 'Ga'):
        N = 31*np.ones(2, dtype=np.int32)
        mat = {'inclusions': ['square', 'otherwise'], 'positions': [np.zeros(2), ''], 'params': [0.6*np.ones(2), ''],
               'vals': [11*np.eye(2), 1.*np.eye(2)], 'Y': np.ones(2), 'order': None}
        pb = {'name': 'p', 'physics': 'scalar', 'material': mat,
              'solve': {'kind': kind, 'N': N, 'primaldual': ['primal', 'dual']},
              'postprocess': [{'kind': 'GaNi'}, {'kind': 'Ga', 'order': None}] if kind == 'GaNi' else [{'kind': 'Ga', 'order': None}],
              'solver': {'kind': 'CG', 'tol': 1e-8, 'maxiter': 1e3}}
        prob = quiet(Problem, pb, None)
        quiet(prob.calculate)
        M = Material(mat)
        Nbar = 2*N-1
        for pd in ('primal', 'dual'):
            tag = 'c1b_%s_%s' % (kind, pd)
            A = quiet(M.get_A_GaNi, N, pd) if kind == 'GaNi' else quiet(M.get_A_Ga, Nbar=Nbar, primaldual=pd, order=None)
            out[tag+'_A'] = A.val
            out[tag+'_kit'] = np.array([r['info']['kit'] for r in prob.output['res_'+pd]])
            out[tag+'_normres'] = np.array([r['info']['norm_res'] for r in prob.output['res_'+pd]])
            out[tag+'_AGa'] = quiet(M.get_A_Ga, Nbar=Nbar, primaldual=pd, order=None).val
            for name, v in prob.output['mat_'+pd].items():
                out[tag+'_'+name] = v
            print('  C1b %s %s: kit %s  %s' % (kind, pd, out[tag+'_kit'],
                                             {k: float(v[0, 0]) for k, v in prob.output['mat_'+pd].items()}))
    # ---- C2-shaped: exact integration of a cube inclusion, primal + dual => upper / lower bounds
    for n in (16, 32):
        N = n*np.ones(3, dtype=np.int32)
        Nbar = 2*N-1
        mat = {'inclusions': ['cube', 'otherwise'], 'positions': [np.zeros(3), ''], 'params': [0.7*np.ones(3), ''],
               'vals': [11*np.eye(3), 1.*np.eye(3)], 'Y': np.ones(3), 'order': None}
        pb = {'name': 'p', 'physics': 'scalar', 'material': mat,
              'solve': {'kind': 'Ga', 'N': N, 'primaldual': ['primal', 'dual']},
              'postprocess': [{'kind': 'Ga', 'order': None}],
              'solver': {'kind': 'CG', 'tol': 1e-6, 'maxiter': 1e3}}
        prob = quiet(Problem, pb, None)
        quiet(prob.calculate)
        M = Material(mat)
        for pd in ('primal', 'dual'):
            tag = 'c2_n%d_%s' % (n, pd)
            A = quiet(M.get_A_Ga, Nbar=Nbar, primaldual=pd, order=None).val
            offd = max(np.abs(A[i, j]).max() for i in range(3) for j in range(3) if i != j)
            iso = max(np.abs(A[i, i]-A[0, 0]).max() for i in range(3))
            assert offd == 0. and iso == 0., (offd, iso)
            out[tag+'_a'] = A[0, 0]                       # A = a(x) I exactly: one scalar field travels
            out[tag+'_kit'] = np.array([r['info']['kit'] for r in prob.output['res_'+pd]])
            out[tag+'_normres'] = np.array([r['info']['norm_res'] for r in prob.output['res_'+pd]])
            out[tag+'_AH'] = prob.output['mat_'+pd]['AH_Ga_'+pd]
            print('  C2 n=%d %s: kit %s AH00 %.15g' % (n, pd, out[tag+'_kit'], out[tag+'_AH'][0, 0]))
    # ---- C3 at 64^3 and 128^3
    Cm = ElasticTensor(bulk=1, mu=1).mandel
    Ci = ElasticTensor(bulk=10, mu=5).mandel
    for n, pds in ((64, ('primal', 'dual')), (128, ('primal',))) if big else ():
        N = np.array([n, n, n])
        rng = np.random.default_rng(20240901)
        phase = (rng.random((n, n, n)) < 0.3).astype(float)
        Gs = elasticity_anyN(N, np.ones(3))
        for pd in pds:
            cm, ci = (Cm, Ci) if pd == 'primal' else (np.linalg.inv(Cm), np.linalg.inv(Ci))
            A = Tensor(name='A', val=np.einsum('ij,...->ij...', cm, 1-phase)+np.einsum('ij,...->ij...', ci, phase),
                       order=2, N=N, multype=21)
            G = Gs[1]+Gs[2] if pd == 'primal' else Gs[3]+Gs[4]
            AH, kits, nres, _ = solve_all(A, G, N, 6, 1e-6)
            if pd == 'dual':
                AH = np.linalg.inv(AH)
            out['c3_n%d_%s_AH' % (n, pd)] = AH
            out['c3_n%d_%s_kit' % (n, pd)] = kits
            out['c3_n%d_%s_normres' % (n, pd)] = nres
            print('  C3 n=%d %s: kit %s AH00 %.15g AH01 %.15g AH33 %.15g' % (n, pd, kits, AH[0, 0], AH[0, 1], AH[3, 3]),
                  flush=True)
    np.savez_compressed(os.path.join(OUT, 'round2.npz'), **out)
    print('round2.npz: %d arrays' % len(out))


# ----------------------------------------------------------------------------- 6. potential (displacement-based) formulation
def gen_potential():
    """ffthompy/tensorsLowRank/homogenisation.py:13-128 (homog_Ga_full, homog_Ga_full_potential,
    homog_GaNi_full_potential) run UNMODIFIED on the SURVEY App. C material (square 0.6, 10:1, order 0, P = 5); the
    module imports ttpy, which is absent here and unused on this path, so an empty stand-in is registered for it."""
    import types
    for m in ('tt', 'tt.core', 'tt.core.vector'):
        sys.modules.setdefault(m, types.ModuleType(m))
    sys.modules['tt.core.vector'].vector = type('vector', (), {})
    from ffthompy import Struct
    from ffthompy.tensorsLowRank.homogenisation import homog_Ga_full_potential, homog_Ga_full, homog_GaNi_full_potential
    out = {}
    for dim, n in ((2, 5), (3, 5), (2, 15), (2, 16), (3, 9)):
        N = n*np.ones(dim, dtype=int)
        Nbar = 2*N-1
        mat = {'inclusions': ['square', 'otherwise'], 'positions': [np.zeros(dim), ''], 'params': [0.6*np.ones(dim), ''],
               'vals': [10*np.eye(dim), 1.*np.eye(dim)], 'Y': np.ones(dim), 'order': 0, 'P': 5*np.ones(dim, dtype=int)}
        Aga = quiet(Material(mat).get_A_Ga, Nbar, 'primal')
        Agani = quiet(Material(mat).get_A_GaNi, N, 'primal')
        pars = Struct(dim=dim, N=N, Y=np.ones(dim), solver=dict(tol=1e-8, maxiter=200))
        rP = quiet(homog_Ga_full_potential, Aga, pars)
        rF = quiet(homog_Ga_full, Aga, pars)
        rG = quiet(homog_GaNi_full_potential, Agani, Aga, pars)
        rG0 = quiet(homog_GaNi_full_potential, Agani, None, pars)
        tag = 'pot_d%d_n%d' % (dim, n)
        out[tag+'_Aga'], out[tag+'_Agani'] = Aga.val, Agani.val
        out[tag+'_AH_Ga_potential'], out[tag+'_kit_Ga_potential'] = rP.AH, rP.info['kit']
        out[tag+'_normres_Ga_potential'] = rP.info['norm_res']
        out[tag+'_Fu_Ga_potential'] = rP.Fu.val
        out[tag+'_e_Ga_potential'] = rP.e.val
        out[tag+'_AH_Ga_gradient'] = rF.AH
        out[tag+'_AH_GaNi_potential_Ga'], out[tag+'_kit_GaNi_potential'] = rG.AH, rG.info['kit']
        out[tag+'_AH_GaNi_potential_GaNi'] = rG0.AH
        print('  potential dim=%d N=%d: Ga %.15g (%d its; gradient-field %.15g), GaNi->Ga %.15g (%d its), GaNi %.15g'
              % (dim, n, rP.AH, rP.info['kit'], rF.AH, rG.AH, rG.info['kit'], rG0.AH))
    np.savez_compressed(os.path.join(OUT, 'potential.npz'), **out)
    print('potential.npz: %d arrays' % len(out))


# ----------------------------------------------------------------------------- 7. material coefficients
def _conf_json(conf):
    import json
    out = {}
    for k, v in conf.items():
        if k in ('inclusions',):
            out[k] = list(v)
        elif k in ('positions', 'params', 'vals'):
            out[k] = [x if isinstance(x, str) else np.asarray(x).tolist() for x in v]
        elif k in ('Y', 'P'):
            out[k] = np.asarray(v).tolist()
        else:
            out[k] = v
    return json.dumps(out)


def gen_materials():
    """Material.get_A_GaNi / get_A_Ga (ffthompy/materials.py:54-124) of the UNMODIFIED reference for every material of
    the example input files plus shifted / anisotropic / 3-D ball / even-grid variants; the configuration travels as
    JSON next to the arrays."""
    os.chdir(REF)
    out, names = {}, []
    confs = []
    for f in ['examples/scalar/scalar_2d.py', 'examples/scalar/scalar_3d.py', 'examples/elasticity/linelas_3d.py']:
        conf = quiet(import_file, f)
        for mname, m in conf.materials.items():
            if 'fun' in m or 'inclusions' not in m:
                continue
            confs.append((os.path.basename(f).split('.')[0]+'_'+mname, m, np.array(conf.N)))
    d2, d3 = np.ones(2), np.ones(3)
    confs += [
        ('shifted_square', {'inclusions': ['square', 'otherwise'], 'positions': [np.array([0.2, -0.15]), ''],
                            'params': [np.array([0.5, 0.3]), ''], 'vals': [np.array([[5., 1.], [1., 3.]]), np.eye(2)],
                            'Y': d2, 'order': None}, np.array([7, 6])),
        ('two_inclusions', {'inclusions': ['ball', 'square', 'otherwise'], 'positions': [np.array([0.25, 0.25]), np.array([-0.25, -0.2]), ''],
                            'params': [0.3, np.array([0.3, 0.2]), ''], 'vals': [7*np.eye(2), 3*np.eye(2), np.eye(2)],
                            'Y': d2, 'order': None}, np.array([9, 10])),
        ('ball_3d', {'inclusions': ['ball', 'otherwise'], 'positions': [np.zeros(3), ''], 'params': [0.7, ''],
                     'vals': [11*np.eye(3), np.eye(3)], 'Y': d3, 'order': None}, np.array([6, 5, 7])),
        ('pyramid_3d', {'inclusions': ['pyramid', 'all'], 'positions': [np.zeros(3), ''], 'params': [0.8*np.ones(3), ''],
                        'vals': [10.*np.eye(3), np.eye(3)], 'Y': d3, 'order': None}, np.array([5, 6, 5])),
        ('rect_cell', {'inclusions': ['square', 'otherwise'], 'positions': [np.zeros(2), ''], 'params': [np.array([1.0, 0.4]), ''],
                       'vals': [11*np.eye(2), np.eye(2)], 'Y': np.array([2., 1.]), 'order': None}, np.array([8, 5])),
    ]
    for tag, m, N in confs:
        m = dict(m)
        mat = Material(m)
        Nbar = 2*N-1
        names.append(tag)
        out[tag+'_conf'] = np.array(_conf_json(m))
        out[tag+'_N'] = N
        for pd in ('primal', 'dual'):
            out['%s_GaNi_%s' % (tag, pd)] = quiet(mat.get_A_GaNi, N, pd).val
            out['%s_Ga_None_%s' % (tag, pd)] = quiet(mat.get_A_Ga, Nbar, pd, None).val
            for order in (0, 1):
                for Pn, P in (('N', N), ('2N', 2*N), ('3', 3*np.ones(N.size, dtype=int))):
                    out['%s_Ga_o%d_P%s_%s' % (tag, order, Pn, pd)] = quiet(mat.get_A_Ga, Nbar, pd, order, P).val
    out['names'] = np.array(names)
    np.savez_compressed(os.path.join(OUT, 'materials.npz'), **out)
    print('materials.npz: %d arrays, %d materials' % (len(out), len(names)))


if __name__ == '__main__':
    if '--materials' in sys.argv:
        gen_materials()
        sys.exit(0)
    if '--potential' in sys.argv:
        gen_potential()
        sys.exit(0)
    if '--round2' in sys.argv:          # leaves the round-1 files untouched
        gen_round2(big='--small' not in sys.argv)
        sys.exit(0)
    gen_projections()
    gen_tensors()
    gen_configs()
    gen_examples()
    gen_round2()
    gen_potential()
    gen_materials()
    os.system('ls -la %s' % OUT)
