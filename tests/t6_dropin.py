"""T6 (SURVEY section 4 / VERDICT round 1 task 1): the reference's OWN integration suite — the 12 example problems of
run_unittests.py:29-67 through `Problem.calculate()`, plus tutorials 02-04 — executed UNMODIFIED from the copy under
baseline/_ref with `ffthompy_b200.install()` active, i.e. ffthompy.tensors / .projections / .general.solver /
.general.solver_pp / .trigpol / .materials / .postprocess are this package's device-backed modules while
applications.py and problem.py (the callers) are the reference's files.  Checked against the reference's pickled goldens
(test_results/python3/*): homogenised matrices to 1e-9 (the reference's own bar, run_unittests.py:62) and CG
iteration counts equal.

    python tests/t6_dropin.py [--log profiles/r02_t6_dropin.log]

Needs a GPU and the reference copy (made by __graft_entry__.build() where /root/reference exists)."""
import contextlib
import io
import os
import pickle
import sys
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
REF = os.environ.get('FFTHOMPY_REFERENCE', os.path.join(ROOT, 'baseline', '_ref'))
EXAMPLES = ['examples/scalar/scalar_2d.py', 'examples/scalar/scalar_3d.py', 'examples/scalar/from_file.py',
            'examples/elasticity/linelas_3d.py']
TUTORIALS = ['tutorials/02_homogenisation.py', 'tutorials/03_exact_integration_simple.py',
             'tutorials/04_exact_integration_fast.py']


def available():
    return os.path.isdir(os.path.join(REF, 'ffthompy'))


def load_goldens():
    """unpickle the reference's goldens with the PURE reference modules, keep plain numbers, then forget the modules"""
    sys.path.insert(0, os.path.join(ROOT, 'oracle'))
    os.environ['FFTHOMPY_REFERENCE'] = REF
    import _refshim
    _refshim.REFERENCE = REF
    _refshim.install()
    gold = {}
    for f in EXAMPLES:
        sys.path.insert(0, os.path.join(REF, os.path.dirname(f)))
    for name in sorted(os.listdir(os.path.join(REF, 'test_results', 'python3'))):
        with open(os.path.join(REF, 'test_results', 'python3', name), 'rb') as fh:
            res = pickle.load(fh)
        g = {}
        for pd in ('primal', 'dual'):
            if 'mat_'+pd in res:
                g['mat_'+pd] = {k: np.array(v) for k, v in res['mat_'+pd].items()}
                g['kit_'+pd] = [int(r['info']['kit']) for r in res['res_'+pd]]
        gold[name] = g
    for m in [m for m in sys.modules if m == 'ffthompy' or m.startswith('ffthompy.')]:
        del sys.modules[m]
    return gold


def run(log=print):
    gold = load_goldens()
    sys.path.insert(0, ROOT)
    import ffthompy_b200
    from ffthompy_b200 import device
    device.init(0)
    ffthompy_b200.install()
    import ffthompy.tensors
    assert ffthompy.tensors.Tensor.__module__.startswith('ffthompy_b200'), 'splice not active'
    from ffthompy.problem import Problem, import_file
    import ffthompy.applications as apps
    assert apps.Tensor.__module__.startswith('ffthompy_b200') and apps.linear_solver.__module__.startswith('ffthompy_b200')
    assert apps.Material.__module__.startswith('ffthompy_b200') and apps.postprocess.__module__.startswith('ffthompy_b200')
    assert apps.scalar.__module__ == 'ffthompy.applications'     # the reference's own callers
    os.chdir(REF)
    n0 = device.launch_count()
    worst, fails = 0., []
    for f in EXAMPLES:
        with contextlib.redirect_stdout(io.StringIO()):
            conf = import_file(f)
        for cp in conf.problems:
            tag = os.path.basename(f).split('.')[0]+'_'+cp['name']
            t0 = time.perf_counter()
            with contextlib.redirect_stdout(io.StringIO()):
                prob = Problem(cp, conf)
                prob.calculate()
            dt = time.perf_counter()-t0
            g = gold[tag]
            for pd in prob.solve['primaldual']:
                kits = [int(r['info']['kit']) for r in prob.output['res_'+pd]]
                ok_k = kits == g['kit_'+pd]
                dmax = 0.
                for kw, ref in g['mat_'+pd].items():
                    dmax = max(dmax, float(np.abs(prob.output['mat_'+pd][kw]-ref).max()))
                worst = max(worst, dmax)
                ok = ok_k and dmax < 1e-9
                if not ok:
                    fails.append((tag, pd))
                log('%-24s %-6s kit %-28s (golden %s)  max|dA_H| %.2e  %s' % (tag, pd, kits, 'equal' if ok_k else g['kit_'+pd],
                                                                             dmax, 'ok' if ok else 'FAIL'))
            log('%-24s Problem.calculate(): %.2f s' % (tag, dt))
    # tutorials: executed verbatim; known answers of SURVEY App. C
    expect = {'tutorials/02_homogenisation.py': 3.92394827320454, 'tutorials/03_exact_integration_simple.py': 1.865763624318734,
              'tutorials/04_exact_integration_fast.py': 2.464008025892713}
    for filen in TUTORIALS:
        out = io.StringIO()
        t0 = time.perf_counter()
        scope = {'__name__': 'test'}
        with contextlib.redirect_stdout(out):
            exec(compile(open(filen).read(), filen, 'exec'), scope)
        dt = time.perf_counter()-t0
        vals = []
        for line in out.getvalue().splitlines():
            for tok in line.replace('=', ' ').replace('[', ' ').replace(']', ' ').replace(',', ' ').split():
                try:
                    vals.append(float(tok))
                except ValueError:
                    pass
        hit = min((abs(v-expect[filen]) for v in vals), default=np.inf)
        ok = hit < 1e-9
        if not ok:
            fails.append((filen, 'value'))
        log('%-44s printed value closest to %.15g: off by %.2e  %s  (%.2f s)' % (filen, expect[filen], hit,
                                                                                'ok' if ok else 'FAIL', dt))
    log('kernels of libffthom_b200.so launched: %d; worst |dA_H| over all problems %.2e; failures: %s'
        % (device.launch_count()-n0, worst, fails or 'none'))
    return fails


if __name__ == '__main__':
    if not available():
        print('reference copy not found under %s (run __graft_entry__.build() where /root/reference exists)' % REF)
        sys.exit(2)
    lines = []

    def log(s):
        print(s, flush=True)
        lines.append(s)
    fails = run(log)
    if '--log' in sys.argv:
        path = os.path.join(ROOT, sys.argv[sys.argv.index('--log')+1])
        with open(path, 'w') as fh:
            fh.write('# T6: unmodified reference callers (baseline/_ref) through ffthompy_b200.install() on cuda:0\n')
            fh.write('\n'.join(lines)+'\n')
    sys.exit(1 if fails else 0)
