"""`Material.get_A_GaNi / get_A_Ga` on the device against arrays written by the UNMODIFIED
reference (ffthompy/materials.py:54-124 via oracle/make_golden.py --materials) for every inclusion material of the
example input files plus shifted / anisotropic / 3-D ball / pyramid / rectangular-cell variants.

GaNi (nodal) values are sums of phase matrices times 0/1 (or tent-function) topologies evaluated with the reference's
own comparisons: BIT-IDENTICAL.  Ga values pass through the device FFT: 1e-12 relative."""
import json

import numpy as np
import pytest

from conftest import Golden

pytestmark = pytest.mark.gpu


def _names():
    return [str(n) for n in Golden()['materials']['names']]


def _conf(g, tag):
    c = json.loads(str(g[tag+'_conf']))
    for k in ('positions', 'params', 'vals'):
        c[k] = [x if isinstance(x, str) else (np.array(x) if np.ndim(x) else x) for x in c[k]]
    c['Y'] = np.array(c['Y'])
    if 'P' in c:
        c['P'] = np.array(c['P'])
    return c


@pytest.fixture(scope='module', autouse=True)
def _device():
    from ffthompy_b200 import device
    device.init(0)


@pytest.mark.parametrize('tag', _names())
def test_material_coefficients_match_the_reference(golden, tag):
    from ffthompy_b200.materials import Material
    g = golden['materials']
    N = np.array(g[tag+'_N'])
    Nbar = 2*N-1
    mat = Material(_conf(g, tag))
    for pd in ('primal', 'dual'):
        A = mat.get_A_GaNi(N, pd)
        ref = g['%s_GaNi_%s' % (tag, pd)]
        assert A.name == 'A_GaNi' or pd == 'dual'
        assert A.val.shape == ref.shape and A.origin == 0
        if pd == 'primal':
            assert np.array_equal(A.val, ref), np.abs(A.val-ref).max()
        else:
            assert np.abs(A.val-ref).max() <= 1e-13*np.abs(ref).max()
        A = mat.get_A_Ga(Nbar, pd, None)
        ref = g['%s_Ga_None_%s' % (tag, pd)]
        assert A.name == 'A_Ga' and A.val.shape == ref.shape
        assert np.abs(A.val-ref).max() <= 1e-12*np.abs(ref).max(), ('order None', pd)
        for order in (0, 1):
            for Pn, P in (('N', N), ('2N', 2*N), ('3', 3*np.ones(N.size, dtype=int))):
                A = mat.get_A_Ga(Nbar, pd, order, P)
                ref = g['%s_Ga_o%d_P%s_%s' % (tag, order, Pn, pd)]
                assert np.abs(A.val-ref).max() <= 1e-12*np.abs(ref).max(), (order, Pn, pd)


def test_material_checks_and_topologies():
    from ffthompy_b200.materials import Material
    from ffthompy_b200.trigpol import Grid
    with pytest.raises(ValueError):
        Material({'inclusions': ['square'], 'positions': [np.zeros(2)], 'params': [np.ones(2)], 'vals': [np.eye(2)]})
    with pytest.raises(ValueError):
        Material({'inclusions': ['square', 'otherwise'], 'positions': [np.zeros(2)], 'params': [np.ones(2), ''],
                  'vals': [np.eye(2), np.eye(2)], 'Y': np.ones(2)})
    with pytest.raises(ValueError):
        Material({'inclusions': ['square', 'otherwise'], 'positions': [np.zeros(2), ''], 'params': [2*np.ones(2), ''],
                  'vals': [np.eye(2), np.eye(2)], 'Y': np.ones(2)})
    m = Material({'inclusions': ['square', 'square', 'otherwise'], 'positions': [np.zeros(2), np.array([0.1, 0.1]), ''],
                  'params': [0.6*np.ones(2), 0.6*np.ones(2), ''], 'vals': [np.eye(2)]*3, 'Y': np.ones(2)})
    with pytest.raises(NotImplementedError):        # overlapping inclusions (materials.py:296-297)
        m.get_A_GaNi(np.array([9, 9]))
    m = Material({'inclusions': ['circle', 'otherwise'], 'positions': [np.zeros(2), ''], 'params': [0.5, ''],
                  'vals': [2*np.eye(2), np.eye(2)], 'Y': np.ones(2)})
    topo = m.get_topologies(Grid.get_coordinates(np.array([15, 15]), np.ones(2)))
    assert abs(topo[0].sum()/225-np.pi*0.25**2) < 0.02 and np.all(topo[0]+topo[1] == 1)
    f = Material({'fun': lambda c: np.einsum('ij,...->ij...', np.eye(2), 1+c[0]**2), 'Y': np.ones(2)})
    A = f.get_A_GaNi(np.array([5, 5]))
    assert A.val.shape == (2, 2, 5, 5)
