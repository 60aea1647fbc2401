"""the potential (displacement-based) formulation with per-iteration grad -> enlarge -> A ->
project -> div and the 1/|2 pi xi|-preconditioned Fourier-space CG, against fixtures written by the UNMODIFIED
reference (ffthompy/tensorsLowRank/homogenisation.py:41-128 via oracle/make_golden.py --potential).
A_H 1e-10 relative (and equal to the gradient-field Ga solve of the same coefficients on odd grids, 1e-12),
CG iteration counts equal, potential and minimiser 1e-9."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

CASES = [(2, 5), (3, 5), (2, 15), (2, 16), (3, 9)]


@pytest.fixture(scope='module', autouse=True)
def _device():
    from ffthompy_b200 import device
    device.init(0)


@pytest.mark.parametrize('dim,n', CASES)
def test_potential_formulation_matches_the_reference(golden, dim, n):
    from ffthompy_b200.tensors import Tensor
    from ffthompy_b200.homogenisation import Struct, homog_Ga_full_potential, homog_Ga_full, homog_GaNi_full_potential
    g = golden['potential']
    tag = 'pot_d%d_n%d' % (dim, n)
    N = n*np.ones(dim, dtype=int)
    Nbar = 2*N-1
    Aga = Tensor(name='Aga', val=g[tag+'_Aga'].copy(), order=2, N=Nbar, multype=21)
    Agani = Tensor(name='Agani', val=g[tag+'_Agani'].copy(), order=2, N=N, multype=21)
    pars = Struct(dim=dim, N=N, Y=np.ones(dim), solver=dict(tol=1e-8, maxiter=200))
    rP = homog_Ga_full_potential(Aga, pars)
    assert rP.info['kit'] == int(g[tag+'_kit_Ga_potential'])
    ref = float(g[tag+'_AH_Ga_potential'])
    assert abs(rP.AH-ref) <= 1e-10*abs(ref)
    assert abs(rP.Fu.mean()) < 1e-12
    assert np.abs(rP.Fu.val-g[tag+'_Fu_Ga_potential']).max() < 1e-9
    assert np.abs(rP.e.val-g[tag+'_e_Ga_potential']).max() < 1e-9
    rF = homog_Ga_full(Aga, pars)
    assert abs(rF.AH-float(g[tag+'_AH_Ga_gradient'])) <= 1e-10*abs(ref)
    if n % 2:   # both formulations minimise over the same space on odd grids
        assert abs(rF.AH-rP.AH) <= 1e-12*abs(ref)
    rG = homog_GaNi_full_potential(Agani, Aga, pars)
    assert rG.info['kit'] == int(g[tag+'_kit_GaNi_potential'])
    assert abs(rG.AH-float(g[tag+'_AH_GaNi_potential_Ga'])) <= 1e-10*abs(rG.AH)
    rG0 = homog_GaNi_full_potential(Agani, None, pars)
    assert abs(rG0.AH-float(g[tag+'_AH_GaNi_potential_GaNi'])) <= 1e-10*abs(rG0.AH)
