"""ffthompy_b200 — B200-native (sm_100a) implementation of FFTHomPy's Fourier-Galerkin solve
loop behind FFTHomPy's own operator API.

    from ffthompy_b200.tensors import Tensor, DFT, Operator, grad, div, potential
    import ffthompy_b200.projections as proj
    from ffthompy_b200.general.solver import linear_solver
    from ffthompy_b200 import trigpol, matvecs
    from ffthompy_b200.materials import Material              # get_A_GaNi / get_A_Ga on the device
    from ffthompy_b200.postprocess import postprocess         # one-pass A_H assembly, bounds driver
    from ffthompy_b200 import homogenisation                  # potential (displacement-based) formulation

mirror ffthompy.tensors / ffthompy.projections / ffthompy.general.solver / ffthompy.trigpol /
ffthompy.matvecs (same names, arguments and error behaviour); `install()` splices them into an
importable reference tree so that ffthompy.applications and the tutorials run on the GPU
unmodified.  All arithmetic is hand-written CUDA behind the C ABI of
libffthom_b200.so (include/ffthom_b200.h); there is no CPU fallback.
"""
import sys

__version__ = '0.1.0'

from .general.base import PrintControl, Timer  # noqa: F401,E402


def install(reference_package='ffthompy'):
    """Register this package's modules under the reference's module names, so that
    `import ffthompy.applications` (and anything else in the reference tree that imports
    ffthompy.tensors / .projections / .general.solver / .trigpol) binds the B200 objects.
    Call before importing the reference's callers."""
    import importlib
    ref = importlib.import_module(reference_package)  # the reference tree must be importable
    from . import tensors, projections, trigpol, postprocess, materials
    from .tensors import objects, operators, projection, fft
    from .general import solver, solver_pp
    mapping = {
        'tensors': tensors, 'tensors.objects': objects, 'tensors.operators': operators,
        'tensors.projection': projection, 'tensors.fft': fft, 'projections': projections,
        'general.solver': solver, 'general.solver_pp': solver_pp, 'trigpol': trigpol,
        'postprocess': postprocess, 'materials': materials,
    }
    for name, mod in mapping.items():
        sys.modules[reference_package+'.'+name] = mod
    ref.tensors = tensors
    ref.projections = projections
    ref.trigpol = trigpol
    return mapping
