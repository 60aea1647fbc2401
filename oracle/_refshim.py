"""Compatibility shim that lets the UNMODIFIED reference (/root/reference, FFTHomPy) import
under Python 3.12 / NumPy 2 / SciPy 1.18.  Test infrastructure only:
used by oracle/make_golden.py and by tests that cross-check the oracle when the reference
tree is present (it does not exist on the GPU box)."""
import collections
import collections.abc
import os
import sys
import time
import warnings

import numpy as np
import scipy as sp

REFERENCE = os.environ.get('FFTHOMPY_REFERENCE', '/root/reference')


def available():
    return os.path.isdir(os.path.join(REFERENCE, 'ffthompy'))


def install():
    """Patch removed aliases and put the reference on sys.path.  Idempotent."""
    if not available():
        raise RuntimeError('reference tree not found at %s' % REFERENCE)
    warnings.filterwarnings('ignore')
    for n, t in (('int', int), ('float', float), ('complex', complex), ('bool', bool)):
        if not hasattr(np, n):
            setattr(np, n, t)
    if not hasattr(time, 'clock'):
        time.clock = time.perf_counter
    if not hasattr(collections, 'Callable'):
        collections.Callable = collections.abc.Callable
    if not hasattr(sp, 'setdiff1d'):
        sp.setdiff1d = np.setdiff1d
    if REFERENCE not in sys.path:
        sys.path.insert(0, REFERENCE)
    return REFERENCE
