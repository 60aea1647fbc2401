import ast
import os
import sys

import numpy as np
import pytest

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), '..'))
for p in (ROOT, os.path.join(ROOT, 'oracle'), os.path.dirname(__file__)):
    if p not in sys.path:
        sys.path.insert(0, p)

GOLDEN = os.path.join(os.path.dirname(__file__), 'golden')


def pytest_configure(config):
    config.addinivalue_line('markers', 'gpu: needs a CUDA device (run on the B200 box with -m gpu)')


class Golden(object):
    """lazy access to tests/golden/*.npz (generated from the unmodified reference by oracle/make_golden.py)"""
    _cache = {}

    def __getitem__(self, name):
        if name not in self._cache:
            self._cache[name] = np.load(os.path.join(GOLDEN, name+'.npz'), allow_pickle=False)
        return self._cache[name]

    def example_meta(self):
        return [ast.literal_eval(str(m)) for m in self['examples']['meta']]


@pytest.fixture(scope='session')
def golden():
    return Golden()


def example_tags():
    g = Golden()
    return [m[0] for m in g.example_meta()]
