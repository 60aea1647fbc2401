"""Small host-side helpers with the interface of ffthompy/general/base.py (Timer, PrintControl, Representation):
callers of the solve loop (`linear_solver` records `info['time']`, `Tensor.__repr__`) find the names they expect.
Written for this package; only the call signatures follow the reference."""
import contextlib
import io
import sys
import time


class Representation(object):
    """mixin: a class lists attribute names, `_repr` renders them one per line (ffthompy/general/base.py:28-40)"""

    def _repr(self, keys, skip='    '):
        width = max((len(k) for k in keys), default=0)
        rows = ['Class : %s ' % type(self).__name__]
        for k in keys:
            v = getattr(self, k)
            rows.append('%s%s = %s' % (skip, k.ljust(width), v() if callable(v) else v))
        return '\n'.join(rows)+'\n'


class PrintControl(object):
    """silence / restore stdout around noisy calls (ffthompy/general/base.py:42-61: disable() ... enable())"""

    def __init__(self, flag=True):
        self.flag = True
        self._saved = None

    def activate(self):
        self.flag = True

    def deactivate(self):
        self.flag = False

    def disable(self):
        if self.flag and self._saved is None:
            self._saved = sys.stdout
            sys.stdout = io.StringIO()

    def enable(self):
        if self.flag and self._saved is not None:
            sys.stdout = self._saved
            self._saved = None

    @contextlib.contextmanager
    def quiet(self):
        self.disable()
        try:
            yield
        finally:
            self.enable()


class Timer(object):
    """stop-watch with the (cpu, wall, wall) triple per measurement that `info['time']` carries
    (ffthompy/general/base.py:63-81; the removed time.clock is replaced by process_time / perf_counter)"""
    _clocks = (time.process_time, time.perf_counter, time.time)

    def __init__(self, name='time', start=True):
        self.name = name
        self.vals = []
        if start:
            self.start()

    def start(self):
        self.vals = []
        self.ttin = [c() for c in self._clocks]

    def measure(self, print_time=True):
        self.vals.append([c()-t0 for c, t0 in zip(self._clocks, self.ttin)])
        if print_time:
            print(self)

    def __repr__(self):
        return 'time (%s): %s' % (self.name, self.vals)
