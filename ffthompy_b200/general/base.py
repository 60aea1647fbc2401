"""Host-side glue mirrored from ffthompy/general/base.py (Timer, PrintControl,
Representation) so that callers of the solve loop find the same helpers."""
import os
import sys
import time

import numpy as np


class Representation():
    def _repr(self, keys, skip=4*' '):
        """general/base.py:29-40"""
        ss = "Class : {0} \n".format(self.__class__.__name__)
        nstr = np.array([len(key) for key in keys]).max()
        for key in keys:
            attr = getattr(self, key)
            if callable(attr):
                ss += '{0}{1}{3} = {2}\n'.format(skip, key, str(attr()), (nstr-len(key))*' ')
            else:
                ss += '{0}{1}{3} = {2}\n'.format(skip, key, str(attr), (nstr-len(key))*' ')
        return ss


class PrintControl():
    """general/base.py:42-61"""
    flag = True

    def __init__(self, flag=True):
        self.flag = True

    def activate(self):
        self.flag = True

    def deactivate(self):
        self.flag = False

    def disable(self):
        if self.flag:
            sys.stdout = open(os.devnull, 'w')

    def enable(self):
        if self.flag:
            sys.stdout.close()
            sys.stdout = sys.__stdout__


class Timer():
    """general/base.py:63-81 (time.clock no longer exists; perf_counter takes its slot)"""

    def __init__(self, name='time', start=True):
        self.name = name
        if start:
            self.start()

    def start(self):
        self.vals = []
        self.ttin = [time.process_time(), time.perf_counter(), time.time()]

    def measure(self, print_time=True):
        self.vals.append([time.process_time()-self.ttin[0],
                          time.perf_counter()-self.ttin[1],
                          time.time()-self.ttin[2]])
        if print_time:
            print(self)

    def __repr__(self):
        return 'time (%s): %s' % (self.name, str(self.vals))
