#!/usr/bin/env python
"""Shared-memory bank-conflict model and padding search for the SoA lines of the 512-length last-axis kernels
(k_fwd_last_reg3 / k_inv_last_reg3, csrc/fh_reg3.cuh: `LinePad<512>`).  Runs without a GPU.

Model (checked against ncu: the round-1/2 layout `p + p/8` gives 40 wavefronts per element and direction against an
ideal 20, ncu measured 965 M wavefronts of which 491 M conflicts = 2.04 x): 8-byte accesses, sixteen 8-byte banks per
128-byte wavefront, a warp is served as two half warps, each costing max-multiplicity-of-a-bank wavefronts.

Patterns of one line of N = 512 = 8 x 8 x 8 (64 threads per line, lane u = butterfly index):
  ph0    stores of sigma in phase 0            position 2 u (+1)
  p1r/w  pass 1 reads u + 64 r, writes 64 q + u
  p2     pass 2  64 q + jp + 8 r   (lane -> (q, jp) in either order)
  p3     pass 3  64 q + 8 q2 + jp  (lane -> (q, q2) in either order)
  fin    R2C / C2R separation: position of frequency k and of N - k for 32 consecutive k
"""
import itertools

import numpy as np

N = 512


def wf(positions):
    p = np.array(positions)
    return sum(np.bincount(p[16*h:16*h+16] % 16, minlength=16).max() for h in (0, 1))


def pos_of_freq(k):
    q, r1 = k % 8, k//8
    return q*64+(r1 % 8)*8+r1//8


def avg(pad, f):
    tot = 0
    for half in (0, 1):
        for c in range(8):
            tot += wf([pad(f(u, c)) for u in range(32*half, 32*half+32)])
    return tot/16.


P2 = {'q=u/8,jp=u%8': lambda u, r: 64*(u//8)+(u % 8)+8*r, 'q=u%8,jp=u/8': lambda u, r: 64*(u % 8)+(u//8)+8*r}
P3 = {'q=u/8,q2=u%8': lambda u, jp: 64*(u//8)+8*(u % 8)+jp, 'q=u%8,q2=u/8': lambda u, jp: 64*(u % 8)+8*(u//8)+jp}


def score(pad):
    ph0 = avg(pad, lambda u, c: 2*u+128*(c % 4))
    p1r = avg(pad, lambda u, r: u+64*r)
    p1w = avg(pad, lambda u, q: 64*q+u)
    p2 = min((avg(pad, f), n) for n, f in P2.items())
    p3 = min((avg(pad, f), n) for n, f in P3.items())
    fk = avg(pad, lambda u, c: pos_of_freq(u+32*c))
    fn = avg(pad, lambda u, c: pos_of_freq((N-(u+32*c)) % N))
    return 2*ph0+p1r+p1w+2*p2[0]+2*p3[0]+fk+fn, dict(ph0=ph0, p1r=p1r, p1w=p1w, p2=p2, p3=p3, fin_k=fk, fin_Nk=fn)


if __name__ == '__main__':
    print('round-1/2 layout p + p/8 :', score(lambda p: p+(p >> 3)))
    best = []
    for a, b, c, d, e in itertools.product(range(3), range(3), range(5), range(5), range(5)):
        def pad(p, a=a, b=b, c=c, d=d, e=e):
            return p+a*(p >> 3)+b*(p >> 4)+c*(p >> 5)+d*(p >> 6)+e*(p >> 7)
        if pad(511) > 511+96:
            continue
        best.append((score(pad), (a, b, c, d, e)))
    best.sort(key=lambda x: x[0][0])
    print('best paddings p + a p/8 + b p/16 + c p/32 + d p/64 + e p/128 (ideal total 20):')
    for (t, det), k in best[:5]:
        print(' total %.1f  (a,b,c,d,e)=%s  %s' % (t, k, det))
