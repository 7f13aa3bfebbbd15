"""The xi_tide table of IMRPhenomNSBH for the reference run under the shim (TEST INFRASTRUCTURE, container only).

The reference reads ``WFfiles/xiTide_Table_200.h5`` (gwfast/waveforms.py:2798, 3345-3373); that blob is not part of the reference
tree here, and in its absence the reference tabulates it itself with ``IMRPhenomNSBH._tabulate_xiTide`` (waveforms.py:3286-3343:
for every node of a 200^3 grid in (compactness, mass ratio, BH spin) the largest real positive root of the order-10 polynomial
of arXiv:1509.00512 eq. (8), by ``numpy.roots``).  That function is deterministic, so its output IS the table; it takes ~10
minutes, and as shipped it stops on ``time.time()`` (``time`` is never imported in waveforms.py) and on the h5py write.

``install()`` therefore (i) puts ``time`` into the reference module's namespace, and (ii) wraps ``_tabulate_xiTide`` so that the
UNMODIFIED function body runs once with ``store=False`` and its result is cached under ``oracle/_ref/`` (git-ignored); later calls
return the cached arrays.  No arithmetic of the reference is replaced.
"""
import os
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
CACHE = os.path.join(HERE, '_ref', 'xiTide_Table_200.npy')


def grids(res=200, Compmin=.1, qmax=100.):
    # waveforms.py:3308-3310
    return np.linspace(Compmin, .5, res), np.linspace(1., qmax, res), np.linspace(-1., 1., res)


def install(wf):
    """wf: the reference's waveforms module (oracle.reference.load()[0])."""
    if getattr(wf.IMRPhenomNSBH, '_oracle_cached', False):
        return
    wf.time = time
    original = wf.IMRPhenomNSBH._tabulate_xiTide

    def cached(self, res=200, store=True, Compmin=.1, qmax=100.):
        if res != 200 or Compmin != .1 or qmax != 100.:
            return original(self, res=res, store=False, Compmin=Compmin, qmax=qmax)
        if os.path.exists(CACHE):
            xi = np.load(CACHE)
        else:
            xi = original(self, res=res, store=False, Compmin=Compmin, qmax=qmax)[0]
            os.makedirs(os.path.dirname(CACHE), exist_ok=True)
            np.save(CACHE, xi)
        return (xi,) + grids(res, Compmin, qmax)

    wf.IMRPhenomNSBH._tabulate_xiTide = cached
    wf.IMRPhenomNSBH._oracle_cached = True


def table():
    """the cached reference table (None if it has not been generated in this container)"""
    return np.load(CACHE) if os.path.exists(CACHE) else None
