"""Load the UNMODIFIED reference (``/root/reference/gwfast``) under the numpy/dual shim.

TEST INFRASTRUCTURE.  Works only where ``/root/reference`` is mounted (the build
container); used by ``oracle/make_golden.py`` to generate ``tests/golden/*.npz`` and by
container-side tests that pin ``oracle.port`` against the reference.  Nothing that runs on
the GPU box may call this.
"""
import contextlib
import io
import os
import sys

REFERENCE_ROOT = os.environ.get("GWFAST_REFERENCE_ROOT", "/root/reference")
_SHIM = os.path.join(os.path.dirname(os.path.abspath(__file__)), "refshim")


def available():
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "gwfast"))


def load():
    """Return the reference modules ``(waveforms, signal, network, utils, glob)``."""
    if not available():
        raise RuntimeError("reference tree not mounted at %s" % REFERENCE_ROOT)
    repo = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    for p in (repo, _SHIM, REFERENCE_ROOT):
        if p not in sys.path:
            sys.path.insert(0, p)
    import scipy.integrate as igt
    if not hasattr(igt, "cumtrapz"):          # gwfast/signal.py:134 (removed in scipy>=1.14)
        igt.cumtrapz = igt.cumulative_trapezoid
    import numpy as onp
    if not hasattr(onp, "trapz"):             # gwfast/signal.py:929
        onp.trapz = onp.trapezoid
    if "NaN" not in onp.__dict__:             # gwfast/fisherTools.py:371 (removed in numpy 2.0)
        onp.NaN = onp.nan
    with contextlib.redirect_stdout(io.StringIO()):
        from gwfast import waveforms, signal, network
        from gwfast import gwfastUtils as utils
        from gwfast import gwfastGlobals as glob
    from oracle import nsbh_table
    nsbh_table.install(waveforms)             # IMRPhenomNSBH: `time` import and a cache around the reference's own xi_tide tabulation
    return waveforms, signal, network, utils, glob
