"""jax.numpy -> numpy (TEST INFRASTRUCTURE; see jax/__init__.py)."""
from numpy import *  # noqa: F401,F403
import numpy as _np

# names the reference uses that `from numpy import *` does not export or that need care
newaxis = _np.newaxis
pi = _np.pi
euler_gamma = _np.euler_gamma
trapezoid = _np.trapezoid
abs = _np.abs
sum = _np.sum
any = _np.any
all = _np.all
min = _np.min
max = _np.max
round = _np.round
complex128 = _np.complex128
float64 = _np.float64
ndarray = _np.ndarray


def fabs(x, *a, **k):
    """numpy.fabs, tolerant of complex input with zero imaginary part.  GWSignal.FisherMatr(return_SNR_derivatives=True) calls
    GWstrain with the parameters cast to complex128 (gwfast/signal.py:815-816, 1017); IMRPhenomD's np.fabs (waveforms.py:1134, 1193)
    then raises TypeError in numpy and in JAX alike -- a defect of the reference on that path.  The golden SNR derivatives
    are generated with this one call made lenient (the imaginary parts are identically zero)."""
    if _np.iscomplexobj(x):
        x = _np.real(x)
    return _np.fabs(x, *a, **k)


def asarray(x, *a, **k):
    """numpy.asarray that lets dual numbers through.  IMRPhenomNSBH stacks three per-event quantities for its table interpolator with
    ``np.asarray((np.asarray(Comp), np.asarray(q), np.asarray(chi1))).T`` (gwfast/waveforms.py:3111); jax.numpy.asarray accepts
    tracers there, numpy.asarray would turn the oracle's Dual objects into an object array."""
    from oracle.dual import Dual, stack0
    if isinstance(x, Dual):
        return x
    if isinstance(x, (tuple, list)) and True in [isinstance(t, Dual) for t in x]:
        return stack0(x)
    return _np.asarray(x, *a, **k)
