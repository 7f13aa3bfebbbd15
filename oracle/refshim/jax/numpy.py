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
