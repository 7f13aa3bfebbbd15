"""Forward-mode dual-number ndarray (TEST INFRASTRUCTURE, not product code).

This module is part of ``oracle/``: it is only ever imported by ``tests/``,
``__graft_entry__.smoke()`` and the ``cpu_baseline`` / ``--impl reference`` legs
of ``bench.py``.  The product path (``gwfast_b200``) never imports it.

A :class:`Dual` carries a value array ``v`` and a tangent array ``d`` with
``d.shape == v.shape + (nt,)`` (tangent axis LAST, so per-event duals of shape
``(N,)`` broadcast against ``(res, N)`` grids).  It implements the numpy
``__array_ufunc__`` / ``__array_function__`` protocols for exactly the calls the
reference hot path makes through ``jax.numpy`` (census: SURVEY.md App. C), which
lets the reference's own ``GWSignal.GWstrain`` run unmodified under
``oracle/refshim`` and lets ``oracle/port`` differentiate its restatement the
same way ``jax.jacrev`` does mathematically (gwfast/signal.py:1175-1189).
"""
import numpy as np

_HANDLED = {}


def _implements(np_function):
    def decorator(func):
        _HANDLED[np_function] = func
        return func
    return decorator


def _val(x):
    return x.v if isinstance(x, Dual) else x


def _ex(a):
    """append a broadcast axis for the tangent dimension"""
    a = np.asarray(a)
    return a[..., None]


def _nt_of(*xs):
    for x in xs:
        if isinstance(x, Dual):
            return x.d.shape[-1]
    return None


def _tan(x, nt, shape=None):
    """tangent array of x (zeros for plain operands)"""
    if isinstance(x, Dual):
        return x.d
    return 0.0


class Dual:
    __array_priority__ = 1000.0

    def __init__(self, v, d):
        self.v = np.asarray(v)
        self.d = np.asarray(d)

    # ------------------------------------------------------------ attributes
    @property
    def shape(self):
        return self.v.shape

    @property
    def ndim(self):
        return self.v.ndim

    @property
    def dtype(self):
        return self.v.dtype

    @property
    def real(self):
        return Dual(self.v.real, self.d.real)

    @property
    def imag(self):
        return Dual(self.v.imag, self.d.imag)

    @property
    def T(self):
        axes = tuple(range(self.v.ndim))[::-1]
        return Dual(self.v.T, np.transpose(self.d, axes + (self.v.ndim,)))

    def reshape(self, *shape):
        if len(shape) == 1 and isinstance(shape[0], (tuple, list)):
            shape = tuple(shape[0])
        v = self.v.reshape(shape)
        return Dual(v, self.d.reshape(v.shape + (self.d.shape[-1],)))

    def astype(self, dt):
        return Dual(self.v.astype(dt), self.d.astype(dt))

    def conj(self):
        return Dual(np.conj(self.v), np.conj(self.d))

    def __len__(self):
        return len(self.v)

    def __getitem__(self, idx):
        return Dual(self.v[idx], self.d[idx])

    def __repr__(self):
        return "Dual(v=%r, nt=%d)" % (self.v, self.d.shape[-1])

    # ------------------------------------------------------------ arithmetic
    def __neg__(self):
        return Dual(-self.v, -self.d)

    def __pos__(self):
        return self

    def __abs__(self):
        return _abs(self)

    def __add__(self, o):
        if isinstance(o, Dual):
            return Dual(self.v + o.v, self.d + o.d)
        o = np.asarray(o)
        v = self.v + o
        return Dual(v, self.d + np.zeros(v.shape + (1,), dtype=v.dtype))

    __radd__ = __add__

    def __sub__(self, o):
        if isinstance(o, Dual):
            return Dual(self.v - o.v, self.d - o.d)
        o = np.asarray(o)
        v = self.v - o
        return Dual(v, self.d + np.zeros(v.shape + (1,), dtype=v.dtype))

    def __rsub__(self, o):
        o = np.asarray(o)
        v = o - self.v
        return Dual(v, -self.d + np.zeros(v.shape + (1,), dtype=v.dtype))

    def __mul__(self, o):
        if isinstance(o, Dual):
            return Dual(self.v * o.v, self.d * _ex(o.v) + _ex(self.v) * o.d)
        o = np.asarray(o)
        return Dual(self.v * o, self.d * _ex(o))

    __rmul__ = __mul__

    def __truediv__(self, o):
        if isinstance(o, Dual):
            # the VALUE is the plain quotient, bit for bit what the undifferentiated code computes: comparisons such as
            # `fgrid < fcutPar` at the last grid sample (Mf = 0.2 up to the last bit, waveforms.py:2290, 2178) must take the same branch
            # in the derivative pass as in the value pass (a reciprocal-multiply here switched IMRPhenomHM's (2,2) mode on in the
            # derivative of a sample whose value had it off)
            q = self.v / o.v
            return Dual(q, (self.d - _ex(q) * o.d) / _ex(o.v))
        o = np.asarray(o)
        return Dual(self.v / o, self.d / _ex(o))

    def __rtruediv__(self, o):
        o = np.asarray(o)
        q = o / self.v
        return Dual(q, -_ex(q / self.v) * self.d)

    def __pow__(self, p):
        return _power(self, p)

    def __rpow__(self, b):
        return _power(b, self)

    # comparisons act on values and return plain boolean arrays
    def __lt__(self, o):
        return self.v < _val(o)

    def __le__(self, o):
        return self.v <= _val(o)

    def __gt__(self, o):
        return self.v > _val(o)

    def __ge__(self, o):
        return self.v >= _val(o)

    def __eq__(self, o):
        return self.v == _val(o)

    def __ne__(self, o):
        return self.v != _val(o)

    __hash__ = None

    # ------------------------------------------------------------ numpy protocols
    def __array_ufunc__(self, ufunc, method, *inputs, **kwargs):
        if method != '__call__' or kwargs.get('out') is not None:
            return NotImplemented
        f = _UFUNCS.get(ufunc)
        if f is None:
            raise NotImplementedError("Dual: ufunc %s not supported" % ufunc.__name__)
        return f(*inputs)

    def __array_function__(self, func, types, args, kwargs):
        f = _HANDLED.get(func)
        if f is None:
            raise NotImplementedError("Dual: function %s not supported" % func.__name__)
        return f(*args, **kwargs)


# ---------------------------------------------------------------- ufunc table
def _power(a, p):
    if isinstance(a, Dual) and not isinstance(p, Dual):
        p = np.asarray(p)
        v = a.v ** p
        with np.errstate(all='ignore'):
            dv = p * a.v ** (p - 1)
        return Dual(v, _ex(dv) * a.d)
    if isinstance(p, Dual) and not isinstance(a, Dual):
        a = np.asarray(a)
        v = a ** p.v
        return Dual(v, _ex(v * np.log(a)) * p.d)
    v = a.v ** p.v
    with np.errstate(all='ignore'):
        return Dual(v, _ex(p.v * a.v ** (p.v - 1)) * a.d + _ex(v * np.log(a.v)) * p.d)


def _abs(a):
    if np.iscomplexobj(a.v):
        r = np.abs(a.v)
        with np.errstate(all='ignore'):
            return Dual(r, np.real(_ex(np.conj(a.v)) * a.d) / _ex(r))
    return Dual(np.abs(a.v), _ex(np.sign(a.v)) * a.d)


def _sqrt(a):
    """d sqrt(x) = dx / (2 sqrt(x)).  Where x = 0 AND dx = 0 exactly -- the constant branch of
    ``sqrt(where(eta < 0.25, 1 - 4 eta, 0.))`` at eta >= 0.25 (gwfast/waveforms.py:759, 1013) -- the tangent is 0, which is what
    ``jax.jacrev`` returns there: the cotangent g / (2 * 0) that reaches the ``where`` is dropped by the select's transpose (the
    unselected branch receives zeros, nothing is multiplied by 0), so the reference differentiates Seta to 0 at eta = 0.25."""
    v = np.sqrt(a.v)
    with np.errstate(all='ignore'):
        d = _ex(0.5 / v) * a.d
    return Dual(v, np.where(_ex(v == 0) & (a.d == 0), 0.0, d))


def _unary(fun, dfun):
    def g(a):
        v = fun(a.v)
        with np.errstate(all='ignore'):
            return Dual(v, _ex(dfun(a.v, v)) * a.d)
    return g


def _as_dual(x, like):
    if isinstance(x, Dual):
        return x
    x = np.asarray(x)
    return Dual(x, np.zeros(x.shape + (like.d.shape[-1],), dtype=x.dtype))


def _arctan2(y, x):
    like = y if isinstance(y, Dual) else x
    y, x = _as_dual(y, like), _as_dual(x, like)
    r2 = x.v * x.v + y.v * y.v
    return Dual(np.arctan2(y.v, x.v), (_ex(x.v) * y.d - _ex(y.v) * x.d) / _ex(r2))


def _zero_tangent(fun):
    def g(a):
        return Dual(fun(a.v), np.zeros_like(a.d))
    return g


_UFUNCS = {
    np.add: lambda a, b: a + b if isinstance(a, Dual) else b + a,
    np.subtract: lambda a, b: a - b if isinstance(a, Dual) else b.__rsub__(a),
    np.multiply: lambda a, b: a * b if isinstance(a, Dual) else b * a,
    np.divide: lambda a, b: a / b if isinstance(a, Dual) else b.__rtruediv__(a),
    np.power: _power,
    np.negative: lambda a: -a,
    np.positive: lambda a: a,
    np.absolute: _abs,
    np.fabs: _abs,
    np.conjugate: lambda a: a.conj(),
    np.exp: _unary(np.exp, lambda x, v: v),
    np.log: _unary(np.log, lambda x, v: 1.0 / x),
    np.sqrt: _sqrt,
    np.cbrt: _unary(np.cbrt, lambda x, v: v / (3.0 * x)),
    np.sin: _unary(np.sin, lambda x, v: np.cos(x)),
    np.cos: _unary(np.cos, lambda x, v: -np.sin(x)),
    np.tan: _unary(np.tan, lambda x, v: 1.0 + v * v),
    np.arctan: _unary(np.arctan, lambda x, v: 1.0 / (1.0 + x * x)),
    np.tanh: _unary(np.tanh, lambda x, v: 1.0 - v * v),
    np.square: _unary(np.square, lambda x, v: 2.0 * x),
    np.reciprocal: _unary(np.reciprocal, lambda x, v: -v * v),
    np.arctan2: _arctan2,
    np.floor: _zero_tangent(np.floor),
    np.sign: _zero_tangent(np.sign),
    np.less: lambda a, b: _val(a) < _val(b),
    np.less_equal: lambda a, b: _val(a) <= _val(b),
    np.greater: lambda a, b: _val(a) > _val(b),
    np.greater_equal: lambda a, b: _val(a) >= _val(b),
    np.equal: lambda a, b: _val(a) == _val(b),
    np.not_equal: lambda a, b: _val(a) != _val(b),
    np.isnan: lambda a: np.isnan(a.v),
    np.isfinite: lambda a: np.isfinite(a.v),
}


# ---------------------------------------------------------------- array functions
@_implements(np.where)
def _where(cond, x, y):
    cond = np.asarray(_val(cond))
    like = x if isinstance(x, Dual) else y
    x, y = _as_dual(x, like), _as_dual(y, like)
    return Dual(np.where(cond, x.v, y.v), np.where(_ex(cond), x.d, y.d))


@_implements(np.interp)
def _interp(x, xp, fp, left=None, right=None, period=None):
    # piecewise-linear; tangent = slope of the containing segment (right-continuous
    # at the nodes), zero outside the table -- what jax.numpy.interp differentiates to.
    xp = np.asarray(xp)
    fp = np.asarray(fp)
    v = np.interp(x.v, xp, fp, left=left, right=right)
    i = np.clip(np.searchsorted(xp, x.v, side='right') - 1, 0, len(xp) - 2)
    slope = (fp[i + 1] - fp[i]) / (xp[i + 1] - xp[i])
    slope = np.where((x.v < xp[0]) | (x.v > xp[-1]), 0.0, slope)
    return Dual(v, _ex(slope) * x.d)


@_implements(np.searchsorted)
def _searchsorted(a, v, side='left', sorter=None):
    # indices are piecewise constant: values only (gwfastUtils.py:1196, RegularGridInterpolator_JAX._find_indices)
    return np.searchsorted(_val(a), _val(v), side=side, sorter=sorter)


@_implements(np.amin)
def _amin(a, axis=None, **kw):
    i = np.expand_dims(np.argmin(a.v, axis=axis), axis)
    v = np.squeeze(np.take_along_axis(a.v, i, axis=axis), axis=axis)
    d = np.squeeze(np.take_along_axis(a.d, _ex(i), axis=axis), axis=axis)
    return Dual(v, d)


@_implements(np.amax)
def _amax(a, axis=None, **kw):
    i = np.expand_dims(np.argmax(a.v, axis=axis), axis)
    v = np.squeeze(np.take_along_axis(a.v, i, axis=axis), axis=axis)
    d = np.squeeze(np.take_along_axis(a.d, _ex(i), axis=axis), axis=axis)
    return Dual(v, d)


_HANDLED[np.min] = _amin
_HANDLED[np.max] = _amax


@_implements(np.real)
def _real(a):
    return a.real


@_implements(np.imag)
def _imag(a):
    return a.imag


@_implements(np.conj)
def _conj(a):
    return a.conj()


@_implements(np.abs)
def _npabs(a):
    return _abs(a)


def _norm_axis(axis, ndim):
    return axis if axis >= 0 else axis + ndim


@_implements(np.sum)
def _sum(a, axis=None, **kw):
    if axis is None:
        return Dual(a.v.sum(), a.d.reshape(-1, a.d.shape[-1]).sum(axis=0))
    ax = _norm_axis(axis, a.v.ndim)
    return Dual(a.v.sum(axis=ax), a.d.sum(axis=ax))


@_implements(np.expand_dims)
def _expand_dims(a, axis):
    ax = axis if axis >= 0 else axis + a.v.ndim + 1
    return Dual(np.expand_dims(a.v, ax), np.expand_dims(a.d, ax))


@_implements(np.moveaxis)
def _moveaxis(a, source, destination):
    n = a.v.ndim
    return Dual(np.moveaxis(a.v, source, destination),
                np.moveaxis(a.d, _norm_axis(source, n), _norm_axis(destination, n)))


@_implements(np.squeeze)
def _squeeze(a, axis=None):
    if axis is None:
        axis = tuple(i for i, s in enumerate(a.v.shape) if s == 1)
    elif isinstance(axis, int):
        axis = (_norm_axis(axis, a.v.ndim),)
    return Dual(np.squeeze(a.v, axis=axis), np.squeeze(a.d, axis=axis))


@_implements(np.transpose)
def _transpose(a, axes=None):
    if axes is None:
        return a.T
    return Dual(np.transpose(a.v, axes), np.transpose(a.d, tuple(axes) + (a.v.ndim,)))


@_implements(np.nan_to_num)
def _nan_to_num(a, **kw):
    return Dual(np.nan_to_num(a.v, **kw), np.nan_to_num(a.d, **kw))


@_implements(np.zeros_like)
def _zeros_like(a, **kw):
    return np.zeros_like(a.v, **kw)


@_implements(np.ones_like)
def _ones_like(a, **kw):
    return np.ones_like(a.v, **kw)


@_implements(np.shape)
def _shape(a):
    return a.v.shape


@_implements(np.ndim)
def _ndim(a):
    return a.v.ndim


@_implements(np.iscomplexobj)
def _iscomplexobj(a):
    return np.iscomplexobj(a.v)


@_implements(np.any)
def _any(a, **kw):
    return np.any(a.v, **kw)


@_implements(np.broadcast_to)
def _broadcast_to(a, shape):
    return Dual(np.broadcast_to(a.v, shape), np.broadcast_to(a.d, tuple(shape) + (a.d.shape[-1],)))


def _stack_like(fun):
    def g(arrs, *args, **kw):
        like = next(x for x in arrs if isinstance(x, Dual))
        arrs = [_as_dual(x, like) for x in arrs]
        vs = np.broadcast_arrays(*[x.v for x in arrs])
        ds = [np.broadcast_to(x.d, v.shape + (like.d.shape[-1],)) for x, v in zip(arrs, vs)]
        return Dual(fun(vs, *args, **kw), fun(ds, *args, **kw))
    return g


_HANDLED[np.stack] = _stack_like(np.stack)


@_implements(np.vstack)
def _vstack(arrs):
    like = next(x for x in arrs if isinstance(x, Dual))
    arrs = [_as_dual(x, like) for x in arrs]
    return Dual(np.vstack([np.atleast_2d(x.v) for x in arrs]),
                np.concatenate([x.d if x.v.ndim >= 2 else x.d[None] for x in arrs], axis=0))


# ---------------------------------------------------------------- seeding helpers
def stack0(xs):
    """np.asarray of a tuple of duals / arrays: stacked along a new leading axis"""
    like = next(x for x in xs if isinstance(x, Dual))
    ds = [_as_dual(x, like) for x in xs]
    shape = np.broadcast_shapes(*[x.v.shape for x in ds])
    nt = like.d.shape[-1]
    return Dual(np.stack([np.broadcast_to(x.v, shape) for x in ds]), np.stack([np.broadcast_to(x.d, shape + (nt,)) for x in ds]))


def seed(args, argnums):
    """Return ``args`` with those listed in ``argnums`` replaced by identity-seeded duals."""
    nt = len(argnums)
    out = list(args)
    for k, j in enumerate(argnums):
        a = np.asarray(args[j])
        d = np.zeros(a.shape + (nt,), dtype=a.dtype)
        d[..., k] = 1.0
        out[j] = Dual(a, d)
    return out
