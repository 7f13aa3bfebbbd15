"""Minimal numpy-backed stand-in for the `jax` package (TEST INFRASTRUCTURE).

Exists only so that the reference's own modules (``/root/reference/gwfast``) can be
imported and run *unmodified* in a container that has no JAX: ``jax.numpy`` is numpy,
``vmap(jacrev(f, argnums))`` is forward-mode differentiation with ``oracle.dual.Dual``
(mathematically the same Jacobian; gwfast/signal.py:1175-1189), ``custom_jvp`` keeps the
reference's hand-written tangent rule (gwfast/waveforms.py:1702-1719), ``jit`` is identity.
Never imported by the product package.
"""
import numpy as _np
from oracle.dual import Dual, seed as _seed
from . import numpy  # noqa: F401  (jax.numpy)
from . import interpreters  # noqa: F401


class _Config:
    def update(self, *a, **k):
        pass


config = _Config()


def local_device_count():
    return 1


def device_count():
    return 1


def devices(*a, **k):          # gwfast/fisherTools.py:11 calls jax.devices('cpu') at import
    return ['cpu:0']


def jit(f, *a, **k):
    return f


def pmap(f, *a, **k):
    raise NotImplementedError("pmap is imported by the reference but never used")


class _Jac:
    """jacrev(f, argnums[, holomorphic]) -- evaluated in forward mode, batched."""

    def __init__(self, f, argnums=0, holomorphic=False):
        self.f = f
        self.argnums = (argnums,) if isinstance(argnums, int) else tuple(argnums)

    def batched(self, fgridsT, *params):
        # vmap maps axis 0 of every argument; the reference's GWstrain broadcasts
        # (res,N) grids against (N,) parameters, so one batched call is equivalent.
        args = [_np.asarray(fgridsT).T] + [_np.asarray(p) for p in params]
        out = self.f(*_seed(args, self.argnums))
        return tuple(_np.asarray(out.d[..., k]).T for k in range(len(self.argnums)))


def jacrev(f, argnums=0, holomorphic=False, **k):
    return _Jac(f, argnums, holomorphic)


jacfwd = jacrev


def vmap(g, *a, **k):
    if isinstance(g, _Jac):
        return g.batched
    raise NotImplementedError("shim vmap only wraps jacrev/jacfwd")


class custom_jvp:
    """custom_jvp with defjvps(None, fn): zero tangent for arg 0, fn's rule for arg 1."""

    def __init__(self, fun):
        self.fun = fun
        self.jvps = None

    def defjvps(self, *jvps):
        self.jvps = jvps

    def __call__(self, *args):
        prim = [a.v if isinstance(a, Dual) else a for a in args]
        out = self.fun(*prim)
        if not any(isinstance(a, Dual) for a in args):
            return out
        nt = next(a.d.shape[-1] for a in args if isinstance(a, Dual))
        tang = _np.zeros(_np.shape(out) + (nt,), dtype=_np.asarray(out).dtype)
        for a, rule in zip(args, self.jvps):
            if rule is None or not isinstance(a, Dual):
                continue
            for k in range(nt):
                tang[..., k] = tang[..., k] + rule(a.d[..., k], out, *prim)
        return Dual(out, tang)
