"""Multi-GPU sharding of a catalog: one process per GPU (torchrun), contiguous event shards, ONE all-gather at the end.

Events are independent (no cross-event term anywhere in gwfast/signal.py or network.py), so the only exchange step of the
path is the final gather of the results -- the same partition as the reference's batch pool
(run/calculate_forecasts_from_catalog.py:900-1007, 1016-1024), with NCCL over NVLink instead of result files.  With the
``nccl`` backend the gather runs on device tensors; with ``gloo`` (CPU tests) on host tensors.
"""
import numpy as np


def shard_bounds(n, world, rank):
    """contiguous slice [lo, hi) of rank `rank`: first (n % world) ranks get one extra event"""
    base, extra = divmod(int(n), int(world))
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def shard_events(evParams, world, rank):
    n = len(np.atleast_1d(evParams['Mc']))
    lo, hi = shard_bounds(n, world, rank)
    return {k: np.ascontiguousarray(np.asarray(v)[lo:hi]) for k, v in evParams.items()}, lo, hi


def all_gather_event_axis(local, n_total, dist=None, device=None):
    """All-gather arrays whose LAST axis is this rank's shard of the event axis into the full (..., n_total) array.

    `local`: numpy array (host) or torch tensor (device).  Shards may be uneven: every rank pads to the largest shard,
    one all_gather_into_tensor moves the padded blocks, and the padding is dropped when the result is assembled.
    """
    import torch
    if dist is None:
        import torch.distributed as dist
    world, rank = dist.get_world_size(), dist.get_rank()
    was_numpy = isinstance(local, np.ndarray)
    t = torch.from_numpy(np.ascontiguousarray(local)) if was_numpy else local
    if device is not None:
        t = t.to(device)
    lead = tuple(t.shape[:-1])
    nmax = max(shard_bounds(n_total, world, r)[1] - shard_bounds(n_total, world, r)[0] for r in range(world))
    pad = torch.zeros(lead + (nmax,), dtype=t.dtype, device=t.device)
    pad[..., :t.shape[-1]] = t
    out = torch.empty((world,) + lead + (nmax,), dtype=t.dtype, device=t.device)
    dist.all_gather_into_tensor(out.view(-1), pad.contiguous().view(-1))
    parts = []
    for r in range(world):
        lo, hi = shard_bounds(n_total, world, r)
        parts.append(out[r][..., :hi - lo])
    full = torch.cat(parts, dim=-1)
    return full.cpu().numpy() if was_numpy or device is not None else full


class DistributedDetNet(object):
    """A DetNet whose SNR / FisherMatr shard the catalog over the ranks of the default process group.

    Every rank passes the FULL events dict (as every worker of the reference's pool loads the catalog) and receives the
    FULL result; the work done by a rank is its contiguous shard.
    """

    def __init__(self, net, dist=None, device=None):
        self.net = net
        self.device = device
        if dist is None:
            import torch.distributed as dist
        self.dist = dist

    def _device(self):
        if self.device is not None:
            return self.device
        if self.dist.get_backend() == 'nccl':
            import torch
            return torch.device('cuda', torch.cuda.current_device())
        return None

    def SNR(self, evParams, res=1000):
        n = len(np.atleast_1d(evParams['Mc']))
        local, lo, hi = shard_events(evParams, self.dist.get_world_size(), self.dist.get_rank())
        snr = self.net.SNR(local, res=res) if hi > lo else np.zeros(0)
        return all_gather_event_axis(np.asarray(snr, dtype=np.float64), n, self.dist, self._device())

    def FisherMatr(self, evParams, gather='host', **kwargs):
        """``gather='host'`` (default): the full (nP, nP, N) numpy array on every rank.
        ``gather='device'`` (NCCL): returns ``(F_local, F_all)`` -- this rank's shard as a numpy array, and the full matrix as a
        device tensor gathered over NVLink straight from the engine's device-resident result (no host round trip); hosts only read
        their own shard, so the step does not move world x the result through every rank's PCIe link."""
        n = len(np.atleast_1d(evParams['Mc']))
        local, lo, hi = shard_events(evParams, self.dist.get_world_size(), self.dist.get_rank())
        nP = next(iter(self.net.signals.values())).wf_model.nParams
        if gather == 'device' and self._device() is not None:
            return fisher_with_device_gather(self.net, local, n, self.dist, **kwargs)
        F = self.net.FisherMatr(local, **kwargs) if hi > lo else np.zeros((nP, nP, 0))
        full = all_gather_event_axis(np.asarray(F, dtype=np.float64), n, self.dist, self._device())
        return (F, full) if gather == 'device' else full


def fisher_with_device_gather(net, local_events, n_total, dist=None, **kwargs):
    """``net.FisherMatr(local_events)`` on this rank's shard plus one all-gather of the device-resident result: returns
    ``(F_local numpy (nP, nP, m), F_all device tensor (nP, nP, n_total))``."""
    import torch
    from . import _engine
    _engine.STASH_DEVICE = True
    try:
        F = net.FisherMatr(local_events, **kwargs)
        dev = _engine.state().last_fisher_device
    finally:
        _engine.STASH_DEVICE = False
        _engine.state().last_fisher_device = None
    F = np.asarray(F, dtype=np.float64)
    if dev is not None and dev.shape[0] == 1 and tuple(dev.shape[1:]) == F.shape:
        local_dev = dev[0]
    else:       # per-arm / re-indexed results (return_all, duty factors, NewtInspiral): upload the host result
        local_dev = torch.from_numpy(np.ascontiguousarray(F)).to(_engine.state().device)
    return F, all_gather_event_axis(local_dev, n_total, dist)
