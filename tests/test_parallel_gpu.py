"""Multi-GPU plumbing on the device (one-rank NCCL group: the N > 1 logic is covered by the gloo tests): the all-gather taken
straight from the engine's device-resident Fisher matrices returns exactly what the host path returns."""
import numpy as np
import pytest

from conftest import make_network, copy_events, load_golden


@pytest.mark.gpu
def test_device_gather_matches_host_path():
    import torch
    import torch.distributed as dist
    from gwfast_b200 import synthetic, parallel
    if not dist.is_initialized():
        dist.init_process_group('nccl', init_method='tcp://127.0.0.1:29533', world_size=1, rank=0, device_id=torch.device('cuda', 0))
    try:
        cfg = dict(model=dict(cls='IMRPhenomD'), network='ET+2CE', rot=True, fmin=2.)
        net = make_network('engine', cfg)
        ev = synthetic.bbh_catalog(257, 99)
        F = net.FisherMatr(copy_events(ev))
        dnet = parallel.DistributedDetNet(net)
        F_local, F_all = dnet.FisherMatr(copy_events(ev), gather='device')
        assert isinstance(F_all, torch.Tensor) and F_all.is_cuda and tuple(F_all.shape) == F.shape
        assert np.array_equal(F_local, F) and np.array_equal(F_all.cpu().numpy(), F)
        assert np.array_equal(dnet.FisherMatr(copy_events(ev)), F)
        assert np.array_equal(dnet.SNR(copy_events(ev)), net.SNR(copy_events(ev)))
        # results that are re-indexed on the host (NewtInspiral's 8 parameters) take the upload branch
        cfgn, evn, _ = load_golden('newt_et')
        netn = make_network('engine', cfgn)
        Fn = netn.FisherMatr(copy_events(evn), **cfgn.get('fisher_kw', {}))
        Fl, Fa = parallel.DistributedDetNet(netn).FisherMatr(copy_events(evn), gather='device', **cfgn.get('fisher_kw', {}))
        assert np.array_equal(Fl, Fn) and np.array_equal(Fa.cpu().numpy(), Fn)
    finally:
        dist.destroy_process_group()
