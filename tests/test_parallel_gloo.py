"""Sharding + final all-gather of the multi-GPU path, exercised with world_size = 2 on CPU (gloo).  The per-shard compute is
the oracle port (no GPU in this container); the GPU ranks run the same DistributedDetNet with the CUDA DetNet underneath."""
import os
import socket
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_shard_bounds_cover_the_catalog():
    from gwfast_b200.parallel import shard_bounds
    for n in (0, 1, 7, 10, 10000, 100003):
        for world in (1, 2, 3, 4, 8):
            b = [shard_bounds(n, world, r) for r in range(world)]
            assert b[0][0] == 0 and b[-1][1] == n
            assert all(b[i][1] == b[i + 1][0] for i in range(world - 1))
            sizes = [hi - lo for lo, hi in b]
            assert max(sizes) - min(sizes) <= 1


def _worker(rank, world, port, n, q):
    sys.path.insert(0, ROOT)
    import warnings
    warnings.filterwarnings('ignore')
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group('gloo', rank=rank, world_size=world)
    from gwfast_b200 import synthetic
    from gwfast_b200.parallel import DistributedDetNet
    from oracle.port import waveforms as PW, detector as PD
    net = PD.Network(synthetic.build_network(PD.Detector, PW.TaylorF2_RestrictedPN(), 'ETSL'))
    ev = synthetic.bns_catalog(n, 11)
    dnet = DistributedDetNet(net)
    snr = dnet.SNR(dict(ev), res=200)
    F = dnet.FisherMatr(dict(ev), res=200)
    q.put((rank, snr, F))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize('n', [7, 12])
def test_two_rank_gather_equals_single_process(n):
    import torch.multiprocessing as mp
    sys.path.insert(0, ROOT)
    from gwfast_b200 import synthetic
    from oracle.port import waveforms as PW, detector as PD
    s = socket.socket()
    s.bind(('127.0.0.1', 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, n, q)) for r in range(2)]
    for p in procs:
        p.start()
    got = [q.get(timeout=300) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    net = PD.Network(synthetic.build_network(PD.Detector, PW.TaylorF2_RestrictedPN(), 'ETSL'))
    ev = synthetic.bns_catalog(n, 11)
    snr, F = net.SNR(dict(ev), res=200), net.FisherMatr(dict(ev), res=200)
    for rank, s_r, F_r in got:
        # events are independent: the sharded result is bitwise the single-process one, on every rank
        assert s_r.shape == (n,) and F_r.shape == (11, 11, n)
        assert np.array_equal(s_r, snr) and np.array_equal(F_r, F)
