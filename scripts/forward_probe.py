"""Cost of the in-kernel forwarding of finished rows (gwf_fisher_out.peer_fisher) with LOCAL slots standing in for the peers: separates the
instruction/issue cost of the extra stores from the cost of sending them over NVLink."""
import os, sys, ctypes as C
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import bench
from gwfast_b200 import synthetic
torch.cuda.set_device(0)
ev = synthetic.bbh_catalog(10000, synthetic.SEEDS['C2'])
c = bench.Case('p', 'IMRPhenomD', 'ET+2CE', ev)
flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device='cuda')
def timeit(tag):
    for _ in range(3): c.fisher()
    ts = []
    for _ in range(12):
        flush.fill_(1); torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); c.fisher(reuse=True); e1.record(); torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1))
    print('%-28s fisher_kernel %.4f ms (min %.4f)' % (tag, np.median(ts), min(ts)), flush=True)
timeit('no peers')
for npeer in (1, 2, 4, 8):
    bufs = [torch.zeros((c.n, c.npack), dtype=torch.float64, device='cuda') for _ in range(npeer)]
    arr = (C.c_void_p * npeer)(*[b.data_ptr() for b in bufs])
    c.fo.peer_fisher = C.cast(arr, C.c_void_p); c.fo.npeers = npeer
    timeit('%d local slots' % npeer)
    assert torch.equal(bufs[-1], c.packed)
