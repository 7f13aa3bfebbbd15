"""Two-rank probe of parallel.PeerGather (torchrun --nproc-per-node 2 scripts/peer_probe.py)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, torch.distributed as dist
local = int(os.environ['LOCAL_RANK'])
torch.cuda.set_device(local)
dist.init_process_group('nccl', device_id=torch.device('cuda', local))
from gwfast_b200 import parallel
rank, world = dist.get_rank(), dist.get_world_size()
n, npack, nP = 1000, 66, 11
pg = parallel.PeerGather(n, npack, dist)
print(rank, 'available', pg.available(), getattr(pg, '_error', None), flush=True)
if pg.available():
    print(rank, 'ptrs', [hex(p) for p in pg._ptrs], pg.gathered.device, tuple(pg.gathered.shape), flush=True)
    packed = torch.arange(n * npack, dtype=torch.float64, device='cuda:%d' % local).view(n, npack) + 1e6 * rank
    full = torch.empty((nP, nP, n), dtype=torch.float64, device='cuda:%d' % local)
    pg.unpack_and_scatter(packed, n, nP, full, n, torch.cuda.current_stream())
    torch.cuda.synchronize()
    print(rank, 'kernel ok', flush=True)
    g = pg.finish()
    for r in range(world):
        want = torch.arange(n * npack, dtype=torch.float64, device='cuda:%d' % local).view(n, npack) + 1e6 * r
        print(rank, 'slot', r, 'equal', bool(torch.equal(g[r], want)), flush=True)
pg.close()
dist.barrier()
dist.destroy_process_group()
