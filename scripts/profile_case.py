"""One configuration through the fused C-ABI step a few times (for ncu captures).  usage: profile_case.py NETWORK MODEL [n]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from gwfast_b200 import synthetic
netname, mname = sys.argv[1], sys.argv[2]
n = int(sys.argv[3]) if len(sys.argv) > 3 else 10000
if mname == 'IMRPhenomD_NRTidalv2':
    ev = synthetic.bns_catalog(n, synthetic.SEEDS['C3'], tidal=True)
elif mname == 'IMRPhenomNSBH':
    ev = synthetic.nsbh_catalog(n, synthetic.SEEDS['NSBH'])
elif mname == 'TaylorF2_RestrictedPN':
    ev = synthetic.bns_catalog(n, synthetic.SEEDS['C1'])
else:
    ev = synthetic.bbh_catalog(n, synthetic.SEEDS['C4' if mname == 'IMRPhenomHM' else 'C2'])
lvk = netname == 'LVK-O4'
torch.cuda.set_device(0)
c = bench.Case('p', mname, netname, ev, rot=not lvk, fmin=10. if lvk else 2.)
for _ in range(3):
    c.fisher()
    c.unpack()
    c.snr()
torch.cuda.synchronize()
print('done', mname, netname, n)
