import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'tests'))
import numpy as np
from conftest import load_golden, make_network, copy_events
for name in ('deriv_hm_strain_lvk', 'deriv_hm_strain_et'):
    cfg, ev, out = load_golden(name)
    net = make_network('engine', cfg)
    res = cfg.get('res', 1000)
    F, D = net.FisherMatr(copy_events(ev), res=res, return_derivatives=True)
    for k in D:
        ref = out['deriv__' + k]
        scale = np.max(np.abs(ref), axis=-1, keepdims=True) + 1e-300
        err = np.abs(D[k] - ref) / scale
        print(name, k, 'per-row max err', ['%.1e' % e for e in err.max(axis=(1, 2))])
        i, n, s = np.unravel_index(np.argmax(err), err.shape)
        print('   worst row %d event %d sample %d of %d: got %s ref %s ; neighbours err %s' % (i, n, s, res, D[k][i, n, s], ref[i, n, s], ['%.1e' % e for e in err[i, n, max(0, s - 2):s + 3]]))
        fe = np.abs(F[k] - out['fisher__' + k]) / np.sqrt(np.einsum('iin->in', out['fisher__' + k])[:, None] * np.einsum('iin->in', out['fisher__' + k])[None])
        print('   fisher err', fe.max())
        break
