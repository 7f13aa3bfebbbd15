"""Synthetic event catalogs and the networks of BASELINE.json's configs; used by tests and bench."""
import os
import numpy as np

from . import gwfastGlobals as glob

SEEDS = {'C1': 20260001, 'C2': 20260002, 'C3': 20260003, 'C4': 20260004, 'C5': 20260005, 'NSBH': 20260006}


def _angles(rng, N):
    return dict(theta=np.arccos(rng.uniform(-1, 1, N)), phi=rng.uniform(0, 2 * np.pi, N), iota=np.arccos(rng.uniform(-1, 1, N)),
                psi=rng.uniform(0, np.pi, N), tcoal=rng.uniform(0, 1, N), Phicoal=rng.uniform(0, 2 * np.pi, N))


def bbh_catalog(N, seed):
    """detector-frame BBH: Mc~logU(5,100), q~U(1,8), dL~U(0.5,30) Gpc, |chi|<0.8, isotropic angles."""
    rng = np.random.default_rng(seed)
    q = rng.uniform(1, 8, N)
    ev = dict(Mc=np.exp(rng.uniform(np.log(5.), np.log(100.), N)), eta=np.minimum(q / (1 + q) ** 2, 0.2499), dL=rng.uniform(0.5, 30, N))
    ev.update(_angles(rng, N))
    ev.update(chi1z=rng.uniform(-0.8, 0.8, N), chi2z=rng.uniform(-0.8, 0.8, N))
    return ev


def bns_catalog(N, seed, tidal=False):
    """detector-frame BNS: Mc~N(1.156,0.056)(1+z), z~U(0.01,3), eta~U(0.24,0.2499), |chi|<0.05, Lambda~U(5,2000)."""
    rng = np.random.default_rng(seed)
    z = rng.uniform(0.01, 3, N)
    ev = dict(Mc=rng.normal(1.156, 0.056, N) * (1 + z), eta=rng.uniform(0.24, 0.2499, N), dL=rng.uniform(0.05, 25, N))
    ev.update(_angles(rng, N))
    ev.update(chi1z=rng.uniform(-0.05, 0.05, N), chi2z=rng.uniform(-0.05, 0.05, N))
    if tidal:
        ev.update(Lambda1=rng.uniform(5, 2000, N), Lambda2=rng.uniform(5, 2000, N))
    return ev


def nsbh_catalog(N, seed):
    """detector-frame NSBH: m_BH~U(3,12)(1+z), m_NS~U(1.1,2.0)(1+z), z~U(0.01,1.5), chi_BH~U(-0.6,0.9), chi_NS~U(-0.05,0.05),
    Lambda_NS~U(50,3000) (a tenth of the events below 1: the polynomial branches of the compactness and quadrupole fits), Lambda_BH = 0."""
    rng = np.random.default_rng(seed)
    z = rng.uniform(0.01, 1.5, N)
    m1, m2 = rng.uniform(3., 12., N) * (1 + z), rng.uniform(1.1, 2.0, N) * (1 + z)
    ev = dict(Mc=(m1 * m2) ** 0.6 / (m1 + m2) ** 0.2, eta=m1 * m2 / (m1 + m2) ** 2, dL=rng.uniform(0.05, 12, N))
    ev.update(_angles(rng, N))
    lam = rng.uniform(50., 3000., N)
    small = rng.uniform(0, 1, N) < 0.1
    lam[small] = rng.uniform(0.05, 0.95, small.sum())
    ev.update(chi1z=rng.uniform(-0.6, 0.9, N), chi2z=rng.uniform(-0.05, 0.05, N), Lambda1=np.zeros(N), Lambda2=lam)
    return ev


def psd_path(rel):
    return os.path.join(glob.detPath, rel)


# network name -> list of (detector key, site key, PSD file relative to psds/)
NETWORKS = {
    'ETSL': [('ETSL', 'ETSL', 'ET-0000A-18.txt')],
    'ET': [('ET', 'ETS', 'ET-0000A-18.txt')],
    'ET+2CE': [('ET', 'ETS', 'ET-0000A-18.txt'), ('CE1Id', 'CE1Id', 'ce_strain/cosmic_explorer.txt'),
               ('CE2NM', 'CE2NM', 'ce_strain/cosmic_explorer_20km.txt')],
    'LVK-O4': [('H1', 'H1', 'observing_scenarios_paper/aligo_O4high.txt'), ('L1', 'L1', 'observing_scenarios_paper/aligo_O4high.txt'),
               ('Virgo', 'Virgo', 'observing_scenarios_paper/avirgo_O4high_NEW.txt'),
               ('KAGRA', 'KAGRA', 'observing_scenarios_paper/kagra_80Mpc.txt')],
    # next-generation layouts with log-uniform PSD tables (they take the shape-specialised kernels): three / four L-shaped detectors,
    # two triangles (a shape without its own instantiation: run-time loop bounds)
    '2L-ET+CE': [('ETSL', 'ETSL', 'ET-0000A-18.txt'), ('ETMRL', 'ETMRLpar', 'ET-0000A-18.txt'), ('CE1Id', 'CE1Id', 'ce_strain/cosmic_explorer.txt')],
    '2L-ET+2CE': [('ETSL', 'ETSL', 'ET-0000A-18.txt'), ('ETMRL', 'ETMRLpar', 'ET-0000A-18.txt'), ('CE1Id', 'CE1Id', 'ce_strain/cosmic_explorer.txt'),
                  ('CE2NSW', 'CE2NSW', 'ce_strain/cosmic_explorer_20km.txt')],
    '2ET': [('ETS', 'ETS', 'ET-0000A-18.txt'), ('ETMR', 'ETMR', 'ET-0000A-18.txt')],
}


def build_network(signal_cls, wf_model, name, useEarthMotion=True, fmin=2., psd_root=None, **kw):
    """{'det': GWSignal} for one of NETWORKS, built with any GWSignal-compatible class (engine, oracle port, reference)."""
    root = psd_root or glob.detPath
    out = {}
    for key, site, rel in NETWORKS[name]:
        s = glob.detectors[site]
        out[key] = signal_cls(wf_model, psd_path=os.path.join(root, rel), detector_shape=s['shape'], det_lat=s['lat'], det_long=s['long'],
                              det_xax=s['xax'], verbose=False, useEarthMotion=useEarthMotion, fmin=fmin, is_ASD=True, **kw)
    return out
