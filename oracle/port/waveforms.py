"""numpy restatement of the reference waveform models on the hot path (TEST INFRASTRUCTURE).

Every function works on plain ndarrays or on ``oracle.dual.Dual`` operands (numpy protocol
dispatch), so the same code gives values and forward-mode tangents.  It is pinned against
the reference's own code by ``tests/test_oracle_port.py`` (container) and the committed
``tests/golden`` fixtures (anywhere).  The product package never imports it.

Restated here (reference file:line):
  * TaylorF2_RestrictedPN   gwfast/waveforms.py:697-953
  * IMRPhenomD              gwfast/waveforms.py:959-1333
  * IMRPhenomD_NRTidalv2    gwfast/waveforms.py:1339-1832
  * IMRPhenomHM             gwfast/waveforms.py:1838-2749   (oracle/port/phenomhm.py)
The fit tables are those of arXiv:1508.07253 Tab. 5 as typed in the reference.
"""
import os
import numpy as np

from .constants import GMSUN_C3, GMSUN_C2_GPC, C_GPC_S, F_ISCO_COEFF

_HERE = os.path.dirname(os.path.abspath(__file__))
PI = np.pi
EULER = np.euler_gamma


def _seta(eta):
    # waveforms.py:759 (and everywhere): sqrt(where(eta<0.25, 1-4 eta, 0)); at eta >= 0.25 value 0 and tangent 0 (oracle/dual.py:_sqrt)
    return np.sqrt(np.where(eta < 0.25, 1.0 - 4.0 * eta, 0.))


def quad_mon(Lam):
    """spin-induced quadrupole from the tidal deformability; waveforms.py:779, 1394."""
    lg = np.log(np.where(Lam < 1., 1., Lam))
    hi = np.exp(0.1940 + 0.09163 * lg + 0.04812 * lg * lg - 4.283e-3 * lg * lg * lg + 1.245e-4 * lg * lg * lg * lg)
    lo = 1. + Lam * (0.427688866723244 + Lam * (-0.324336526985068 + Lam * 0.1107439432180572))
    return np.where(Lam < 1., lo, hi)


def oct_mon_minus1(qm):
    """spin-induced octupole minus the BBH baseline; waveforms.py:1564."""
    lq = np.log(qm)
    return -1. + np.exp(0.003131 + 2.071 * lq - 0.7152 * lq * lq + 0.2458 * lq * lq * lq - 0.03309 * lq * lq * lq * lq)


def lamt_dellam_from_lam12(L1, L2, eta):
    """gwfastUtils.py:398-417."""
    e2 = eta * eta
    s = _seta(eta)
    lt = (8. / 13.) * ((1. + 7. * eta - 31. * e2) * (L1 + L2) + s * (1. + 9. * eta - 11. * e2) * (L1 - L2))
    dl = 0.5 * (s * (1. - 13272. / 1319. * eta + 8944. / 1319. * e2) * (L1 + L2)
                + (1. - 15910. / 1319. * eta + 32850. / 1319. * e2 + 3380. / 1319. * e2 * eta) * (L1 - L2))
    return lt, dl


def lam12_from_lamt_dellam(lt, dl, eta):
    """gwfastUtils.py:419-448 (2x2 inverse of the map above, with the closed-form determinant)."""
    e2 = eta * eta
    s = _seta(eta)
    a = (8. / 13.) * (1. + 7. * eta - 31. * e2)
    b = (8. / 13.) * s * (1. + 9. * eta - 11. * e2)
    c = s * (1. - (13272. / 1319.) * eta + (8944. / 1319.) * e2) * 0.5
    d = (1. - (15910. / 1319.) * eta + (32850. / 1319.) * e2 + (3380. / 1319.) * (e2 * eta)) * 0.5
    det = (306656. / 1319.) * (eta ** 5) - (5936. / 1319.) * (eta ** 4)
    return ((c - d) * lt + (b - a) * dl) / det, ((-d - c) * lt + (b + a) * dl) / det


def mceta_from_m1m2(m1, m2):
    """gwfastUtils.py:466-479."""
    return ((m1 * m2) ** (3. / 5.)) / ((m1 + m2) ** (1. / 5.)), (m1 * m2) / ((m1 + m2) * (m1 + m2))


def m1m2_from_mceta(Mc, eta):
    """gwfastUtils.py:450-464."""
    s = _seta(eta)
    M = Mc / (eta ** (3. / 5.))
    return 0.5 * M * (1. + s), 0.5 * M * (1. - s)


# ----------------------------------------------------------------------------- PN pieces
def pn_phase_coeffs(eta, chi1, chi2, qm1=1., qm2=1., spin_ho_3p5=False):
    """3.5PN TaylorF2 phasing coefficients c_k (coefficient of v^k); waveforms.py:784-812 / 1059-1077.

    Returns a dict with c2..c7, the two log coefficients, and ``ss6`` = the 3PN spin-spin block
    that IMRPhenomD removes again (waveforms.py:1077).
    """
    e2 = eta * eta
    s = _seta(eta)
    m1, m2 = 0.5 * (1.0 + s), 0.5 * (1.0 - s)
    c12, c22, c1c2 = chi1 * chi1, chi2 * chi2, chi1 * chi2
    xs, xa = 0.5 * (chi1 + chi2), 0.5 * (chi1 - chi2)
    c = {}
    c['c2'] = 3715. / 756. + (55. * eta) / 9.
    c['c3'] = -16. * PI + (113. * s * xa) / 3. + (113. / 3. - (76. * eta) / 3.) * xs
    c['c4'] = (5. * (3058.673 / 7.056 + 5429. / 7. * eta + 617. * e2) / 72. + 247. / 4.8 * eta * c1c2 - 721. / 4.8 * eta * c1c2
               + (-720. / 9.6 * qm1 + 1. / 9.6) * m1 * m1 * c12 + (-720. / 9.6 * qm2 + 1. / 9.6) * m2 * m2 * c22
               + (240. / 9.6 * qm1 - 7. / 9.6) * m1 * m1 * c12 + (240. / 9.6 * qm2 - 7. / 9.6) * m2 * m2 * c22)
    t5 = (732985. / 2268. - 24260. * eta / 81. - 340. * e2 / 9.) * xs + (732985. / 2268. + 140. * eta / 9.) * s * xa
    c['c5'] = 38645. * PI / 756. - 65. * PI * eta / 9. - t5
    c['c5l'] = c['c5'] * 3.

    def ss_block(q1, q2):
        return ((326.75 / 1.12 + 557.5 / 1.8 * eta) * eta * c1c2
                + (4703.5 / 8.4 + 2935. / 6. * m1 - 120. * m1 * m1) * m1 * m1 * q1 * c12
                + (-4108.25 / 6.72 - 108.5 / 1.2 * m1 + 125.5 / 3.6 * m1 * m1) * m1 * m1 * c12
                + (4703.5 / 8.4 + 2935. / 6. * m2 - 120. * m2 * m2) * m2 * m2 * q2 * c22
                + (-4108.25 / 6.72 - 108.5 / 1.2 * m2 + 125.5 / 3.6 * m2 * m2) * m2 * m2 * c22)

    c['c6'] = (11583.231236531 / 4.694215680 - 640. / 3. * PI * PI - 684.8 / 2.1 * EULER
               + eta * (-15737.765635 / 3.048192 + 225.5 / 1.2 * PI * PI) + e2 * 76.055 / 1.728 - e2 * eta * 127.825 / 1.296
               - np.log(4.) * 684.8 / 2.1 + PI * chi1 * m1 * (1490. / 3. + m1 * 260.) + PI * chi2 * m2 * (1490. / 3. + m2 * 260.)
               + ss_block(qm1, qm2))
    c['ss6'] = ss_block(1., 1.)
    c['c6l'] = -6848. / 21.
    c7 = (77096675. * PI / 254016. + 378515. * PI * eta / 1512. - 74045. * PI * e2 / 756.)
    lin_s = -25150083775. / 3048192. + 10566655595. * eta / 762048. - 1042165. * e2 / 3024. + 5345. * e2 * eta / 36.
    lin_a = -25150083775. / 3048192. + 26804935. * eta / 6048. - 1985. * e2 / 48.
    if spin_ho_3p5:
        xs2, xa2 = xs * xs, xa * xa
        c['c7'] = (c7 + (lin_s + (14585. / 8. - 7270. * eta + 80. * e2) * xa2) * xs + (14585. / 24. - 475. * eta / 6. + 100. * e2 / 3.) * xs2 * xs
                   + s * (lin_a * xa + (14585. / 24. - 2380. * eta) * xa2 * xa + (14585. / 8. - 215. * eta / 2.) * xa * xs2))
    else:
        c['c7'] = c7 + lin_s * xs + s * (lin_a * xa)
    return c


def tau_star(f, Mc, eta):
    """3.5PN time to coalescence in seconds; waveforms.py:878-901 (identical at :1299, :1769, :2712)."""
    Ms = Mc * GMSUN_C3 / (eta ** (3. / 5.))
    v = (PI * Ms * f) ** (1. / 3.)
    e2 = eta * eta
    fac = 5. / 256 * Ms / (eta * (v ** 8.))
    t05 = (1. + (743. / 252. + 11. / 3. * eta) * (v * v) - 32. / 5. * PI * (v * v * v)
           + (3058673. / 508032. + 5429. / 504. * eta + 617. / 72. * e2) * (v ** 4) - (7729. / 252. - 13. / 3. * eta) * PI * (v ** 5))
    t6 = (-10052469856691. / 23471078400. + 128. / 3. * PI * PI + 6848. / 105. * EULER + (3147553127. / 3048192. - 451. / 12. * PI * PI) * eta
          - 15211. / 1728. * e2 + 25565. / 1296. * e2 * eta + 3424. / 105. * np.log(16. * v * v)) * (v ** 6)
    t7 = (-15419335. / 127008. - 75703. / 756. * eta + 14809. / 378. * e2) * PI * (v ** 7)
    return fac * (t05 + t6 + t7)


# ----------------------------------------------------------------------------- models
class _Model:
    """parameter-ordering contract of WaveFormModel; waveforms.py:68-147."""
    is_tidal = False
    is_HigherModes = False
    is_holomorphic = False
    objType = 'BBH'

    def __init__(self, is_chi1chi2=True):
        self.is_chi1chi2 = is_chi1chi2
        names = ['Mc', 'eta', 'dL', 'theta', 'phi', 'iota', 'psi', 'tcoal', 'Phicoal']
        names += ['chi1z', 'chi2z'] if is_chi1chi2 else ['chiS', 'chiA']
        if self.is_tidal:
            names += ['LambdaTilde', 'deltaLambda']
        self.ParNums = {n: i for i, n in enumerate(names)}
        self.nParams = len(names)

    def tau_star(self, f, **ev):
        return tau_star(f, ev['Mc'], ev['eta'])


class TaylorF2_RestrictedPN(_Model):
    is_holomorphic = True

    def __init__(self, fHigh=None, is_tidal=False, use_3p5PN_SpinHO=False, phiref_vlso=False,
                 which_ISCO='Schw', use_QuadMonTid=False, **kw):
        self.is_tidal = is_tidal
        self.objType = 'BNS' if is_tidal else 'BBH'
        self.fcutPar = F_ISCO_COEFF if fHigh is None else fHigh
        self.use_3p5PN_SpinHO, self.phiref_vlso = use_3p5PN_SpinHO, phiref_vlso
        self.which_ISCO, self.use_QuadMonTid = which_ISCO, use_QuadMonTid
        super().__init__(**kw)

    def Phi(self, f, **ev):
        """waveforms.py:743-862."""
        eta, chi1, chi2 = ev['eta'], ev['chi1z'], ev['chi2z']
        Ms = ev['Mc'] * GMSUN_C3 / (eta ** (3. / 5.))
        v = (PI * Ms * f) ** (1. / 3.)
        if self.is_tidal and self.use_QuadMonTid:
            q1, q2 = quad_mon(ev['Lambda1']), quad_mon(ev['Lambda2'])
        else:
            q1 = q2 = 1.
        c = pn_phase_coeffs(eta, chi1, chi2, q1, q2, self.use_3p5PN_SpinHO)
        if self.phiref_vlso:
            c5 = c['c5'] * (1. - 3. * np.log(1. / np.sqrt(6.)))
            phiR = 0.
        else:
            c5, phiR = c['c5'], PI
        tidal = 0.
        if self.is_tidal:
            lt, dl = lamt_dellam_from_lam12(ev['Lambda1'], ev['Lambda2'], eta)
            tidal = (-0.5 * 39. * lt) * (v ** 10.) + (-3115. / 64. * lt + 6595. / 364. * _seta(eta) * dl) * (v ** 12.)
        lv = np.log(v)
        series = (1. + c['c2'] * v * v + c['c3'] * v ** 3 + c['c4'] * v ** 4 + (c5 + c['c5l'] * lv) * v ** 5
                  + (c['c6'] + c['c6l'] * lv) * v ** 6 + c['c7'] * v ** 7 + tidal)
        return 3. / (128. * eta) * series / (v ** 5.) + phiR - PI * 0.25

    def Ampl(self, f, **ev):
        """waveforms.py:864-876 (Newtonian amplitude, clightGpc convention)."""
        return np.sqrt(5. / 24.) * (PI ** (-2. / 3.)) * C_GPC_S / ev['dL'] * (GMSUN_C3 * ev['Mc']) ** (5. / 6.) * (f ** (-7. / 6.))

    def fcut(self, **ev):
        """waveforms.py:903-953."""
        eta = ev['eta']
        if self.which_ISCO == 'Schw':
            return self.fcutPar / (ev['Mc'] / (eta ** (3. / 5.)))
        e2 = eta * eta
        Mtot = ev['Mc'] / (eta ** (3. / 5.))
        chi1, chi2 = ev['chi1z'], ev['chi2z']
        s_ = _seta(eta)
        m1, m2 = 0.5 * (1.0 + s_), 0.5 * (1.0 - s_)
        s = (m1 * m1 * chi1 + m2 * m2 * chi2) / (m1 * m1 + m2 * m2)
        atot = (chi1 + chi2 * (m2 / m1) * (m2 / m1)) / ((1. + m2 / m1) * (1. + m2 / m1))
        aeff = atot + 0.41616 * eta * (chi1 + chi2)

        def r_isco(chi):
            z1 = 1.0 + ((1.0 - chi * chi) ** (1. / 3.)) * ((1.0 + chi) ** (1. / 3.) + (1.0 - chi) ** (1. / 3.))
            z2 = np.sqrt(3.0 * chi * chi + z1 * z1)
            root = np.sqrt((3.0 - z1) * (3.0 + z1 + 2.0 * z2))
            return np.where(chi > 0., 3.0 + z2 - root, 3.0 + z2 + root)

        r = r_isco(aeff)
        e_ns = eta * (0.055974469826360077 + 0.5809510763115132 * eta - 0.9606726679372312 * e2 + 3.352411249771192 * e2 * eta)
        e_tot = (e_ns * (1. + (-0.0030302335878845507 - 2.0066110851351073 * eta + 7.7050567802399215 * e2) * s)) / \
                (1. + (-0.6714403054720589 - 1.4756929437702908 * eta + 7.304676214885011 * e2) * s)
        Mfin = Mtot * (1. - e_tot)
        L = 2. / (3. * np.sqrt(3.)) * (1. + 2. * np.sqrt(3. * r - 2.))
        E = np.sqrt(1. - 2. / (3. * r))
        chif = atot + eta * (L - 2. * atot * (E - 1.)) + (-3.821158961 - 1.2019 * aeff - 1.20764 * aeff * aeff) * e2 \
            + (3.79245 + 1.18385 * aeff + 4.90494 * aeff * aeff) * e2 * eta
        om = 1. / (((r_isco(chif)) ** (3. / 2.)) + chif)
        return om / (PI * Mfin * GMSUN_C3)


# ---- IMRPhenomD -------------------------------------------------------------------
# arXiv:1508.07253 Tab. 5 fit tables, one row per phenomenological coefficient:
# value = t0 + t1*eta + xi*(t2 + t3*eta + t4*eta^2) + xi^2*(t5 + t6*eta + t7*eta^2) + xi^3*(t8 + t9*eta + t10*eta^2)
PHENOMD_FITS = {
    'sigma1': (2096.551999295543, 1463.7493168261553, 1312.5493286098522, 18307.330017082117, -43534.1440746107, -833.2889543511114, 32047.31997183187, -108609.45037520859, 452.25136398112204, 8353.439546391714, -44531.3250037322),
    'sigma2': (-10114.056472621156, -44631.01109458185, -6541.308761668722, -266959.23419307504, 686328.3229317984, 3405.6372187679685, -437507.7208209015, 1631817.1307344697, -7462.648563007646, -114585.25177153319, 674402.4689098676),
    'sigma3': (22933.658273436497, 230960.00814979506, 14961.083974183695, 1194018.1342318142, -3104223.9693052764, -3038.166617199259, 1872032.2849093592, -7309145.012085539, 42738.22871475411, 467502.018616601, -3064853.498512499),
    'sigma4': (-14621.71522218357, -377812.8579387104, -9608.682631509726, -1710892.5257214056, 4332924.601416521, -22366.683262266528, -2501971.6386377467, 10274495.902259542, -85360.30079034246, -570025.3441737515, 4396844.346849777),
    'beta1': (97.89747327985583, -42.659730877489224, 153.48421037904913, -1417.0620760768954, 2752.8614143665027, 138.7406469558649, -1433.6585075135881, 2857.7418952430758, 41.025109467376126, -423.680737974639, 850.3594335657173),
    'beta2': (-3.282701958759534, -9.051384468245866, -12.415449742258042, 55.4716447709787, -106.05109938966335, -11.953044553690658, 76.80704618365418, -155.33172948098394, -3.4129261592393263, 25.572377569952536, -54.408036707740465),
    'beta3': (-2.5156429818799565e-05, 1.9750256942201327e-05, -1.8370671469295915e-05, 2.1886317041311973e-05, 8.250240316860033e-05, 7.157371250566708e-06, -5.5780000112270685e-05, 0.00019142082884072178, 5.447166261464217e-06, -3.220610095021982e-05, 7.974016714984341e-05),
    'alpha1': (43.31514709695348, 638.6332679188081, -32.85768747216059, 2415.8938269370315, -5766.875169379177, -61.85459307173841, 2953.967762459948, -8986.29057591497, -21.571435779762044, 981.2158224673428, -3239.5664895930286),
    'alpha2': (-0.07020209449091723, -0.16269798450687084, -0.1872514685185499, 1.138313650449945, -2.8334196304430046, -0.17137955686840617, 1.7197549338119527, -4.539717148261272, -0.049983437357548705, 0.6062072055948309, -1.682769616644546),
    'alpha3': (9.5988072383479, -397.05438595557433, 16.202126189517813, -1574.8286986717037, 3600.3410843831093, 27.092429659075467, -1786.482357315139, 5152.919378666511, 11.175710130033895, -577.7999423177481, 1808.730762932043),
    'alpha4': (-0.02989487384493607, 1.4022106448583738, -0.07356049468633846, 0.8337006542278661, 0.2240008282397391, -0.055202870001177226, 0.5667186343606578, 0.7186931973380503, -0.015507437354325743, 0.15750322779277187, 0.21076815715176228),
    'alpha5': (0.9974408278363099, -0.007884449714907203, -0.059046901195591035, 1.3958712396764088, -4.516631601676276, -0.05585343136869692, 1.7516580039343603, -5.990208965347804, -0.017945336522161195, 0.5965097794825992, -2.0608879367971804),
    'gamma1': (0.006927402739328343, 0.03020474290328911, 0.006308024337706171, -0.12074130661131138, 0.26271598905781324, 0.0034151773647198794, -0.10779338611188374, 0.27098966966891747, 0.0007374185938559283, -0.02749621038376281, 0.0733150789135702),
    'gamma2': (1.010344404799477, 0.0008993122007234548, 0.283949116804459, -4.049752962958005, 13.207828172665366, 0.10396278486805426, -7.025059158961947, 24.784892370130475, 0.03093202475605892, -2.6924023896851663, 9.609374464684983),
    'gamma3': (1.3081615607036106, -0.005537729694807678, -0.06782917938621007, -0.6689834970767117, 3.403147966134083, -0.05296577374411866, -0.9923793203111362, 4.820681208409587, -0.006134139870393713, -0.38429253308696365, 1.7561754421985984),
    'rho1': (3931.8979897196696, -17395.758706812805, 3132.375545898835, 343965.86092361377, -1216256.5819981997, -70698.00600428853, 1383907.177859705, -3966276.1890979446, -60017.52423652596, 803515.1181825735, -2091710.365941658),
    'rho2': (-40105.47653771657, 112253.0169706701, 23561.696065836168, -3476180.699403351, 11375936.70849482, 754313.1127166454, -13084760.44625268, 36444584.853928134, 596226.612472288, -7427790.1143564405, 18928977.514040343),
    'rho3': (83208.35471266537, -191237.7264145924, -210916.2454782992, 8717975.08352568, -26914942.420669552, -1988980.6527362722, 30888029.960154563, -83908702.79256162, -1453503.1953446497, 17063528.990822166, -42748659.731120914),
    'v2': (0.8149838730507785, 2.5747553517454658, 1.1610198035496786, -2.3627771785551537, 6.771038707057573, 0.7570782938606834, -2.7256896890432474, 7.1140380397149965, 0.1766934149293479, -0.7978690983168183, 2.1162391502005153),
}


def _fit(name, eta, e2, xi):
    t = PHENOMD_FITS[name]
    return t[0] + t[1] * eta + (t[2] + t[3] * eta + t[4] * e2 + (t[5] + t[6] * eta + t[7] * e2) * xi + (t[8] + t[9] * eta + t[10] * e2) * xi * xi) * xi


_QNM = None


def qnm_tables(wf_dir=None):
    """QNM ringdown tables (a, fring, fdamp), 1003 rows; waveforms.py:988-990.

    Read from the data directory shipped with the repo (``gwfast_b200/data/WFfiles``, verbatim copies of
    the reference's three data tables) so that the port also runs where the reference tree is absent.
    """
    global _QNM
    if _QNM is None:
        d = wf_dir or os.path.join(os.path.dirname(os.path.dirname(_HERE)), 'gwfast_b200', 'data', 'WFfiles')
        _QNM = tuple(np.loadtxt(os.path.join(d, 'QNMData_%s.txt' % k)) for k in ('a', 'fring', 'fdamp'))
    return _QNM


def final_spin(eta, chi1, chi2):
    """waveforms.py:1256-1276."""
    s_ = _seta(eta)
    m1, m2 = 0.5 * (1.0 + s_), 0.5 * (1.0 - s_)
    s = m1 * m1 * chi1 + m2 * m2 * chi2
    af1 = eta * (3.4641016151377544 - 4.399247300629289 * eta + 9.397292189321194 * eta * eta - 13.180949901606242 * eta * eta * eta)
    af2 = eta * (s * ((1.0 / eta - 0.0850917821418767 - 5.837029316602263 * eta) + (0.1014665242971878 - 2.0967746996832157 * eta) * s))
    af3 = eta * (s * ((-1.3546806617824356 + 4.108962025369336 * eta) * s * s + (-0.8676969352555539 + 2.064046835273906 * eta) * s * s * s))
    return af1 + af2 + af3


def radiated_energy(eta, chi1, chi2):
    """waveforms.py:1278-1297."""
    s_ = _seta(eta)
    m1, m2 = 0.5 * (1.0 + s_), 0.5 * (1.0 - s_)
    s = (m1 * m1 * chi1 + m2 * m2 * chi2) / (m1 * m1 + m2 * m2)
    e_ns = eta * (0.055974469826360077 + 0.5809510763115132 * eta - 0.9606726679372312 * eta * eta + 3.352411249771192 * eta * eta * eta)
    return (e_ns * (1. + (-0.0030302335878845507 - 2.0066110851351073 * eta + 7.7050567802399215 * eta * eta) * s)) / \
           (1. + (-0.6714403054720589 - 1.4756929437702908 * eta + 7.304676214885011 * eta * eta) * s)


PHI_JOIN_INS = 0.018      # waveforms.py:980
AMP_JOIN_INS = 0.014      # waveforms.py:978
MF_CUT = 0.2              # waveforms.py:982


class _PhenomDCore:
    """f-independent IMRPhenomD quantities for a batch of events; waveforms.py:1002-1136, 1165-1245."""

    def __init__(self, eta, chi1, chi2, qm1=1., qm2=1.):
        self.eta = eta
        e2 = eta * eta
        s = _seta(eta)
        xs, xa = 0.5 * (chi1 + chi2), 0.5 * (chi1 - chi2)
        xi = -1.0 + (xs * (1.0 - eta * 76.0 / 113.0) + s * xa)
        a, tr, td = qnm_tables()
        aeff = final_spin(eta, chi1, chi2)
        erad = radiated_energy(eta, chi1, chi2)
        self.fring = np.interp(aeff.real, a, tr) / (1.0 - erad)
        self.fdamp = np.interp(aeff.real, a, td) / (1.0 - erad)
        for k in PHENOMD_FITS:
            setattr(self, k, _fit(k, eta, e2, xi))
        # ---- phase (waveforms.py:1059-1136)
        c = pn_phase_coeffs(eta, chi1, chi2, qm1, qm2, False)
        c['c6'] = c['c6'] - c['ss6']
        self.pn = c
        self.norm = 3. / (128. * eta)
        fj = PHI_JOIN_INS
        self.C2Int = self.dphi_ins(fj) - self.dphi_int(fj)
        self.C1Int = self.phi_ins(fj) - self.phi_int_raw(fj) / eta - self.C2Int * fj
        fm = 0.5 * self.fring
        self.fMRDJoin = fm
        self.C2MRD = (self.C2Int + self.dphi_int(fm)) - self.dphi_mrd(fm)
        self.C1MRD = (self.phi_int_raw(fm) / eta + self.C1Int + self.C2Int * fm) - self.phi_mrd_raw(fm) / eta - self.C2MRD * fm
        g2, g3 = self.gamma2, self.gamma3
        root = np.sqrt(np.where(g2 >= 1.0, 0., 1.0 - g2 * g2))
        # the phase uses |.| in both branches (waveforms.py:1134), the amplitude only in the first (:1193)
        self.fpeak_amp = np.where(g2 >= 1.0, np.fabs(self.fring - (self.fdamp * g3) / g2), self.fring + (self.fdamp * (-1.0 + root) * g3) / g2)
        self.fpeak_phi = np.fabs(self.fpeak_amp)
        self.t0 = self.dphi_mrd(self.fpeak_phi)
        # ---- amplitude (waveforms.py:1204-1245)
        sp1 = 1.0 + s
        c12, c22 = chi1 * chi1, chi2 * chi2
        A = {}
        A[2] = ((-969. + 1804. * eta) * (PI ** (2. / 3.))) / 672.
        A[3] = ((chi1 * (81. * sp1 - 44. * eta) + chi2 * (81. - 81. * s - 44. * eta)) * PI) / 48.
        A[4] = ((-27312085.0 - 10287648. * c22 - 10287648. * c12 * sp1 + 10287648. * c22 * s
                 + 24. * (-1975055. + 857304. * c12 - 994896. * chi1 * chi2 + 857304. * c22) * eta + 35371056 * e2) * (PI ** (4. / 3.))) / 8.128512e6
        A[5] = ((PI ** (5. / 3.)) * (chi2 * (-285197. * (-1. + s) + 4. * (-91902. + 1579. * s) * eta - 35632. * e2)
                                     + chi1 * (285197. * sp1 - 4. * (91902. + 1579. * s) * eta - 35632. * e2) + 42840. * (-1.0 + 4. * eta) * PI)) / 32256.
        A[6] = -((PI ** 2.) * (-336. * (-3248849057.0 + 2943675504. * c12 - 3339284256. * chi1 * chi2 + 2943675504. * c22) * e2 - 324322727232. * e2 * eta
                               - 7. * (-177520268561. + 107414046432. * c22 + 107414046432. * c12 * sp1 - 107414046432. * c22 * s
                                       + 11087290368. * (chi1 + chi2 + chi1 * s - chi2 * s) * PI)
                               + 12. * eta * (-545384828789. - 176491177632. * chi1 * chi2 + 202603761360. * c22 + 77616. * c12 * (2610335. + 995766. * s)
                                              - 77287373856. * c22 * s + 5841690624. * (chi1 + chi2) * PI + 21384760320. * PI * PI))) / 6.0085960704e10
        A[7], A[8], A[9] = self.rho1, self.rho2, self.rho3
        self.A = A
        self.amp0 = np.sqrt(2.0 * eta / 3.0) * (PI ** (-1. / 6.))
        # intermediate amplitude: quartic through (f1,v1,d1), (f2,v2), (f3,v3,d3); waveforms.py:1199-1245
        f1, f3 = AMP_JOIN_INS, self.fpeak_amp
        f2 = f1 + 0.5 * (f3 - f1)
        v1, d1 = self.amp_ins(f1), self.damp_ins(f1)
        v3, d3 = self.amp_mrd(f3), self.damp_mrd(f3)
        self.f1, self.f2, self.f3 = f1, f2, f3
        # Newton form on the node sequence f1,f1,f2,f3,f3 (same quartic as the delta_0..4 of the reference)
        a12 = (self.v2 - v1) / (f2 - f1)
        a23 = (v3 - self.v2) / (f3 - f2)
        b112 = (a12 - d1) / (f2 - f1)
        b123 = (a23 - a12) / (f3 - f1)
        b233 = (d3 - a23) / (f3 - f2)
        c1123 = (b123 - b112) / (f3 - f1)
        c1233 = (b233 - b123) / (f3 - f1)
        self.nw = (v1, d1, b112, c1123, (c1233 - c1123) / (f3 - f1))

    # -- phase pieces; x is the dimensionless frequency Mf
    def phi_ins(self, x):
        """waveforms.py:1081-1094, 1145 (inspiral branch)."""
        c, n, eta = self.pn, self.norm, self.eta
        x13 = x ** (1. / 3.)
        lg = np.log(PI * x) / 3.
        pn = (c['c5'] * n + c['c7'] * n * (PI ** (2. / 3.)) * (x13 * x13) + c['c6'] * n * (PI ** (1. / 3.)) * x13
              + c['c6l'] * n * (PI ** (1. / 3.)) * x13 * lg + c['c5l'] * n * lg + c['c4'] * n * (PI ** (-1. / 3.)) / x13
              + c['c3'] * n * (PI ** (-2. / 3.)) / (x13 * x13) + c['c2'] * n / PI / x + n * (PI ** (-5. / 3.)) * x ** (-5. / 3.))
        return pn + (self.sigma1 * x + self.sigma2 * 0.75 * x ** (4. / 3.) + self.sigma3 * 0.6 * x ** (5. / 3.) + self.sigma4 * 0.5 * x * x) / eta

    def dphi_ins(self, x):
        """waveforms.py:1108-1109."""
        c, n, eta = self.pn, self.norm, self.eta
        px = PI * x
        d = (2.0 * c['c7'] * n * (px ** (7. / 3.)) + (c['c6'] * n + c['c6l'] * n * (1.0 + np.log(px) / 3.)) * (px ** 2.) + c['c5l'] * n * (px ** (5. / 3.))
             - c['c4'] * n * (px ** (4. / 3.)) - 2. * c['c3'] * n * px - 3. * c['c2'] * n * (px ** (2. / 3.)) - 5. * n) * PI / (3. * (px ** (8. / 3.)))
        return d + (self.sigma1 + self.sigma2 * (x ** (1. / 3.)) + self.sigma3 * (x ** (2. / 3.)) + self.sigma4 * x) / eta

    def phi_int_raw(self, x):
        return self.beta1 * x - self.beta3 / (3. * x * x * x) + self.beta2 * np.log(x)

    def dphi_int(self, x):
        return (self.beta1 + self.beta3 / (x ** 4) + self.beta2 / x) / self.eta

    def phi_mrd_raw(self, x):
        return -(self.alpha2 / x) + (4.0 / 3.0) * (self.alpha3 * (x ** (3. / 4.))) + self.alpha1 * x + self.alpha4 * np.arctan((x - self.alpha5 * self.fring) / self.fdamp)

    def dphi_mrd(self, x):
        u = x - self.alpha5 * self.fring
        return (self.alpha1 + self.alpha2 / (x * x) + self.alpha3 / (x ** (1. / 4.)) + self.alpha4 / (self.fdamp * (1. + u * u / (self.fdamp * self.fdamp)))) / self.eta

    def phase(self, x, xref, apply_fcut=True):
        """three-region phase minus the t0 / phiRef alignment; waveforms.py:1139-1153."""
        def regions(y, with_cut):
            mrd = self.phi_mrd_raw(y) / self.eta + self.C1MRD + self.C2MRD * y
            if with_cut:
                mrd = np.where(y < MF_CUT, mrd, 0.)
            return np.where(y < PHI_JOIN_INS, self.phi_ins(y),
                            np.where(y < self.fMRDJoin, self.phi_int_raw(y) / self.eta + self.C1Int + self.C2Int * y, mrd))
        phi_ref = regions(xref, apply_fcut)
        lin = -self.t0 * (x - xref) - phi_ref
        if apply_fcut:
            return regions(x, True), np.where(x < MF_CUT, 1., 0.), lin
        return regions(x, False), 1., lin

    # -- amplitude pieces
    def amp_ins(self, x):
        A = self.A
        x13 = x ** (1. / 3.)
        x23 = x13 * x13
        return 1. + x23 * A[2] + (x ** (4. / 3.)) * A[4] + (x ** (5. / 3.)) * A[5] + (x ** (7. / 3.)) * A[7] + (x ** (8. / 3.)) * A[8] \
            + x * (A[3] + x * A[6] + x * x * A[9])

    def damp_ins(self, x):
        """d/dx of amp_ins (waveforms.py:1217, written there as an explicit expression)."""
        A = self.A
        return (2. / 3.) * A[2] / (x ** (1. / 3.)) + A[3] + (4. / 3.) * A[4] * (x ** (1. / 3.)) + (5. / 3.) * A[5] * (x ** (2. / 3.)) + 2. * A[6] * x \
            + (7.0 / 3.0) * (x ** (4. / 3.)) * A[7] + (8.0 / 3.0) * (x ** (5. / 3.)) * A[8] + 3. * (x * x) * A[9]

    def amp_mrd(self, x):
        fd3 = self.fdamp * self.gamma3
        u = x - self.fring
        return np.exp(-u * self.gamma2 / fd3) * (fd3 * self.gamma1) / (u * u + fd3 * fd3)

    def damp_mrd(self, x):
        """waveforms.py:1221."""
        fd3 = self.fdamp * self.gamma3
        u = x - self.fring
        den = u * u + fd3 * fd3
        return ((-2. * self.fdamp * u * self.gamma3 * self.gamma1) / den - (self.gamma2 * self.gamma1)) / (np.exp(u * self.gamma2 / fd3) * den)

    def amp_int(self, x):
        v1, d1, b, c, d = self.nw
        u1 = x - self.f1
        return v1 + u1 * (d1 + u1 * (b + (x - self.f2) * (c + (x - self.f3) * d)))

    def amplitude(self, x, apply_fcut=True):
        mrd = self.amp_mrd(x)
        if apply_fcut:
            mrd = np.where(x < MF_CUT, mrd, 0.)
        return np.where(x < AMP_JOIN_INS, self.amp_ins(x), np.where(x < self.fpeak_amp, self.amp_int(x), mrd))


class IMRPhenomD(_Model):
    def __init__(self, fRef=None, apply_fcut=True, **kw):
        self.fRef, self.apply_fcut, self.fcutPar = fRef, apply_fcut, MF_CUT
        super().__init__(**kw)

    def _core(self, ev):
        return _PhenomDCore(ev['eta'], ev['chi1z'], ev['chi2z'])

    def _mf(self, f, ev):
        M = ev['Mc'] / (ev['eta'] ** (3. / 5.))
        x = M * GMSUN_C3 * f
        xref = np.amin(x, axis=0) if self.fRef is None else M * GMSUN_C3 * self.fRef
        return M, x, xref

    def Phi(self, f, **ev):
        """waveforms.py:992-1153."""
        core = self._core(ev)
        M, x, xref = self._mf(f, ev)
        phis, win, lin = core.phase(x, xref, self.apply_fcut)
        return phis + win * lin

    def Ampl(self, f, **ev):
        """waveforms.py:1155-1254."""
        core = self._core(ev)
        M, x, xref = self._mf(f, ev)
        overall = 2. * np.sqrt(5. / (64. * PI)) * M * GMSUN_C2_GPC * M * GMSUN_C3 / ev['dL']
        return overall * core.amp0 * (x ** (-7. / 6.)) * core.amplitude(x, self.apply_fcut)

    def fcut(self, **ev):
        """waveforms.py:1324-1333."""
        return self.fcutPar / (ev['Mc'] * GMSUN_C3 / (ev['eta'] ** (3. / 5.)))


# ---- IMRPhenomD_NRTidalv2 -----------------------------------------------------------
def _kappa2T(eta, L1, L2):
    """waveforms.py:1543 / 1679 / 1815."""
    s = _seta(eta)
    Xa, Xb = 0.5 * (1.0 + s), 0.5 * (1.0 - s)
    return (3.0 / 13.0) * ((1.0 + 12.0 * Xb / Xa) * (Xa ** 5) * L1 + (1.0 + 12.0 * Xa / Xb) * (Xb ** 5) * L2)


def _f_merger(eta, k2T):
    """dimensionless merger frequency of the Planck taper; waveforms.py:1690-1700 / 1818-1828."""
    s = _seta(eta)
    q = 0.5 * (1.0 + s - 2.0 * eta) / eta
    num = 1.0 + 3.35411203e-2 * k2T + 4.31460284e-5 * k2T * k2T
    den = 1.0 + 7.54224145e-2 * k2T + 2.23626859e-4 * k2T * k2T
    return (0.3586 / np.sqrt(q)) * (num / den) / (2. * PI)


def planck_taper(x, y, end_zero=False):
    """waveforms.py:1702-1722 incl. the custom JVP: zero tangent w.r.t. x, hand-written rule w.r.t. y.

    ``end_zero``: the documented carve-out of SURVEY.md App. A-3 -- the reference's last grid sample sits exactly on the
    end of the taper (x = 1.2 y) where ``where(x > yp, 0, ...)`` returns 0 or 1 depending on last-bit rounding; with
    ``end_zero`` the taper is 0 for x >= 1.2 y (1 - 1e-12), which is what the CUDA engine implements.
    """
    from ..dual import Dual
    xv = x.v if isinstance(x, Dual) else x
    yv = y.v if isinstance(y, Dual) else y
    a = 1.2
    yp = a * yv
    with np.errstate(all='ignore'):
        e = np.exp((yp - yv) / (xv - yv) + (yp - yv) / (xv - yp))
        val = np.nan_to_num(np.where(xv < yv, 1., np.where(xv > yp, 0., 1. - 1. / (e + 1.))))
        if end_zero:
            val = np.where(xv >= yp * (1. - 1e-12), 0., val)
        if not isinstance(y, Dual):
            return val
        dy = np.where(xv < yv, 0., np.where(xv > yp, 0., e * ((-1. + a) / (xv - yv) + (-1. + a) / (xv - yp) + (-yv + yp) / ((xv - yv) ** 2)
                                                               + 1.2 * (-yv + yp) / ((xv - yp) ** 2)) / ((e + 1.) ** 2)))
        dy = np.nan_to_num(dy)
        if end_zero:
            dy = np.where(xv >= yp * (1. - 1e-12), 0., dy)
    # the outer nan_to_num of the reference (waveforms.py:1722) also acts on the tangent
    return Dual(val, np.nan_to_num(dy[..., None] * y.d))


class IMRPhenomD_NRTidalv2(IMRPhenomD):
    is_tidal = True
    objType = 'BNS'

    def __init__(self, fRef=None, apply_fcut=True, taper_end_zero=False, **kw):
        self.taper_end_zero = taper_end_zero
        super().__init__(fRef=fRef, apply_fcut=apply_fcut, **kw)

    def _lams(self, ev, like):
        if 'Lambda1' in ev:
            return ev['Lambda1'], ev['Lambda2']
        z = np.zeros(np.shape(like))
        return z, z

    def _core(self, ev):
        if 'Lambda1' in ev:
            return _PhenomDCore(ev['eta'], ev['chi1z'], ev['chi2z'], quad_mon(ev['Lambda1']), quad_mon(ev['Lambda2']))
        return _PhenomDCore(ev['eta'], ev['chi1z'], ev['chi2z'])

    def Phi(self, f, **ev):
        """waveforms.py:1374-1572."""
        eta, chi1, chi2 = ev['eta'], ev['chi1z'], ev['chi2z']
        core = self._core(ev)
        M, x, xref = self._mf(f, ev)
        L1, L2 = self._lams(ev, eta)
        if 'Lambda1' in ev:
            q1, q2 = quad_mon(L1), quad_mon(L2)
        else:
            q1 = q2 = np.ones(np.shape(eta))
        s = _seta(eta)
        m1, m2 = 0.5 * (1.0 + s), 0.5 * (1.0 - s)
        k2T = _kappa2T(eta, L1, L2)
        px = PI * x
        p23, p43, p53 = px ** (2. / 3.), px ** (4. / 3.), px ** (5. / 3.)
        num = 1.0 + (-12.615214237993088 * p23) + (19.0537346970349 * px) + (-21.166863146081035 * p43) + (90.55082156324926 * p53) + (-60.25357801943598 * px * px)
        den = 1.0 + (-15.11120782773667 * p23) + (22.195327350624694 * px) + (8.064109635305156 * p43)
        tidal = -k2T * 2.4375 / (m1 * m2) * p53 * num / den
        c12, c22 = chi1 * chi1, chi2 * chi2
        o1, o2 = oct_mon_minus1(q1), oct_mon_minus1(q2)
        ss = -400. * PI * (q1 - 1.) * c12 * m1 * m1 - 400. * PI * (q2 - 1.) * c22 * m2 * m2
        sss = (10. * ((m1 * m1 + 308. / 3. * m1) * chi1 + (m2 * m2 - 89. / 3. * m2) * chi2) * (q1 - 1.) * m1 * m1 * c12
               + 10. * ((m2 * m2 + 308. / 3. * m2) * chi2 + (m1 * m1 - 89. / 3. * m1) * chi1) * (q2 - 1.) * m2 * m2 * c22
               - 440. * o1 * m1 * m1 * m1 * c12 * chi1 - 440. * o2 * m2 * m2 * m2 * c22 * chi2)
        phis, win, lin = core.phase(x, xref, self.apply_fcut)
        return phis + win * (lin + tidal + (ss + sss) * core.norm * p23)

    def Ampl(self, f, **ev):
        """waveforms.py:1574-1724."""
        eta = ev['eta']
        core = self._core(ev) if False else _PhenomDCore(eta, ev['chi1z'], ev['chi2z'])   # amplitude ignores QuadMon
        M, x, xref = self._mf(f, ev)
        L1, L2 = self._lams(ev, eta)
        k2T = _kappa2T(eta, L1, L2)
        overall = 2. * np.sqrt(5. / (64. * PI)) * M * GMSUN_C2_GPC * M * GMSUN_C3 / ev['dL']
        xt = (PI * x) ** (2. / 3.)
        poly = (1.0 + 4.157407407407407 * xt + 2519.111111111111 * (xt ** 2.89)) / (1. + 13477.8073677 * (xt ** 4))
        amp_tidal = -9.0 * k2T * (xt ** 3.25) * poly
        taper = planck_taper(x, _f_merger(eta, k2T), self.taper_end_zero)
        return overall * (core.amp0 * (x ** (-7. / 6.)) * core.amplitude(x, True) + 2 * np.sqrt(PI / 5.) * amp_tidal) * taper

    def fcut(self, **ev):
        """waveforms.py:1794-1832."""
        eta = ev['eta']
        M = ev['Mc'] / (eta ** (3. / 5.))
        L1, L2 = self._lams(ev, eta)
        return 1.2 * _f_merger(eta, _kappa2T(eta, L1, L2)) / (M * GMSUN_C3)
