"""Physical constants of the reference, digit for digit (TEST INFRASTRUCTURE).

Follows gwfast/gwfastGlobals.py:38-111.  The constants are not mutually consistent
(SURVEY.md App. A-17); each is kept exactly as the reference writes it.
"""
import numpy as np

GMSUN_C3 = 4.925491025543575903411922162094833998e-6      # s        gwfastGlobals.py:38
GMSUN_C2 = 1.476625061404649406193430731479084713e3       # m        gwfastGlobals.py:44
GPC_M = 3.085677581491367278913937957796471611e25         # m        gwfastGlobals.py:50
GMSUN_C2_GPC = GMSUN_C2 / GPC_M                           # Gpc      gwfastGlobals.py:68
R_EARTH_KM = 6371.00                                      # km       gwfastGlobals.py:75
C_KM_S = 2.99792458e5                                     # km/s     gwfastGlobals.py:105
C_GPC_S = C_KM_S / 3.0856778570831e+22                    # Gpc/s    gwfastGlobals.py:111
DAY_S = 3600. * 24.                                       # "day" used for tcoal, signal.py:446
F_ISCO_COEFF = 1. / (6. * np.pi * np.sqrt(6.) * GMSUN_C3)  # Hz*Msun  waveforms.py:731

# detector sites used by the tests / bench (gwfastGlobals.py:143-220)
SITES = {
    'L1': dict(lat=30.563, long=-90.774, xax=242.71636956358617, shape='L'),
    'H1': dict(lat=46.455, long=-119.408, xax=170.99924234706103, shape='L'),
    'Virgo': dict(lat=43.631, long=10.504, xax=115.56756342034298, shape='L'),
    'KAGRA': dict(lat=36.412, long=137.306, xax=15.396, shape='L'),
    'ETS': dict(lat=40. + 31. / 60., long=9. + 25. / 60., xax=0., shape='T'),
    'ETSL': dict(lat=40. + 31. / 60., long=9. + 25. / 60., xax=45., shape='L'),
    'CE1Id': dict(lat=43.827, long=-112.825, xax=-45., shape='L'),
    'CE2NM': dict(lat=33.160, long=-106.480, xax=-105., shape='L'),
}
