"""oracle/ -- TEST INFRASTRUCTURE ONLY.

CPU checker for the B200 Fisher/SNR path.  Only ``tests/``, ``__graft_entry__.smoke()``
and the ``cpu_baseline`` / ``--impl reference`` legs of ``bench.py`` may import anything
from here; ``gwfast_b200`` never does.

* ``oracle.dual``      forward-mode dual ndarray
* ``oracle.refshim``   fake ``jax``/``h5py``/``numdifftools`` so the reference's own code
                       (``/root/reference/gwfast``) runs unmodified in-container
* ``oracle.reference`` loader for the reference under the shim (container only)
* ``oracle.nsbh_table`` IMRPhenomNSBH's xi_tide table as the reference's own ``_tabulate_xiTide`` produces it (cached under ``oracle/_ref``)
* ``oracle.port``      numpy restatement of the hot path that travels to the GPU box,
                       pinned against the reference by ``tests/golden`` fixtures
"""
