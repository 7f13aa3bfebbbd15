"""h5py stub (TEST INFRASTRUCTURE): the reference imports h5py for catalog IO (out of scope) and for IMRPhenomNSBH's xi_tide table, which
the oracle obtains from the reference's own tabulation instead (oracle/nsbh_table.py) -- the file is never opened."""


class File:
    def __init__(self, *a, **k):
        raise ImportError("h5py is not available in this container (oracle shim stub)")
