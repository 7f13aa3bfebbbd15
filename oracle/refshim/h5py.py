"""h5py stub: only the out-of-scope catalog IO / IMRPhenomNSBH table use it (TEST INFRASTRUCTURE)."""


class File:
    def __init__(self, *a, **k):
        raise ImportError("h5py is not available in this container (oracle shim stub)")
