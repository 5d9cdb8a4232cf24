"""innfer_b200 -- B200-native (sm_100a) engine for the RRDB/ESRGAN hot path of victorca25/iNNfer.

Layout: ``csrc/`` CUDA kernels + C-ABI, ``_native`` ctypes binding, ``engine`` the network handle,
``architectures`` / ``utils`` / ``run`` the host-side mirror of the reference's Python API.
"""
__version__ = "0.1.0"
