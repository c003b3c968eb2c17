"""realvsr_b200 -- B200-native (sm_100a) implementation of RealVSR's EDVR hot path.

Layout:
  csrc/            CUDA kernels + the C ABI (include/rvsr_b200.h)
  _lib.py          ctypes binding (no fallback: raises if the library is missing)
  engine.py        Python handle on the C++ inference engine
  archs/           host-side mirror of the reference's codes/models/archs API
                   (EDVR_arch, arch_util, dcn.deform_conv) -- same class names, ctor
                   arguments and state_dict keys
  dist.py          window sharding across ranks (one process per GPU)
"""
__version__ = "0.1.0"
