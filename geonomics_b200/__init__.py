"""geonomics_b200 -- B200-native (sm_100a) per-timestep update loop for Geonomics models.

Layout:
  csrc/          CUDA kernels + the C-ABI (libgnxb200.so, declared in include/gnx_b200.h)
  _lib.py        ctypes binding of the C-ABI (no CPU fallback)
  device.py      DeviceSpecies: HBM-resident structure-of-arrays state of one Species
  density.py     setup of the density-grid stack + its triangulation
  genome_pack.py bit-packing of genotypes / recombination paths
  build.py       in-tree nvcc build of the shared library
"""
from .build import build_library, library_path  # noqa: F401

__version__ = '0.1.0'
