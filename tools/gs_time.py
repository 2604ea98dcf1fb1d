import sys, os, numpy as np
sys.path.insert(0, os.getcwd()); sys.path.insert(0, os.path.join(os.getcwd(), 'tests'))
from geonomics_b200 import _lib
_lib.LIB_PATH = os.path.join(os.getcwd(), 'geonomics_b200', 'libgnxb200_timing.so')
from geonomics_b200 import workloads
from geonomics_b200.device import DeviceSpecies, _DT
import ctypes as C
cfg = workloads.CONFIGS['c2']
w = workloads.build(cfg, 1)
dev = DeviceSpecies(w['land_dim'], w['rasters'], w['prm'], w['gen_arch'], capacity=1600000, seed=1)
g = workloads.random_packed_genomes(cfg['N'], w['L'], 2)
dev.upload(w['pop']['x'], w['pop']['y'], w['pop']['age'], w['pop']['sex'], w['pop']['idx'], genomes_packed=g)
dev.step(5); dev.sync()
out = np.zeros(64, dtype=np.int64)
_lib.check(dev._L.gnx_read_field(dev._ctx, 31, out.ctypes.data_as(C.c_void_p), out.nbytes), 'read')
d = np.diff(out[:48])
print('phase cycles:', d.tolist())
print('preamble', out[61]-out[60], 'total', out[62]-out[60])
print('gs_iters', dev.counters()['gs_iters'])
