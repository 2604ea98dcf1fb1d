"""How crowded does the mating neighbourhood get as the c2 population evolves?  (diagnostic)"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from geonomics_b200 import workloads
from geonomics_b200.device import DeviceSpecies
cfg = dict(workloads.CONFIGS['c2'])
w = workloads.build(cfg, cfg['seed'])
N0, L = cfg['N'], w['L']
dev = DeviceSpecies(w['land_dim'], w['rasters'], w['prm'], w['gen_arch'], capacity=int(1.5 * N0) + 4096, seed=cfg['seed'])
dev.upload(w['pop']['x'], w['pop']['y'], w['pop']['age'], w['pop']['sex'], w['pop']['idx'],
           genomes_packed=workloads.random_packed_genomes(N0, L, cfg['seed'] + 1))
dev.set_debug(True)
for steps in (5, 100, 400, 500):
    dev.step(steps)
    dev.sync()
    c = dev.counters()
    n = c['n']
    nn = dev.read('N_NBRS', n)
    cs = dev.read('CELL_START', 512 * 512 + 1).astype(np.int64)
    cnt = np.diff(cs)
    rows3 = cnt.reshape(512, 512)
    row_run = rows3[:, :-2] + rows3[:, 1:-1] + rows3[:, 2:]
    print('t=%d n=%d nbrs mean %.1f p99 %d max %d | cell count mean %.2f max %d | 3-cell row run p99 %d max %d frac>32 %.4f'
          % (c['t'], n, nn.mean(), np.percentile(nn, 99), nn.max(), cnt.mean(), cnt.max(),
             np.percentile(row_run, 99), row_run.max(), (row_run > 32).mean()))
dev.close()
