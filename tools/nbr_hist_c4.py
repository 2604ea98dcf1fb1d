"""How crowded does the mating neighbourhood get as the c4 population evolves?  (diagnostic)"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from geonomics_b200 import workloads
from geonomics_b200.device import DeviceSpecies
name = sys.argv[1] if len(sys.argv) > 1 else 'c4'
cfg = dict(workloads.CONFIGS[name])
w = workloads.build(cfg, cfg['seed'])
N0, L = cfg['N'], w['L']
dev = DeviceSpecies(w['land_dim'], w['rasters'], w['prm'], w['gen_arch'], capacity=int(1.5 * N0) + 4096, seed=cfg['seed'])
dev.upload(w['pop']['x'], w['pop']['y'], w['pop']['age'], w['pop']['sex'], w['pop']['idx'],
           genomes_packed=workloads.random_packed_genomes(N0, L, cfg['seed'] + 1))
dev.set_debug(True)
import torch, time
for steps in (20, 80, 200, 300, 400, 1000):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    dev.step(steps)
    dev.sync()
    dt = time.perf_counter() - t0
    c = dev.counters()
    n = c['n']
    nn = dev.read('N_NBRS', n)
    print('t=%d n=%d ms/step %.3f nbrs mean %.1f p50 %d p90 %d p99 %d p999 %d max %d ; sum(nbrs^1)/n %.1f'
          % (c['t'], n, 1e3 * dt / steps, nn.mean(), np.percentile(nn, 50), np.percentile(nn, 90), np.percentile(nn, 99),
             np.percentile(nn, 99.9), nn.max(), nn.mean()), flush=True)
    dev.step_records()
dev.close()
