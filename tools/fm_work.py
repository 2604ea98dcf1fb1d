"""Work distribution of the mate search: per mating cell, focals nf and candidates K of its 3x3 block.
Prints how the distance tests split between sparse and crowded cells for several thresholds. (diagnostic)"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from geonomics_b200 import workloads
from geonomics_b200.device import DeviceSpecies
name = sys.argv[1] if len(sys.argv) > 1 else 'c4'
times = [int(a) for a in sys.argv[2:]] or [300, 1000]
cfg = dict(workloads.CONFIGS[name])
w = workloads.build(cfg, cfg['seed'])
N0, L = cfg['N'], w['L']
dev = DeviceSpecies(w['land_dim'], w['rasters'], w['prm'], w['gen_arch'], capacity=int(1.5 * N0) + 4096, seed=cfg['seed'])
dev.upload(w['pop']['x'], w['pop']['y'], w['pop']['age'], w['pop']['sex'], w['pop']['idx'],
           genomes_packed=workloads.random_packed_genomes(N0, L, cfg['seed'] + 1))
cs = w['prm']['mating_radius'] * 1.0000001
ncx, ncy = int(w['land_dim'][0] / cs) + 1, int(w['land_dim'][1] / cs) + 1
done = 0
for t in times:
    dev.step(t - done)
    done = t
    dev.sync()
    dev.step_records()
    st = dev.download(genomes=False)
    cx = np.minimum((st['x'] / cs).astype(np.int64), ncx - 1)
    cy = np.minimum((st['y'] / cs).astype(np.int64), ncy - 1)
    cnt = np.bincount(cy * ncx + cx, minlength=ncx * ncy).reshape(ncy, ncx).astype(np.int64)
    pad = np.pad(cnt, 1)
    rows = pad[:, :-2] + pad[:, 1:-1] + pad[:, 2:]           # 3-cell row ranges
    K = rows[:-2] + rows[1:-1] + rows[2:]
    rmax = np.maximum(np.maximum(rows[:-2], rows[1:-1]), rows[2:])
    nf = cnt
    occ = nf > 0
    tests = nf * K
    print('t=%d n=%d cells occupied %d  mean nf %.1f  tests %.3g  (per focal %.1f)' % (t, len(st['x']), occ.sum(), nf[occ].mean(), tests.sum(), tests.sum() / nf.sum()))
    for thr in (256, 512, 1024, 2048, 4096, 8192):
        heavy = occ & ((rmax > 64) | (tests >= thr))
        items = np.ceil(nf[heavy] / 32).sum()
        nch = np.ceil(K[heavy] / 32)
        chunk_tests = (nf[heavy] * nch).sum()
        light = occ & ~heavy
        print('  thr %5d: heavy cells %8d focals %9d items %8d tests %.3g chunk-tests %.3g (x32 = %.3g, pad %.2f) | light focals %9d tests %.3g  max-lane tests/warp est %.3g'
              % (thr, heavy.sum(), nf[heavy].sum(), items, tests[heavy].sum(), chunk_tests, chunk_tests * 32,
                 chunk_tests * 32 / max(tests[heavy].sum(), 1), nf[light].sum(), tests[light].sum(), 0))
    # histogram of nf for heavy (thr 4096)
    heavy = occ & ((rmax > 64) | (tests >= 4096))
    print('  heavy nf percentiles', np.percentile(nf[heavy], [10, 50, 90, 99]), 'K percentiles', np.percentile(K[heavy], [10, 50, 90, 99]))
    light = occ & ~heavy
    print('  light nf percentiles', np.percentile(nf[light], [10, 50, 90, 99]), 'K percentiles', np.percentile(K[light], [10, 50, 90, 99]))
dev.close()
