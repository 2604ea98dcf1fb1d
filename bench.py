#!/usr/bin/env python
"""bench.py -- individual-generations/second of the per-timestep update loop on B200.

  python bench.py --gpus N --steps K --warmup W [--workload c2|c3|c4] [--impl reference]

A "step" is one full time step (model.py:603-667 queue for one species: age, movement,
mate search, births with recombination, density, logistic mortality) over the synthetic
population of the named workload.  `value` is whole-job throughput with the state resident
in HBM; `e2e` is the same metric through the C-ABI host-buffer call gnx_walk_host (host->
device copy of the population, one step, device->host copy back, every step).  With N>1
(torchrun, one rank per GPU) every rank advances its own replicate population of the same
workload -- the path shards by independent replicate iterations, so
there is no data-path collective and scaling is "weak".
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = 'individual_generations_per_second'
UNIT = 'individual-generations/s'


# ------------------------------------------------------------------------------------------
# algorithmic HBM bytes per kernel launch (DESIGN.md "Kernels")
#   n = individuals at step start, B = births, P = pairs, npre = n + B, W = bytes per packed
#   homologue, T = traits, cells = mating-grid cells, YX = landscape cells
# ------------------------------------------------------------------------------------------
def algorithmic_bytes(W, T, cells, YX):
    return {
        # entries walked = last step's n_pre (its dead are dropped here); x,y rw; id r; alive r; key w
        'k_move_key': lambda s: s['npre'] * (1 + 4) + s['n'] * (32 + 8),
        'scan_cells.reduce': lambda s: cells * 4,
        'scan_cells.apply': lambda s: cells * 8,
        'k_bucket': lambda s: s['npre'] * 4 + s['n'] * (8 + 16),           # key r; id r; bucket w
        # bucket r; whole record (x,y 16, fit 8, age 4, slot 4, sex 1, z 8T) r + w (id comes from the bucket); key w
        'k_regrid': lambda s: s['n'] * (16 + (33 + 8 * T) + (41 + 8 * T) + 4),
        'k_find_mates': lambda s: s['n'] * (16 + 4 + 8 + 4),            # x,y; key; id; mate w
        'scan_pairs.reduce': lambda s: s['n'] * 8,
        'scan_pairs.apply': lambda s: s['n'] * 8 + s['P'] * (8 + 32 + 16 + 12),
        'k_gametes': lambda s: s['B'] * (4 * W + 2 * W + 8 + 8 * T + 4 + 8 + 8),   # rows; keys; z w; pair; slots
        'k_newborns': lambda s: s['B'] * (16 + 29 + 4),                          # midpoint r; record w; pair
        'k_density_counts': lambda s: (s['npre'] + s['P']) * 16,
        'k_raster_N': lambda s: YX * 8,
        'k_raster_d': lambda s: YX * 24,
        'k_death': lambda s: s['npre'] * (16 + 8 * T + 8 + 8 * T + 8 + 4 + 8 + 1),
        'scan_mortality.reduce': lambda s: s['npre'] * 1,
        'scan_mortality.apply': lambda s: s['npre'] * 1 + (s['npre'] - s['deaths']) * 2 * (41 + 8 * T),
    }


class ClockSampler:
    """SM clock and throttle reasons sampled DURING the timed region.

    The timed region of the default workload is a few milliseconds, shorter than one period of
    `nvidia-smi -lms`, so the samples come from NVML in this process (nvidia_ml_py: one poll every ~0.5 ms
    from a thread; the stepping calls release the GIL) and only those taken between `mark_start()` and
    `mark_end()` -- the host times that bracket the K timed steps -- are used.  `nvidia-smi` is the fallback
    when NVML cannot be loaded."""
    Q = ('index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,'
         'clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,'
         'clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap')
    REASONS = ((0x8, 'hw_slowdown'), (0x40, 'hw_thermal_slowdown'), (0x20, 'sw_thermal_slowdown'),
               (0x4, 'sw_power_cap'))

    def __init__(self, gpu_index):
        self.gpu = gpu_index
        self.proc = None
        self.lines = []
        self.samples = []            # (host time, sm MHz, reasons bitmask)
        self.t0 = self.t1 = None
        self.nvml = None
        self.handle = None
        self.smax = None
        self._stop = False

    def _nvml_handle(self):
        import pynvml
        pynvml.nvmlInit()
        h = None
        try:
            import torch
            uuid = str(torch.cuda.get_device_properties(self.gpu).uuid)
            for cand in ('GPU-' + uuid, uuid):
                try:
                    h = pynvml.nvmlDeviceGetHandleByUUID(cand.encode())
                    break
                except Exception:
                    h = None
        except Exception:
            h = None
        if h is None:
            h = pynvml.nvmlDeviceGetHandleByIndex(self.gpu)
        return pynvml, h

    def start(self):
        try:
            self.nvml, self.handle = self._nvml_handle()
            self.smax = float(self.nvml.nvmlDeviceGetMaxClockInfo(self.handle, self.nvml.NVML_CLOCK_SM))
            self.thread = threading.Thread(target=self._poll, daemon=True)
            self.thread.start()
            return
        except Exception:
            self.nvml = None
        try:
            self.proc = subprocess.Popen(['nvidia-smi', '-i', str(self.gpu), '--query-gpu=' + self.Q,
                                          '--format=csv,noheader,nounits', '-lms', '20'],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _poll(self):
        n = self.nvml
        while not self._stop:
            try:
                mhz = float(n.nvmlDeviceGetClockInfo(self.handle, n.NVML_CLOCK_SM))
                why = int(n.nvmlDeviceGetCurrentClocksThrottleReasons(self.handle))
                self.samples.append((time.perf_counter(), mhz, why))
            except Exception:
                pass
            time.sleep(0.0005)

    def _read(self):
        for ln in self.proc.stdout:
            self.lines.append(ln.strip())

    def mark_start(self):
        self.t0 = time.perf_counter()

    def mark_end(self):
        self.t1 = time.perf_counter()

    def stop(self):
        if self.nvml is not None:
            self._stop = True
            self.thread.join(timeout=1)
            t0 = self.t0 if self.t0 is not None else -1e300
            t1 = self.t1 if self.t1 is not None else 1e300
            inside = [s for s in self.samples if t0 <= s[0] <= t1]
            used = inside or self.samples[-3:]
            reasons = set()
            for _, _, why in used:
                for bit, nm in self.REASONS:
                    if why & bit:
                        reasons.add(nm)
            return {'sm_mhz': float(np.median([s[1] for s in used])) if used else None, 'sm_max_mhz': self.smax,
                    'samples': len(inside), 'source': 'nvml, polled during the timed region',
                    'reasons': sorted(reasons)}
        if self.proc is None:
            return {'sm_mhz': None, 'sm_max_mhz': None, 'reasons': ['nvidia-smi unavailable']}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, smax, reasons = [], [], set()
        names = ['hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap']
        for ln in self.lines:
            f = [v.strip() for v in ln.split(',')]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1]))
                smax.append(float(f[2]))
            except ValueError:
                continue
            for k, nm in enumerate(names):
                if f[5 + k].lower().startswith('active'):
                    reasons.add(nm)
        return {'sm_mhz': float(np.median(sm)) if sm else None,
                'sm_max_mhz': float(max(smax)) if smax else None,
                'samples': len(sm), 'source': 'nvidia-smi -lms 20', 'reasons': sorted(reasons)}


def load_traffic(workload, scale):
    """Measured DRAM bytes per launch from the committed `ncu --set full` capture of this workload
    (profiles/r02_traffic.json, else r01_traffic.json; written by tools/make_traffic.py); {} if there is none."""
    if scale != 1.0:
        return {}
    for nm in ('r02_traffic.json', 'r01_traffic.json'):       # the newest capture that names this workload
        path = os.path.join(ROOT, 'profiles', nm)
        try:
            k = json.load(open(path)).get(workload, {}).get('kernels', {})
        except Exception:
            k = {}
        if k:
            return k
    return {}


def load_peaks():
    p = os.path.join(ROOT, 'MEASURED_PEAKS.json')
    if os.path.exists(p):
        try:
            return float(json.load(open(p))['hbm_gbs']), 'measured (MEASURED_PEAKS.json hbm_gbs)'
        except Exception:
            pass
    return 6650.0, 'fallback (B200_PROFILING.md 6.65 TB/s)'


# ------------------------------------------------------------------------------------------
# CPU arm: the numpy oracle (port of the reference algorithm) on a bounded sample
# ------------------------------------------------------------------------------------------
def _oracle_worker(args):
    cfg, n_sample, steps, seed = args
    os.environ.setdefault('OMP_NUM_THREADS', '1')
    from geonomics_b200 import workloads, genome_pack
    from oracle import step_oracle as so
    from oracle import draws as od
    c = workloads.scaled(cfg, n_sample)
    c['surfaces'] = False            # the reference's surface tables are O(Y*X*approx_len) setup
    w = workloads.build(c, seed)
    n = c['N']
    L = w['L']
    g = genome_pack.unpack_genomes(workloads.random_packed_genomes(n, L, seed + 1), L)
    prm = dict(w['prm'])
    arch = dict(land_dim=w['land_dim'], rasters=w['rasters'],
                K=w['rasters'][prm['K_layer']] * prm['K_factor'], ww=None,
                traits=w['gen_arch']['traits'], dom=None, paths=w['gen_arch']['paths'])
    z = so.phenotype(g, arch['traits'])
    state = dict(x=w['pop']['x'], y=w['pop']['y'], age=w['pop']['age'], sex=w['pop']['sex'],
                 idx=w['pop']['idx'], g=g, z=z, max_ind_idx=n - 1)
    dgs = so.DensityGridStack(arch['land_dim'], None)
    rng = np.random.default_rng(seed + 2)
    total = 0
    t_total = 0.0
    for s in range(steps + 1):
        n_now = len(state['x'])
        d = od.make_draws(rng, prm, n_now, n_now, len(arch['paths']))
        t0 = time.perf_counter()
        state, im = so.step(state, arch, prm, d, dgs=dgs)
        dt = time.perf_counter() - t0
        if s > 0:                     # first step warms caches / imports
            total += n_now
            t_total += dt
    return total, t_total


def cpu_baseline(cfg, n_sample, steps, procs, seed=123):
    if procs <= 1:
        res = [_oracle_worker((cfg, n_sample, steps, seed))]
        wall = res[0][1]
    else:
        import multiprocessing as mp
        ctx = mp.get_context('spawn')
        t0 = time.perf_counter()
        with ctx.Pool(procs) as pool:
            res = pool.map(_oracle_worker, [(cfg, n_sample, steps, seed + 10 * k) for k in range(procs)])
        wall = max(r[1] for r in res)
        del t0
    total = sum(r[0] for r in res)
    return total / wall, total, wall


def _reference_worker(args):
    """One process = one replica of the workload run by the UNMODIFIED reference (oracle/_ref, or
    /root/reference in the build container) through its own public API: make_model -> walk('burn')
    -> time walk(steps, 'main').  Only import-time dependencies that this image lacks are shimmed
    (oracle/ref_shims.py) and the burn-in length is fixed (its stationarity tests are burn-in control,
    outside the timed region)."""
    cfg, n_sample, steps, warmup, seed = args
    os.environ.setdefault('OMP_NUM_THREADS', '1')
    import contextlib
    import io
    import tempfile
    import warnings
    warnings.filterwarnings('ignore')
    from geonomics_b200 import workloads
    from oracle import ref_shims
    with contextlib.redirect_stdout(io.StringIO()):
        gnx = ref_shims.install()
    import geonomics.sim.burnin as _b
    _b._test_t_threshold = lambda *a, **k: True
    _b._test_adf_threshold = lambda *a, **k: True
    _b.SpatialTester.run_test = lambda self, n, alpha=0.05: True
    c = workloads.scaled(cfg, n_sample)
    X, Y = c['dim']
    tmp = tempfile.mkdtemp(prefix='gnx_ref_bench_')
    path = os.path.join(tmp, 'params.py')
    spp = [{'movement': True, 'movement_surface': bool(c['surfaces']), 'dispersal_surface': bool(c['surfaces']),
            'genomes': True, 'n_traits': c['n_traits'], 'demographic_change': 0, 'parameter_change': False}]
    with contextlib.redirect_stdout(io.StringIO()):
        gnx.make_parameters_file(path, layers=[{'type': 'defined', 'change': False} for _ in range(3)], species=spp)
    txt = open(path).read()
    open(path, 'w').write('import numpy as np\n' + txt)                     # params.py:180 quirk
    p = gnx.read_parameters_file(path)
    p['landscape']['main']['dim'] = (X, Y)
    lyr0 = workloads.smooth_field((X, Y), 7) if c['surfaces'] else np.ones((Y, X))
    rasts = [lyr0, np.tile(np.linspace(0, 1, X), (Y, 1)), np.tile(np.linspace(0, 1, Y)[:, None], (1, X))]
    for n, r in enumerate(rasts):
        p['landscape']['layers']['lyr_%i' % n]['init']['defined']['rast'] = r
    s = p['comm']['species']['spp_0']
    s['init'].update(N=c['N'], K_layer='lyr_0', K_factor=c['N'] / float(lyr0.sum()))
    s['mating'].update(sex=False, sex_ratio=1 / 1, b=c['b'], R=c['R'], n_births_distr_lambda=c['lam'],
                       n_births_fixed=True, mating_radius=c['mating_radius'], choose_nearest_mate=False,
                       inverse_dist_mating=False)
    s['mortality'].update(max_age=None, d_min=0, d_max=1)
    mv = s['movement']
    mv.update(direction_distr_mu=0, direction_distr_kappa=0, movement_distance_distr='wald',
              movement_distance_distr_param1=1.0, movement_distance_distr_param2=1.0,
              dispersal_distance_distr='wald', dispersal_distance_distr_param1=1.0,
              dispersal_distance_distr_param2=1.0)
    if c['surfaces']:
        for k in ('move_surf', 'disp_surf'):
            mv[k].update(layer='lyr_0', mixture=True, vm_distr_kappa=12, approx_len=200)
    g = s['gen_arch']
    g.update(L=c['L'], use_tskit=False, dom=False, n_recomb_sims=1000, r_distr_alpha=c['recomb_rate'],
             r_distr_beta=None, start_p_fixed=0.5, mu_neut=0, mu_delet=0)
    for t in range(c['n_traits']):
        g['traits']['trait_%i' % t].update(layer='lyr_%i' % (1 + t % 2), n_loci=c['loci_per_trait'], phi=c['phi'],
                                           gamma=c['gamma'], mu=0, alpha_distr_mu=0.0, alpha_distr_sigma=0.1,
                                           max_alpha_mag=0.25, univ_adv=False)
    p['model'].update(T=steps + warmup + 1, burn_T=6)
    p['model']['seed'] = {'num': seed}
    with contextlib.redirect_stdout(io.StringIO()):
        mod = gnx.make_model(p, name='ref_bench')
        mod.walk(10000, 'burn', verbose=False)
        mod.walk(max(1, warmup), 'main', verbose=False)
        spp0 = mod.comm[0]
        n_rec = len(spp0.Nt)
        N_before = len(spp0)
        t0 = time.perf_counter()
        mod.walk(steps, 'main', verbose=False)
        dt = time.perf_counter() - t0
    Nt = [N_before] + list(spp0.Nt[n_rec:])
    return float(sum(Nt[:-1])), dt                      # individuals at the start of every timed step


def run_reference(args, cfg):
    rank = int(os.environ.get('RANK', '0'))
    if rank != 0:
        return
    from oracle import ref_shims
    procs = os.cpu_count() or 1
    steps, warmup = max(1, args.steps), max(0, args.warmup)
    t0 = time.perf_counter()
    import multiprocessing as mp
    ctx = mp.get_context('spawn')
    have_ref = ref_shims.reference_root() is not None
    port_value, _, port_wall = cpu_baseline(cfg, args.cpu_sample, min(steps, 8), procs)
    if have_ref:
        n_sample = args.ref_sample
        with ctx.Pool(procs) as pool:
            res = pool.map(_reference_worker, [(cfg, n_sample, steps, warmup, 1000 + k) for k in range(procs)])
        total, wall = sum(r[0] for r in res), max(r[1] for r in res)
        value = total / wall
        kind = 'reference'
        sample = ('%d processes x %d timed steps (after burn-in and %d warm-up steps) of a %d-individual replica of '
                  'the workload (same per-capita parameters and density) run by the unmodified reference package '
                  '(erthward/geonomics 1.4.9, installed in oracle/_ref) through make_model / Model.walk; the '
                  'reference is single-threaded pure Python and its throughput is flat in N (BASELINE.md section '
                  '2), so one replica per host core is its best use of the box' % (procs, steps, warmup, n_sample))
    else:
        value, wall, kind, n_sample = port_value, port_wall, 'port', args.cpu_sample
        sample = ('%d processes x %d steps of a %d-individual replica, numpy oracle port of the reference '
                  'algorithm (reference package not installed in oracle/_ref)' % (procs, min(steps, 8), n_sample))
    elapsed = time.perf_counter() - t0
    line = {
        'impl': 'reference', 'metric': METRIC, 'value': value, 'unit': UNIT, 'n_gpus': args.gpus,
        'steps': args.steps, 'warmup': args.warmup, 'ms_per_step': 1e3 * wall / steps,
        'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f64 + int8 genotypes (numpy)',
        'data': 'synthetic',
        'config': {'workload': args.workload + ': ' + workload_desc(cfg), 'cpu_sample_individuals': n_sample},
        'cpu_baseline': {'value': value, 'unit': UNIT, 'cores': procs, 'kind': kind, 'sample': sample,
                         'port': {'value': port_value, 'unit': UNIT, 'cores': procs,
                                  'sample': 'numpy oracle port of the same algorithm, %d processes x %d steps of a '
                                            '%d-individual replica' % (procs, min(steps, 8), args.cpu_sample)}},
        'e2e': {'value': value, 'unit': UNIT, 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
        'gpu_launches': 0, 'wall_s': elapsed,
    }
    print(json.dumps(line))


def workload_desc(cfg):
    return '%d individuals, %d loci, %d traits x %d loci, %dx%d landscape, mating_radius %g%s' % (
        cfg['N'], cfg['L'], cfg['n_traits'], cfg['loci_per_trait'], cfg['dim'][0], cfg['dim'][1],
        cfg['mating_radius'], ', on-the-fly conductance surfaces' if cfg['surfaces'] else '')


def merge_dense(prof):
    """k_find_mates_dense serves the crowded cells of the same search: one row for both launches."""
    prof = dict(prof)
    if 'k_find_mates_dense' in prof and 'k_find_mates' in prof:
        a, b = prof.pop('k_find_mates_dense'), prof['k_find_mates']
        prof['k_find_mates'] = (b[0], a[1] + b[1], {'thread_per_focal_ms': b[1] / b[0], 'crowded_cells_ms': a[1] / a[0]})
    return prof


def kernel_table(dev, w, prof, precs, steps, workload, scale):
    """Per-kernel rows (CUDA-event time per launch, algorithmic bytes, fraction of the HBM peak)."""
    W = 4 * dev.W
    T = dev.n_traits
    cs = w['prm']['mating_radius'] * 1.0000001
    ncx, ncy = int(w['land_dim'][0] / cs) + 1, int(w['land_dim'][1] / cs) + 1
    cells = ncx * ncy
    YX = w['land_dim'][0] * w['land_dim'][1]
    ab = algorithmic_bytes(W, T, cells, YX)
    mean = {k: float(np.mean([r[k] for r in precs])) for k in ('Nt', 'n_births', 'n_deaths', 'n_pairs')}
    s = dict(n=mean['Nt'] - mean['n_births'] + mean['n_deaths'], B=mean['n_births'], P=mean['n_pairs'],
             deaths=mean['n_deaths'])
    s['npre'] = s['n'] + s['B']
    peak, peak_src = load_peaks()
    table = []
    traffic = load_traffic(workload, scale)
    prof = merge_dense(prof)
    kernel_sum_ms = sum(v[1] for v in prof.values()) / steps
    for name, val in sorted(prof.items(), key=lambda kv: -kv[1][1]):
        cnt, tot_ms = val[0], val[1]
        per_launch_ms = tot_ms / cnt
        row = {'kernel': name, 'launches_per_step': cnt / steps, 'ms_per_launch': per_launch_ms,
               'share_of_kernel_time_sum': tot_ms / steps / kernel_sum_ms}
        if name in ab:
            b = ab[name](s)
            row['algorithmic_bytes'] = b
            row['achieved_GBs'] = b / (per_launch_ms * 1e-3) / 1e9
            row['frac_of_hbm_peak'] = row['achieved_GBs'] / peak
        if name in traffic:
            row['dram_bytes_ncu'] = traffic[name]['dram_bytes_per_launch']
        if len(val) > 2:
            row['launches'] = val[2]
        table.append(row)
    return table, s, kernel_sum_ms, peak, peak_src


def roofline_block(table, s, W, peak, peak_src, kernel_sum_ms):
    top = next((r for r in table if 'achieved_GBs' in r), None)
    gam = next((r for r in table if r['kernel'] == 'k_gametes'), None)
    if top is None:
        return None
    roofline = {'bound': 'hbm', 'kernel': top['kernel'], 'achieved': top['achieved_GBs'], 'peak': peak,
                'unit': 'GB/s', 'frac': top['frac_of_hbm_peak'], 'traffic': top.get('dram_bytes_ncu'),
                'algorithmic_bytes': top['algorithmic_bytes'], 'peak_source': peak_src,
                'share_of_kernel_time_sum': top['share_of_kernel_time_sum'],
                'kernel_time_sum_ms_per_step': kernel_sum_ms,
                'note': 'share = this kernel / sum of the per-kernel CUDA-event times of the profiled pass '
                        '(serialised, events around every launch), not / ms_per_step of the graph-launched run'}
    if gam is not None:
        sec = gam['ms_per_launch'] * 1e-3
        g = {'kernel': 'k_gametes', 'achieved': gam.get('achieved_GBs'), 'frac': gam.get('frac_of_hbm_peak'),
             'ms_per_launch': gam['ms_per_launch'],
             # 4W parents + 2W child + 8 (two keys) per birth
             'frac_by_survey_8d_bytes': s['B'] * (6 * W + 8) / sec / 1e9 / peak,
             'bytes_per_birth_survey_8d': 6 * W + 8}
        if 'dram_bytes_ncu' in gam:
            g['frac_by_measured_dram'] = gam['dram_bytes_ncu'] / sec / 1e9 / peak
        roofline['genotype_streaming_kernel'] = g
    return roofline


def bind_near_gpu(local_rank):
    """Run this rank's host thread (and so its first-touch pinned buffers) on the cores of the NUMA
    node its GPU hangs off: the e2e leg moves ~180 MB per step per rank between host and device."""
    try:
        import pynvml
        pynvml.nvmlInit()
        import torch
        pr = torch.cuda.get_device_properties(local_rank)     # CUDA_VISIBLE_DEVICES may re-number: go by bus id
        h = pynvml.nvmlDeviceGetHandleByPciBusId(('%08x:%02x:%02x.0' % (pr.pci_domain_id, pr.pci_bus_id,
                                                                        pr.pci_device_id)).encode())
        words = pynvml.nvmlDeviceGetCpuAffinity(h, (os.cpu_count() + 63) // 64)
        cpus = {64 * i + b for i, wd in enumerate(words) for b in range(64) if (int(wd) >> b) & 1}
        cpus &= os.sched_getaffinity(0)
        if cpus:
            os.sched_setaffinity(0, cpus)
        return len(cpus)
    except Exception:
        return 0


def make_species(cfg, rank, seed_off=0):
    from geonomics_b200 import workloads
    from geonomics_b200.device import DeviceSpecies
    w = workloads.build(cfg, cfg['seed'] + rank + seed_off)
    N0, L = cfg['N'], w['L']
    cap = int(1.5 * N0) + 4096
    dev = DeviceSpecies(w['land_dim'], w['rasters'], w['prm'], w['gen_arch'], capacity=cap,
                        seed=cfg['seed'] + 7919 * rank + 104729 * (seed_off // 1000))
    dev.upload(w['pop']['x'], w['pop']['y'], w['pop']['age'], w['pop']['sex'], w['pop']['idx'],
               genomes_packed=workloads.random_packed_genomes(N0, L, cfg['seed'] + 1 + rank + seed_off))
    return dev, w, cap


def steady_state_block(name, presteps, steps, prof_steps):
    """One extra workload inside the default line: advanced `presteps` time steps first (evolved
    populations clump, so early steps flatter the mate search), then `steps` timed steps with the
    state resident, then a per-kernel pass."""
    import torch
    from geonomics_b200 import workloads
    cfg = dict(workloads.CONFIGS[name])
    dev, w, cap = make_species(cfg, 0)
    stream = torch.cuda.ExternalStream(dev.stream_ptr)
    done = 0
    while done < presteps:
        c = min(64, presteps - done)
        dev.step(c)
        done += c
    dev.sync()
    dev.step_records()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record(stream)
    dev.step(steps)
    e1.record(stream)
    dev.sync()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    recs = dev.step_records()
    ind = float(sum(r['Nt'] - r['n_births'] + r['n_deaths'] for r in recs))
    births = float(sum(r['n_births'] for r in recs))
    dev.profile(True)
    dev.step(prof_steps)
    dev.sync()
    prof = dev.profile_report()
    dev.profile(False)
    precs = dev.step_records()
    table, s, ksum, peak, peak_src = kernel_table(dev, w, prof, precs, prof_steps, name, 1.0)
    W = 4 * dev.W
    beta = births / ind
    per_ig = 193 + (6 * W + 8 + 29 + 16 + 32) * beta
    out = {'workload': name + ': ' + workload_desc(cfg), 'simulated_steps_before_timing': presteps,
           'steps': steps, 'ms_per_step': ms / steps, 'value': ind / (ms * 1e-3), 'unit': UNIT,
           'births_per_individual': beta,
           'whole_step': {'algorithmic_bytes_per_individual_generation': per_ig,
                          'achieved_GBs': per_ig * ind / (ms * 1e-3) / 1e9,
                          'frac_of_hbm_peak': per_ig * ind / (ms * 1e-3) / 1e9 / peak},
           'roofline': roofline_block(table, s, W, peak, peak_src, ksum), 'kernels': table[:12]}
    dev.close()
    return out


def strips_block(dist, rank, world, name, presteps, steps):
    """The north_star target config as ONE landscape strip-decomposed over the ranks (SURVEY.md
    section 8e-2): strong scaling -- the total work is fixed, every rank owns a band of rows.
    Records cross the strip edges by direct writes into the neighbour's buffer over NVLink peer
    memory; NCCL supplies the barriers and three small collectives (geonomics_b200/strips.py)."""
    import torch
    from geonomics_b200 import workloads, strips
    cfg = dict(workloads.CONFIGS[name])
    w = workloads.build(cfg, cfg['seed'])                      # every rank builds the same landscape / population
    N0, L = cfg['N'], w['L']
    cs, ncx, ncy = strips.mating_grid(w['land_dim'], w['prm']['mating_radius'])
    cap = int(2.0 * N0 / world) + 65536
    st = strips.NcclStrips(w['land_dim'], w['rasters'], w['prm'], w['gen_arch'], capacity=cap, seed=cfg['seed'],
                           migrant_capacity=max(65536, cap // 8), halo_capacity=max(65536, cap // 4))
    own = strips.owner_of(w['pop']['y'], st.bounds, st.cs, st.ncy) == rank
    n_own = int(own.sum())
    p = w['pop']
    st.dev.upload(p['x'][own], p['y'][own], p['age'][own], p['sex'][own], p['idx'][own],
                  genomes_packed=workloads.random_packed_genomes(n_own, L, cfg['seed'] + 1 + rank), max_ind_idx=N0 - 1)
    del w
    done = 0
    while done < presteps:
        c = min(50, presteps - done)
        st.step(c)
        st.sync()
        st.step_records_local()
        done += c
    dist.barrier()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(st.stream)
    st.step(steps)
    e1.record(st.stream)
    st.sync()
    dist.barrier()
    ms = e0.elapsed_time(e1)
    recs = st.step_records_local()
    ind = float(sum(r['Nt'] - r['n_births'] + r['n_deaths'] for r in recs))
    held = float(recs[-1]['Nt'])
    t = torch.tensor([ms, ind, held], device='cuda', dtype=torch.float64)
    tmax, tsum = t.clone(), t.clone()
    dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
    dist.all_reduce(tsum, op=dist.ReduceOp.SUM)
    sync_by_device = st.device_barrier
    st.close()
    ms_all, ind_all = float(tmax[0]), float(tsum[1])
    return {'workload': name + ': ' + workload_desc(cfg), 'scaling': 'strong', 'n_gpus': world,
            'decomposition': '%d horizontal strips of mating-grid rows (cut by carrying capacity), halo = one mating-grid '
                             'row, records by peer writes over NVLink (CUDA IPC); ' % world +
                             ('the 4 barriers and 3 small collectives of a step are one-CTA kernels over peer memory '
                              '(gnx_strip_barrier), no NCCL inside the step' if sync_by_device else
                              'NCCL barriers + 3 small collectives per step'),
            'simulated_steps_before_timing': presteps, 'steps': steps, 'ms_per_step': ms_all / steps,
            'value': ind_all / (ms_all * 1e-3), 'unit': UNIT,
            'load_imbalance_max_over_mean': float(tmax[2]) / (float(tsum[2]) / world)}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=None,
                    help='timed steps (default 1000 for the GPU arm: ~0.45 s, enough clock samples; 20 for --impl reference)')
    ap.add_argument('--warmup', type=int, default=5)
    ap.add_argument('--workload', default='c2')
    ap.add_argument('--impl', default='b200')
    ap.add_argument('--scale', type=float, default=1.0, help='shrink the workload (debug only)')
    ap.add_argument('--cpu-sample', type=int, default=100000)
    ap.add_argument('--cpu-steps', type=int, default=12)
    ap.add_argument('--ref-sample', type=int, default=1500,
                    help='individuals per replica of the unmodified reference (--impl reference; ~0.4 s per step per core)')
    ap.add_argument('--ref-steps', type=int, default=200,
                    help='timed steps of the unmodified reference in the cpu_baseline leg (~10 s of CPU work)')
    ap.add_argument('--no-cpu-baseline', action='store_true')
    ap.add_argument('--e2e-steps', type=int, default=5, help='host-buffer steps per replicate in flight')
    ap.add_argument('--e2e-depth', type=int, default=3, help='replicate populations in flight in the e2e leg')
    ap.add_argument('--presteps', type=int, default=0, help='time steps simulated before the warm-up')
    ap.add_argument('--c4-presteps', type=int, default=300,
                    help='the default line carries a c4 block (north_star target config) measured after this '
                         'many simulated steps; 0 disables it')
    ap.add_argument('--c4-steps', type=int, default=100)
    ap.add_argument('--no-strips', action='store_true',
                    help='with N > 1 ranks the default line also carries c4 strip-decomposed over the N GPUs '
                         '(strong scaling); this switches it off')
    ap.add_argument('--ncu-window', action='store_true',
                    help='bracket ONE extra step after the warm-up with cudaProfilerStart/Stop (for ncu --profile-from-start off)')
    ap.add_argument('--replicates', type=int, default=None,
                    help='replicate populations per GPU, stepped concurrently on their own streams '
                         '(default 8 for c3 = BASELINE configs[2]: 64 replicates on 8 GPUs; 1 otherwise)')
    args = ap.parse_args()
    if args.steps is None:
        args.steps = 20 if args.impl == 'reference' else 1000
    from geonomics_b200 import workloads
    cfg = dict(workloads.CONFIGS[args.workload])
    if args.scale != 1.0:
        cfg = workloads.scaled(cfg, int(cfg['N'] * args.scale))
    if args.impl == 'reference':
        run_reference(args, cfg)
        return

    import torch
    rank = int(os.environ.get('RANK', '0'))
    local_rank = int(os.environ.get('LOCAL_RANK', '0'))
    world = int(os.environ.get('WORLD_SIZE', '1'))
    torch.cuda.set_device(local_rank)
    if world > 1:
        bind_near_gpu(local_rank)
    dist = None
    if world > 1:
        # NCCL writes its version banner to STDOUT when NCCL_DEBUG is VERSION/WARN/INFO; the contract
        # is ONE JSON line on stdout, so file descriptor 1 points at stderr while the communicator
        # is created (init + a first collective), then is restored
        import torch.distributed as dist
        sys.stdout.flush()
        saved_fd = os.dup(1)
        os.dup2(2, 1)
        try:
            dist.init_process_group('nccl', device_id=torch.device('cuda', local_rank))
            t = torch.zeros(1, device='cuda')
            dist.all_reduce(t)
            torch.cuda.synchronize()
        finally:
            sys.stdout.flush()
            os.dup2(saved_fd, 1)
            os.close(saved_fd)

    from geonomics_b200.device import DeviceSpecies
    dev, w, cap = make_species(cfg, rank)
    N0, L = cfg['N'], w['L']
    stream = torch.cuda.ExternalStream(dev.stream_ptr)
    # further replicate populations on this GPU (independent iterations, model.py:115-117): each has
    # its own context and stream; the steps are enqueued round-robin and overlap on the device
    R = args.replicates if args.replicates is not None else (8 if args.workload == 'c3' else 1)
    devs, streams = [dev], [stream]
    for k in range(1, R):
        dk, _, _ = make_species(cfg, rank, 1000 * k)
        devs.append(dk)
        streams.append(torch.cuda.ExternalStream(dk.stream_ptr))

    def step_all(n):
        done = 0
        while done < n:                      # chunks keep every stream's launch queue fed
            c = min(16, n - done)
            for d in devs:
                d.step(c)
            done += c

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- optional: advance the simulation first (steady-state measurements)
    if args.presteps > 0:
        step_all(args.presteps)
    # ---- warm-up
    step_all(args.warmup)
    for d in devs:
        d.sync()
        d.step_records()
    if args.ncu_window:
        torch.cuda.synchronize()
        torch.cuda.profiler.start()
        step_all(1)
        torch.cuda.synchronize()
        torch.cuda.profiler.stop()
        for d in devs:
            d.step_records()
    launches0 = sum(d.launch_count for d in devs)
    # ---- timed region: K steps, state resident in HBM, CUDA events on the launching stream
    sampler = ClockSampler(local_rank)
    sampler.start()
    barrier()
    e0 = torch.cuda.Event(enable_timing=True)
    e1s = [torch.cuda.Event(enable_timing=True) for _ in devs]
    sampler.mark_start()
    e0.record(stream)                        # every stream is idle here (barrier above)
    step_all(args.steps)
    for e1, st in zip(e1s, streams):
        e1.record(st)
    for d in devs:
        d.sync()
    sampler.mark_end()
    barrier()
    ms = max(e0.elapsed_time(e1) for e1 in e1s)
    clocks = sampler.stop()
    launches = sum(d.launch_count for d in devs) - launches0
    # individuals processed per step = population at step start
    ind_gens = births = 0.0
    for d in devs:
        recs = d.step_records()
        assert len(recs) == args.steps, (len(recs), args.steps)
        ind_gens += float(sum(r['Nt'] - r['n_births'] + r['n_deaths'] for r in recs))
        births += float(sum(r['n_births'] for r in recs))
    from geonomics_b200 import parallel
    # job time = slowest rank's device time; job work = sum over ranks (no data-path collective)
    ms_all, ind_all = parallel.reduce_throughput(ms, ind_gens, dist, 'cuda')
    _, births_all = parallel.reduce_throughput(ms, births, dist, 'cuda')
    value = ind_all / (ms_all * 1e-3)

    # ---- per-kernel timing pass (CUDA events around every launch) for the roofline
    prof_steps = min(args.steps, 200)
    dev.profile(True)
    dev.step(prof_steps)
    dev.sync()
    prof = dev.profile_report()
    dev.profile(False)
    precs = dev.step_records()
    T = dev.n_traits
    YX = w['land_dim'][0] * w['land_dim'][1]
    table, s, kernel_sum_ms, peak, peak_src = kernel_table(dev, w, prof, precs, prof_steps, args.workload, args.scale)
    roofline = roofline_block(table, s, 4 * dev.W, peak, peak_src, kernel_sum_ms)

    # ---- e2e: host buffers through the C-ABI, every step (pinned host memory).  The populations are
    # independent replicate iterations (model.py:115-117), `--e2e-depth` of them in flight, each on its
    # own context: gnx_walk_host_begin enqueues one's copy in + step, gnx_walk_host_end waits and copies
    # it out, so one replicate's result travels to the host while the next one's input travels to the
    # device.  Every step still uploads its whole input and downloads its whole result.
    e2e = None
    if args.e2e_steps > 0:
        def pinned(shape, dtype):
            return torch.empty(shape, dtype=dtype, pin_memory=True).numpy()
        per_ind = 8 + 8 + 4 + 1 + 8 + 8 * dev.W + 8 * T + 8
        depth = max(1, args.e2e_depth)
        edevs = [dev]
        for k in range(1, depth):
            edevs.append(make_species(cfg, rank, 5000 * k)[0])
        ebufs = []
        for d in edevs:
            bufs = dict(x=pinned(cap, torch.float64), y=pinned(cap, torch.float64), age=pinned(cap, torch.int32),
                        sex=pinned(cap, torch.int8), idx=pinned(cap, torch.int64),
                        genomes=pinned((cap, 2, dev.W), torch.int32).view(np.uint32),
                        z=pinned((cap, max(1, T)), torch.float64), fit=pinned(cap, torch.float64))
            if d is not dev:
                d.step(args.warmup)
            st = d.download(unpack=False)
            n = len(st['x'])
            for k in ('x', 'y', 'age', 'sex', 'idx', 'fit'):
                bufs[k][:n] = st[k]
            bufs['genomes'][:n] = st['genomes']
            if T:
                bufs['z'].reshape(-1)[:n * T] = st['z'].reshape(-1)
            bufs['n'] = n
            bufs['max_ind_idx'] = st['max_ind_idx']
            d.walk_host(bufs, 1)                            # warm-up
            ebufs.append(bufs)
        n_calls = args.e2e_steps * depth
        barrier()
        t0 = time.perf_counter()
        e2e_ind = 0
        h2d = d2h = 0
        inflight = [False] * depth
        for k in range(n_calls):
            j = k % depth
            if inflight[j]:
                edevs[j].walk_host_end(ebufs[j])
                d2h += ebufs[j]['n'] * per_ind
            n_in = ebufs[j]['n']
            edevs[j].walk_host_begin(ebufs[j], 1)
            inflight[j] = True
            e2e_ind += n_in
            h2d += n_in * per_ind
        for j in range(depth):
            if inflight[j]:
                edevs[j].walk_host_end(ebufs[j])
                d2h += ebufs[j]['n'] * per_ind
        barrier()
        dt = time.perf_counter() - t0
        dt, e2e_ind = parallel.reduce_throughput(dt, float(e2e_ind), dist, 'cuda')
        e2e = {'value': e2e_ind / dt, 'unit': UNIT, 'h2d_bytes_per_step': h2d / n_calls,
               'd2h_bytes_per_step': d2h / n_calls, 'steps': n_calls, 'ms_per_step': dt * 1e3 / n_calls,
               'replicates_in_flight': depth,
               'api': 'gnx_walk_host_begin / gnx_walk_host_end (C-ABI, pinned host SoA buffers in and out every '
                      'step; %d replicate populations in flight, one context each)' % depth}
        for d in edevs:
            d.step_records()
        for d in edevs[1:]:
            d.close()

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        v, total, wall = cpu_baseline(cfg, args.cpu_sample, args.cpu_steps, 1)
        port = {'value': v, 'unit': UNIT, 'cores': 1, 'kind': 'port',
                'sample': '%d steps of a %d-individual replica of the workload (same per-capita parameters '
                          'and density), numpy oracle port of the reference algorithm, 1 process'
                          % (args.cpu_steps, args.cpu_sample), 'seconds': wall}
        cpu = port
        from oracle import ref_shims
        if ref_shims.reference_root() is not None:
            # the unmodified reference package itself (oracle/_ref), one process, in a child so that its
            # global numpy.random state and import shims stay out of this one
            import multiprocessing as mp
            with mp.get_context('spawn').Pool(1) as pool:
                tot_r, wall_r = pool.map(_reference_worker, [(cfg, args.ref_sample, args.ref_steps, 2, 1000)])[0]
            cpu = {'value': tot_r / wall_r, 'unit': UNIT, 'cores': 1, 'kind': 'reference',
                   'sample': '%d timed steps (after burn-in and 2 warm-up steps) of a %d-individual replica of the '
                             'workload (same per-capita parameters and density) run by the unmodified reference '
                             'package (erthward/geonomics 1.4.9, oracle/_ref) through make_model / Model.walk, '
                             '1 process (the reference is single-threaded)' % (args.ref_steps, args.ref_sample),
                   'seconds': wall_r, 'port': port}
    gs_iters = dev.counters()['gs_iters']

    for d in devs[1:]:
        d.close()
    dev.close()
    devs = []
    c4 = None
    if (rank == 0 and world == 1 and args.workload == 'c2' and args.scale == 1.0 and args.c4_presteps > 0
            and args.presteps == 0):
        # the north_star target config on one GPU, at steady state
        c4 = steady_state_block('c4', args.c4_presteps, args.c4_steps, min(50, args.c4_steps))

    c4_strips = None
    if (world > 1 and args.workload == 'c2' and args.scale == 1.0 and args.c4_presteps > 0 and args.presteps == 0
            and not args.no_strips):
        try:
            c4_strips = strips_block(dist, rank, world, 'c4', args.c4_presteps, args.c4_steps)
        except Exception as e:                          # e.g. no peer access between the GPUs of this box
            c4_strips = {'unavailable': '%s: %s' % (type(e).__name__, str(e)[:300])}

    if rank == 0:
        footprint_mb = (N0 * (2 * (8 + 8 + 4 + 1 + 8 + 4 + 8 * T + 8) + 8 * dev.W + 60) + 8 * YX * 6) / 1e6
        line = {
            'metric': METRIC, 'value': value, 'unit': UNIT, 'n_gpus': world, 'steps': args.steps,
            'warmup': args.warmup, 'ms_per_step': ms_all / args.steps, 'higher_is_better': True,
            'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f64 coordinates/phenotypes + u32 bit-packed genotypes',
            'data': 'synthetic',
            'config': {'workload': args.workload + ': ' + workload_desc(cfg),
                       'replicates': world * R,
                       'parallelism': '%d replicate population%s per GPU (own context and stream each), no '
                                      'collective' % (R, '' if R == 1 else 's'),
                       'births_per_individual': births_all / ind_all,
                       'simulated_steps_before_timing': args.presteps + args.warmup,
                       'l2': 'not flushed between steps: a step streams ~%.0f MB of state and work arrays '
                             '(> 126 MB L2)' % footprint_mb,
                       'rng': 'Philox4x32-10'},
            'clocks': clocks, 'e2e': e2e, 'gpu_launches': launches, 'roofline': roofline,
            'cpu_baseline': cpu, 'kernels': table[:14], 'gs_iters': gs_iters,
        }
        if c4 is not None:
            line['c4'] = c4
        if c4_strips is not None:
            line['c4_strips'] = c4_strips
        print(json.dumps(line))
    for d in devs:
        d.close()
    if dist is not None:
        dist.destroy_process_group()


if __name__ == '__main__':
    main()
