"""Multi-process check of the strip decomposition (run under torchrun, one rank per GPU):
the ranks jointly advance one landscape through NcclStrips (records written into the peers'
buffers over NVLink through CUDA IPC, NCCL collectives as barriers); rank 0 also runs the same
population undecomposed and compares -- ids, positions, ages, genomes, phenotypes, fitness and
the step records must be identical.

  python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 \
      --master-port 29511 tools/strips_nccl_check.py [steps]
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'tests'))


def main():
    import torch
    import torch.distributed as dist
    from test_cuda_strips import _workload, _single
    from geonomics_b200 import strips
    steps = int(sys.argv[1]) if len(sys.argv) > 1 else 8
    rank, local = int(os.environ['RANK']), int(os.environ['LOCAL_RANK'])
    torch.cuda.set_device(local)
    dist.init_process_group('nccl', device_id=torch.device('cuda', local))
    world = dist.get_world_size()
    wl = _workload(surfaces=True, n=20000, dim=(128, 96), L=200, seed=21)
    dim, rasters, prm, ga, pop, g = wl
    st = strips.NcclStrips(dim, rasters, prm, ga, capacity=4 * len(pop['x']) // world + 4096, seed=77)
    st.upload_owned(pop['x'], pop['y'], pop['age'], pop['sex'], pop['idx'], g=g)
    st.step(steps)
    st.sync()
    mine = st.dev.download()
    recs = st.step_records_local()
    shares = [None] * world
    dist.all_gather_object(shares, (mine, recs))
    ok = True
    if rank == 0:
        got = strips.merge_states([s[0] for s in shares])
        got_recs = strips.merge_records([s[1] for s in shares])
        ref, ref_recs = _single(wl, steps, seed=77)
        ref = ref[-1]
        ok = np.array_equal(ref['idx'], got['idx'])
        for k in ('x', 'y', 'age', 'sex', 'g', 'z', 'fit'):
            ok = ok and np.array_equal(ref[k], got[k])
        ok = ok and all((a['Nt'], a['n_births'], a['n_deaths'], a['n_pairs']) ==
                        (b['Nt'], b['n_births'], b['n_deaths'], b['n_pairs']) for a, b in zip(ref_recs, got_recs))
        print('strips over %d GPUs (CUDA IPC peer writes; barriers and collectives by %s), %d steps: %s; N %d -> %d; per-rank shares %s'
              % (world, 'gnx_strip_barrier over peer memory' if st.device_barrier else 'NCCL', steps, 'IDENTICAL to the undecomposed run' if ok else 'MISMATCH', len(pop['x']),
                 len(got['idx']), [len(s[0]['idx']) for s in shares]), flush=True)
    st.close()
    dist.destroy_process_group()
    sys.exit(0 if ok else 1)


if __name__ == '__main__':
    main()
