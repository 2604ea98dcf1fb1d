"""GPU: gamete formation / phenotype across genome sizes, including the TMA-staged kernel
(rows >= 128 B, L > 384) and the register-streaming kernel, against the golden-pinned oracle;
and bit-equality of the two kernel variants."""
import numpy as np
import pytest

from parity_util import synthetic_case, run_device_step, compare_step, make_device

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize('L', [1, 64, 129, 300, 512, 1000, 2100, 4500])
def test_full_step_parity_across_genome_sizes(L):
    from oracle import step_oracle as so
    arch, prm, state, draws = synthetic_case(L=L, n=1200, loci_per_trait=min(20, max(1, L // 3)),
                                             n_traits=2 if L >= 4 else 1, seed=L, max_tries=24)
    new_o, im_o = so.step(state, arch, prm, draws)
    assert im_o['B'] > 100
    out = run_device_step(arch, prm, state, draws, staged=True)
    compare_step(out, new_o, im_o)


@pytest.mark.parametrize('L', [512, 1000, 2100])
def test_tma_and_register_kernels_agree(L):
    arch, prm, state, draws = synthetic_case(L=L, n=2500, seed=7 + L, max_tries=24)
    res = []
    for tma in (True, False):
        dev = make_device(arch, prm, capacity=6000, disp_tries=draws['disp_dist'].shape[1])
        try:
            dev.set_gamete_tma(tma)
            dev.upload(state['x'], state['y'], state['age'], state['sex'], state['idx'], g=state['g'],
                       z=state['z'], max_ind_idx=state['max_ind_idx'])
            d = dict(draws)
            d.pop('move_choice', None)
            d.pop('disp_choice', None)
            dev.set_draws(d)
            dev.step(1)
            dev.sync()
            res.append(dev.download())
        finally:
            dev.close()
    a, b = res
    assert len(a['x']) == len(b['x']) > 1000
    assert np.array_equal(a['idx'], b['idx'])
    assert np.array_equal(a['genomes'], b['genomes'])
    assert np.allclose(a['z'], b['z'], rtol=1e-12, atol=1e-14)      # summation order differs
