"""Species change events (a17): the schedule built by geonomics_b200.api._SpeciesChanger against
the UNMODIFIED reference's ops/change.py::_SpeciesChanger (change.py:155-267, 612-742), run on
stand-in species objects -- same K trajectory and same parameter values at every time step, for
every kind of demographic event and for life-history events."""
import os
import sys
import types

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
from oracle import ref_shims  # noqa: E402

pytestmark = pytest.mark.skipif(ref_shims.reference_root() is None, reason='reference package not present')

CASES = {
    'monotonic': {'dem': {0: dict(kind='monotonic', start_t=3, end_t=9, rate=1.05)}},
    'stochastic': {'dem': {0: dict(kind='stochastic', start_t=2, end_t=20, interval=3, distr='uniform',
                                   size_range=(0.5, 1.5))}},
    'stochastic_normal': {'dem': {0: dict(kind='stochastic', start_t=1, end_t=12, interval=None, distr='normal',
                                          size_range=(0.8, 1.4))}},
    'cyclical': {'dem': {0: dict(kind='cyclical', start_t=4, end_t=28, n_cycles=3, size_range=(0.5, 2.0))}},
    'cyclical_minmax': {'dem': {0: dict(kind='cyclical', start_t=0, end_t=20, n_cycles=2, min_size=0.25,
                                        max_size=1.5, increase_first=False)}},
    'custom': {'dem': {0: dict(kind='custom', timesteps=[5, 9, 15], sizes=[2, 5, 0.5])}},
    'two_events_and_life_history': {
        'dem': {0: dict(kind='monotonic', start_t=2, end_t=4, rate=0.9),
                1: dict(kind='custom', timesteps=[6, 8], sizes=[3, 1])},
        'life_hist': {'b': dict(timesteps=[3, 7], vals=[0.4, 0.1]), 'd_max': dict(timesteps=[7], vals=[0.8])}},
}


class _Spp:
    def __init__(self):
        self.K = np.linspace(1, 4, 12).reshape(3, 4)
        self.t = -1
        self.b = 0.2
        self.d_max = 1.0
        self._move_surf = None
        self._disp_surf = None

    def _override_K(self, K):
        self.K = K

    def _set_parameter(self, parameter, val):
        setattr(self, parameter, val)


def _trajectory(make, T=32):
    spp = _Spp()
    np.random.seed(99)
    ch = make(spp)
    out = []
    for t in range(T):
        spp.t = t
        ch(t, spp)
        out.append((spp.K.copy(), spp.b, spp.d_max))
    return out


@pytest.mark.parametrize('case', sorted(CASES))
def test_schedule_matches_reference(case):
    gnx = ref_shims.install()
    from geonomics.ops import change as ref_change
    from geonomics.sim.params import ParametersDict
    from geonomics_b200 import api
    params = CASES[case]

    def make_ref(spp):
        land = types.SimpleNamespace(_changer=None)
        c = ref_change._SpeciesChanger(spp, ParametersDict(params), land=land)
        return lambda t, s: c._make_change(t=t, additional_args={'spp': s})

    def make_ours(spp):
        c = api._SpeciesChanger(spp, params, types.SimpleNamespace(_changer=None))
        return lambda t, s: c._make_change(t, s)

    ref = _trajectory(make_ref)
    ours = _trajectory(make_ours)
    for t, ((Kr, br, dr), (Ko, bo, do)) in enumerate(zip(ref, ours)):
        assert np.array_equal(Kr, Ko), 'K differs at t=%d' % t
        assert br == bo and dr == do, 'life-history parameter differs at t=%d' % t
    assert not np.array_equal(ref[0][0], ref[-1][0]) or case in ('stochastic', 'stochastic_normal', 'cyclical',
                                                                 'cyclical_minmax', 'two_events_and_life_history')
