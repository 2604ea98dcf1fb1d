"""Diagnostic: KS p-values and mean N per time step of the GPU path vs the reference's 100
replicates, for several seed bases (is a failure of tests/test_statistical_parity.py chance?)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'tests', 'golden'))
import numpy as np
from scipy.stats import ks_2samp
import make_stat_golden as msg
from geonomics_b200 import api
ref = np.load(os.path.join(ROOT, 'tests', 'golden', 'stat_reference.npz'))


def fixed_burn(self):
    ok = all(len(s.Nt) >= self.burn_T for s in self.comm.values())
    for s in self.comm.values():
        s.burned = ok
    self.comm.burned = ok
api.Model._check_comm_burned = fixed_burn
bases = [int(a) for a in sys.argv[1:]] or [5000, 6000, 7000]
allN = []
for base in bases:
    Nt = []
    for rep in range(100):
        p = api.make_params_dict(msg.stat_params(), 'stat')
        p['model']['seed'] = {'num': base + rep}
        mod = api.make_model(p)
        mod.walk(10000, 'burn')
        spp = mod.comm[0]
        mod.walk(msg.T, 'main')
        Nt.append(spp.Nt[-msg.T:])
        burnN = spp.Nt[:-msg.T]
    Nt = np.array(Nt)
    allN.append(Nt)
    print('base', base, 'burn len', len(burnN))
    print('  t     refmean  gpumean   p')
    for t in range(0, msg.T, 3):
        print('  %2d  %8.1f %8.1f  %.3g' % (t, ref['Nt'][:, t].mean(), Nt[:, t].mean(),
                                            ks_2samp(ref['Nt'][:, t], Nt[:, t]).pvalue))
A = np.concatenate(allN)
print('pooled %d reps: mean over t ref %.2f gpu %.2f' % (len(A), ref['Nt'].mean(), A.mean()))
for t in range(0, msg.T, 3):
    print('  %2d  %8.1f %8.1f  %.3g' % (t, ref['Nt'][:, t].mean(), A[:, t].mean(), ks_2samp(ref['Nt'][:, t], A[:, t]).pvalue))
