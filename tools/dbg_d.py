import sys, os
sys.path.insert(0, os.getcwd()); sys.path.insert(0, os.path.join(os.getcwd(), 'tests'))
import numpy as np
from golden_io import load_case
from parity_util import run_device_step
from oracle import step_oracle as so
z, arch, prm, state, draws = load_case('sexed')
new_o, im_o = so.step(state, arch, prm, draws)
out = run_device_step(arch, prm, state, draws, staged=True)
d, do = out['d_rast'], im_o['d_rast']
bad = np.argwhere(~np.isclose(d, do, rtol=1e-6, atol=1e-9))
print('n bad', len(bad))
for i, j in bad[:12]:
    print((i, j), 'd', d[i, j], do[i, j], 'N', out['N_rast'][i, j], im_o['N_rast'][i, j], 'np', out['n_pairs_rast'][i, j], im_o['n_pairs_rast'][i, j], 'K', arch['K'][i, j])
