"""Loader for the committed golden step vectors (tests/golden/step_*.npz)."""
import os
import glob
import numpy as np

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden')


def case_names():
    return sorted(os.path.basename(p)[5:-4] for p in glob.glob(os.path.join(GOLDEN_DIR, 'step_*.npz')))


def load_case(name):
    z = np.load(os.path.join(GOLDEN_DIR, 'step_%s.npz' % name))
    nt = int(z['n_traits'])
    traits = []
    for t in range(nt):
        traits.append(dict(loci=z['trait%i_loci' % t], alpha=z['trait%i_alpha' % t],
                           phi=(float(z['trait%i_phi' % t]) if z['trait%i_phi' % t].ndim == 0 else z['trait%i_phi' % t]), gamma=float(z['trait%i_gamma' % t]),
                           lyr_num=int(z['trait%i_lyr' % t]),
                           univ_adv=bool(z['trait%i_univ_adv' % t])))
    arch = dict(land_dim=tuple(int(v) for v in z['land_dim']), rasters=z['rasters'], K=z['K'],
                ww=float(z['ww']), traits=traits, dom=z['dom'], paths=z['paths'],
                move_surf=z['move_surf'] if 'move_surf' in z.files else None,
                disp_surf=z['disp_surf'] if 'disp_surf' in z.files else None)
    if 'res_ratio' in z.files:
        arch['res_ratio'] = tuple(float(v) for v in z['res_ratio'])
    if 'mut_mu_neut' in z.files:
        arch['mutation'] = dict(mu_neut=float(z['mut_mu_neut']), mu_delet=float(z['mut_mu_delet']),
                                trait_mus=[float(v) for v in z['mut_trait_mus']],
                                mutables=[int(v) for v in z['mut_mutables']],
                                nonneut_loci=z['mut_nonneut_loci'], delet_loci=z['mut_delet_loci'],
                                delet_s=z['mut_delet_s'], s_shape=float(z['mut_s_shape']),
                                s_scale=float(z['mut_s_scale']))
        if 'use_tskit' in z.files and int(z['use_tskit']):
            # gen_arch.use_tskit = True: genotype ROWS per non-neutral locus; the evolving tables travel in the
            # mutation dict (oracle/step_oracle.py mutate_tskit)
            for t, tr in enumerate(traits):
                tr['loci_idxs'] = z['trait%i_loci_idxs' % t]
                tr['alpha_distr'] = z['trait%i_alpha_distr' % t]
            arch['mutation'].update(tskit_layout=True, traits=traits, delet_loci_idxs=z['mut_delet_loci_idxs'],
                                    subsetters=z['subsetters'], paths=z['paths'])
    if arch['ww'] == int(arch['ww']):
        arch['ww'] = int(arch['ww'])
    prm = dict(b=float(z['prm_b']), R=float(z['prm_R']), lam=float(z['prm_n_births_distr_lambda']),
               n_births_fixed=bool(z['prm_n_births_fixed']),
               mating_radius=None if float(z['prm_mating_radius']) < 0 else float(z['prm_mating_radius']),
               d_min=float(z['prm_d_min']),
               d_max=float(z['prm_d_max']), sex=bool(z['prm_sex']),
               sex_ratio_p=float(z['prm_sex_ratio_p']),
               max_age=None if z['prm_max_age'] < 0 else int(z['prm_max_age']),
               direction_mu=float(z['prm_direction_distr_mu']),
               direction_kappa=float(z['prm_direction_distr_kappa']),
               burn=bool(z['prm_burn']) if 'prm_burn' in z.files else False,
               choose_nearest=bool(z['prm_choose_nearest']) if 'prm_choose_nearest' in z.files else False,
               inverse_dist=bool(z['prm_inverse_dist']) if 'prm_inverse_dist' in z.files else False)
    if prm['lam'] == int(prm['lam']):
        prm['lam'] = int(prm['lam'])
    state = dict(x=z['in_x'], y=z['in_y'], age=z['in_age'], sex=z['in_sex'], idx=z['in_idx'],
                 g=z['in_g'], z=z['in_z'], max_ind_idx=int(z['in_max_ind_idx']))
    draws = {k[5:]: z[k] for k in z.files if k.startswith('draw_')}
    return z, arch, prm, state, draws
