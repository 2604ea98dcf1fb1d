# A Geonomics parameters file (same structure as gnx.make_parameters_file() writes,
# geonomics/sim/params.py templates), trimmed of comments.  Used by the host-API tests.
import numpy as np

_dim = (40, 40)
_xg = np.tile(np.linspace(0, 1, _dim[0]), (_dim[1], 1))
_yg = np.tile(np.linspace(0, 1, _dim[1])[:, None], (1, _dim[0]))

params = {
    'landscape': {
        'main': {'dim': _dim, 'res': (1, 1), 'ulc': (0, 0), 'prj': None},
        'layers': {
            'lyr_0': {'init': {'defined': {'rast': np.ones((_dim[1], _dim[0])), 'pts': None, 'vals': None,
                                           'interp_method': None}}},
            'lyr_1': {'init': {'defined': {'rast': _xg, 'pts': None, 'vals': None, 'interp_method': None}},
                      'change': {0: {'change_rast': _xg[:, ::-1].copy(), 'start_t': 4, 'end_t': 8,
                                     'n_steps': 3}}},
            'lyr_2': {'init': {'random': {'n_pts': 200, 'interp_method': 'linear'}}},
        },
    },
    'comm': {
        'species': {
            'spp_0': {
                'init': {'N': 1200, 'K_layer': 'lyr_0', 'K_factor': 0.75},
                'mating': {'repro_age': 0, 'sex': False, 'sex_ratio': 1 / 1, 'R': 0.5, 'b': 0.2,
                           'n_births_distr_lambda': 1, 'n_births_fixed': True, 'mating_radius': 2,
                           'choose_nearest_mate': False, 'inverse_dist_mating': False},
                'mortality': {'max_age': None, 'd_min': 0, 'd_max': 1, 'density_grid_window_width': None},
                'movement': {'move': True, 'direction_distr_mu': 0, 'direction_distr_kappa': 0,
                             'movement_distance_distr_param1': 1.0, 'movement_distance_distr_param2': 1.0,
                             'movement_distance_distr': 'wald',
                             'dispersal_distance_distr_param1': 1.0, 'dispersal_distance_distr_param2': 1.0,
                             'dispersal_distance_distr': 'wald'},
                'gen_arch': {
                    'gen_arch_file': None, 'L': 60, 'start_p_fixed': 0.5, 'start_neut_zero': False,
                    'mu_neut': 0, 'mu_delet': 0, 'delet_alpha_distr_shape': 0.2, 'delet_alpha_distr_scale': 0.2,
                    'r_distr_alpha': 0.5, 'r_distr_beta': None, 'dom': False, 'pleiotropy': False,
                    'recomb_rate_custom_fn': None, 'n_recomb_paths_mem': int(1e4), 'n_recomb_paths_tot': int(1e5),
                    'n_recomb_sims': 2000, 'allow_ad_hoc_recomb': False, 'jitter_breakpoints': False,
                    'mut_log': False, 'use_tskit': False, 'tskit_simp_interval': 100,
                    'traits': {
                        'trait_0': {'layer': 'lyr_1', 'phi': 0.1, 'n_loci': 8, 'mu': 0, 'alpha_distr_mu': 0,
                                    'alpha_distr_sigma': 0.1, 'max_alpha_mag': 0.25, 'gamma': 1,
                                    'univ_adv': False},
                    },
                },
            },
        },
    },
    'model': {
        'T': 12, 'burn_T': 10, 'seed': {'num': 7},
        'its': {'n_its': 1, 'rand_landscape': False, 'rand_comm': False, 'rand_genarch': True,
                'repeat_burn': False},
    },
}
