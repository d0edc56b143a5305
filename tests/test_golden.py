"""Committed golden trajectories (tests/golden/*.npz, made by tools/make_golden.py from the CPU oracle):
the oracle must reproduce them (CPU), and the CUDA path through the C-ABI must match them (GPU).
Tolerances: oracle re-run 1e-9 (BLAS summation order may differ between hosts); FP64 CUDA path 1e-7
(dynamic cases accumulate extrusion rounding), tensor path 1e-5 (the north star's bound)."""
import glob
import json
import os

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
FILES = sorted(f for f in glob.glob(os.path.join(HERE, 'golden', 'config*.npz')))
NAMES = [os.path.basename(f)[:-4] for f in FILES]


def _load(name):
    z = np.load(os.path.join(HERE, 'golden', name + '.npz'))
    return z, json.loads(str(z['case']))


def _rel(a, b):
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    return float(np.max(np.abs(a - b) / np.maximum(np.abs(b), 1e-300)))


def test_golden_files_present():
    assert len(FILES) >= 4 and os.path.exists(os.path.join(HERE, 'golden', 'ar_tables.npz'))


@pytest.mark.parametrize('name', NAMES)
def test_oracle_reproduces_golden(name):
    import sys
    sys.path.insert(0, os.path.dirname(HERE))
    from tools.make_golden import run_case
    z, case = _load(name)
    out = run_case(case, z['screen'])
    for k in ('reset_obs', 'obs', 'reward', 'power', 'aux', 'actions'):
        assert _rel(out[k], z[k]) < 1e-9, k
    assert np.array_equal(out['done'], z['done']) and np.array_equal(out['noise_counts'], z['noise_counts'])
    assert np.array_equal(out['obs_f16'], z['obs_f16'])
    assert int(out['final_timestep']) == int(z['final_timestep'])
    assert int(out['final_episode_no']) == int(z['final_episode_no'])


@pytest.mark.gpu
@pytest.mark.parametrize('precision', ['f64', 'tensor', 'fused'])
@pytest.mark.parametrize('name', NAMES)
def test_cuda_path_matches_golden(name, precision):
    import sys
    sys.path.insert(0, os.path.dirname(HERE))
    from adaptive_optics_gym_b200 import AOEnv
    from tools.make_golden import load_ar_tables, oracle_env
    z, case = _load(name)
    kw = case['kw']
    tabs = None
    if kw['atm_type'] == 'dynamic':
        # the committed AR matrices (tests/golden/ar_tables.npz); SH calibration from the oracle
        tabs = load_ar_tables()
        if kw.get('SH_operation'):
            ref = oracle_env(case, z['screen'])
            sh = ref.shwfs
            idx = sh.estimation_subapertures
            tabs.update(sh_recon=ref.reconstruction_matrix,
                        sh_offset=np.array((sh.mla_x[idx], sh.mla_y[idx])) + ref.slopes_ref)
    env = AOEnv(**kw, initial_screen=z['screen'], precision=precision, tables=tabs)
    rtol = {'f64': 1e-7, 'tensor': 1e-5, 'fused': 1e-5}[precision]
    T = kw['timesteps_per_episode']
    noise, pos, i = z['noise'], 0, 0
    for ep in range(case['episodes']):
        env.reset()
        assert _rel(env.last_obs_f64, z['reset_obs'][ep]) < rtol
        for t in range(T):
            if case['action'] == 'sh':
                a = env.SH_step(noise='none')[0]
                assert np.max(np.abs(a - z['actions'][i])) < 1e-7 * np.max(np.abs(z['actions'][i]))
            else:
                a = z['actions'][i]
            cnt = int(z['noise_counts'][i]) * 240
            nz = noise[pos:pos + cnt] if cnt else None
            pos += cnt
            o, r, d, tr, info = env.step(a, extrusion_noise=nz)
            assert d == bool(z['done'][i]) and tr is False
            assert _rel(env.last_obs_f64, z['obs'][i]) < rtol, ('obs', i)
            assert _rel(r, z['reward'][i]) < rtol and _rel(info['power'], z['power'][i]) < rtol
            aux = env.last_strehl if kw.get('rew_type', 'strehl_ratio') == 'strehl_ratio' else env.last_ssim
            assert _rel(aux, z['aux'][i]) < rtol
            if precision == 'f64' and kw['atm_type'] != 'dynamic':
                assert np.array_equal(o.view(np.uint16), z['obs_f16'][i]), 'float16 obs bits'
            i += 1
    assert env.timestep == int(z['final_timestep']) and env.episode_no == int(z['final_episode_no'])
    env.close()
