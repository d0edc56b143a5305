"""GPU: the tensor-core (tcgen05, split-fp16) path stage by stage against FP64 NumPy built from
the oracle's tables -- localises a failure to the field kernel, MFT stage 1 or stage 2."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def test_tensor_path_stages():
    from adaptive_optics_gym_b200 import AOEnv
    from adaptive_optics_gym_b200._lib import AogError
    from oracle.ao_oracle import OracleAOEnv
    from tests.test_parity_gpu import _screen
    kw = dict(atm_fried=0.2, act_dim=64, obs_dim=2, rew_type='strehl_ratio', timesteps_per_episode=5)
    scr = _screen(0, 0.2).astype(np.float32)
    try:
        env = AOEnv(precision='tensor', **kw, initial_screen=scr)
    except AogError as e:
        if 'not built' in str(e):
            pytest.skip('tensor path not built')
        raise
    ref = OracleAOEnv(**kw, initial_screen=scr)
    env.reset(), ref.reset()
    a = np.random.default_rng(1).uniform(-1, 1, 64).astype(np.float32)
    o, r, d, _, info = env.step(a)
    ro, rr, rd, _, rinfo = ref.step(a)
    # DM actuators after normalisation
    np.testing.assert_allclose(env._h.get_field('actuators'), ref.deformable_mirror.actuators, rtol=1e-12)
    # pupil field (unit modulus x aperture)
    phase = ref.layer.phase_for(1.5e-6) + 2 * ref.deformable_mirror.surface * (2 * np.pi / 1.5e-6)
    E_ref = (ref.aperture * np.exp(1j * phase)).reshape(240, 240)
    E = env._h.get_field('tc_pupil').reshape(240, 240)
    assert np.abs(E - E_ref).max() < 5e-6, np.abs(E - E_ref).max()
    # stage 1: T = M1~ . E with unit-modulus twiddles
    M1, M2, norm = ref.propagator_fiber.matrices(ref.pupil_grid, 1.5e-6)
    w = ref.pupil_grid.weight
    T_ref = (M1 / w) @ E_ref
    T = env._h.get_field('tc_stage1').reshape(128, 240)
    err = np.abs(T - T_ref).max()
    assert err < 2e-3, err                      # |T| up to 240; FP32-class accumulation
    assert err / np.abs(T_ref).max() < 2e-5
    # stage 2 + projection: fibre power
    assert abs(info['power'] - rinfo['power']) <= 1e-5 * rinfo['power'], (info['power'], rinfo['power'])
    np.testing.assert_allclose(env.last_obs_f64, ref.last_obs_f64, rtol=1e-5)
    assert abs(r - rr) <= 1e-5 * abs(rr)
    env.close()
