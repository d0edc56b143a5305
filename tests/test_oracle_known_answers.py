"""CPU: the oracle against the analytic known answers of SURVEY.md Appendix D (the only pins
available -- the reference ships no tests; parity is otherwise unpinned)."""
import numpy as np
import pytest

from oracle import ao_oracle as O


@pytest.fixture(scope='module')
def env5(flat_screen):
    e = O.OracleAOEnv(obs_dim=5, act_type='zernike', act_dim=6, rew_type='smf_ssim', initial_screen=flat_screen)
    e.reset()
    return e


def test_aperture_pixel_count(env5):
    assert int(env5.aperture.sum()) == 45244
    assert env5.pupil_grid.delta[0] == pytest.approx(0.5 / 240)


def test_unaberrated_obs_5x5(env5):
    obs = env5.last_obs_f64.reshape(5, 5)
    want = np.array([[9.57057142e-4, 5.51183449e-4, 2.91760398e-3, 5.51183449e-4, 9.57057142e-4],
                     [5.51183449e-4, 1.16313893e-2, 1.86736332e-2, 1.16313893e-2, 5.51183449e-4],
                     [2.91760398e-3, 1.86736332e-2, 3.01752344, 1.86736332e-2, 2.91760398e-3]])
    np.testing.assert_allclose(obs[:3], want, rtol=2e-8)
    np.testing.assert_allclose(obs, obs[::-1, ::-1], rtol=1e-10)     # point symmetry


def test_unaberrated_fiber_power_ssim_reward(env5):
    reward, rew_fiber = env5.reward_function()
    assert rew_fiber == pytest.approx(0.7571639, abs=2e-7)
    assert env5.last_ssim == pytest.approx(0.9621218103725746, abs=1e-12)
    assert reward == pytest.approx(0.8 * rew_fiber + 0.2 * env5.last_ssim, abs=1e-15)
    assert env5.wf_wfs_after_foc.total_power == pytest.approx(0.9631288, abs=2e-7)


def test_ssim_of_zeros_and_window_error():
    ref = np.zeros(25)
    ref[12] = 2.8
    assert O.skimage_ssim_1d(np.zeros(25), ref, 2.8) == pytest.approx(0.6315901942144502, abs=1e-13)
    with pytest.raises(ValueError):
        O.skimage_ssim_1d(np.zeros(4), np.zeros(4), 2.8)


def test_lp_modes(env5):
    M, beta = env5.single_mode_fiber.instance(env5.propagator_fiber.output_grid, 1.5e-6)
    assert M.shape == (128 * 128, 3)
    np.testing.assert_allclose(beta, [4171697.52, 4150088.19, 4150088.19], rtol=3e-9)
    gram = M.T @ M * env5.propagator_fiber.output_grid.weight
    np.testing.assert_allclose(gram, np.eye(3), atol=1e-12)
    V = env5.single_mode_fiber.V(1.5e-6)
    assert V == pytest.approx(2.63894, abs=1e-5)
    u01 = O._lp_find_branch_cuts(0, V)[0]
    u11 = O._lp_find_branch_cuts(1, V)[0]
    assert u01[0] == pytest.approx(1.7011139, abs=1e-6) and u11[0] == pytest.approx(2.5564255, abs=1e-6)
    assert O._lp_find_branch_cuts(2, V) is None


def test_flat_wavefront_strehl_100_and_2x2_obs(flat_screen):
    e = O.OracleAOEnv(obs_dim=2, act_dim=8, rew_type='strehl_ratio', atm_fried=0.2, initial_screen=flat_screen)
    o, info = e.reset()
    assert info == {} and o.dtype == np.float16
    np.testing.assert_allclose(e.last_obs_f64, 0.01575277, rtol=3e-7)
    assert int(np.argmax(e.unaberrated_PSF)) == 120 * 240 + 120
    # flat DM (zero surface) evaluated directly: Strehl 100 -> reward 0
    e.wf_wfs_after_foc = e.propagator_fiber(e.wf_wfs_fiber)
    reward, _ = e.reward_function()
    assert e.last_strehl == pytest.approx(100.0, abs=1e-9) and reward == pytest.approx(0.0, abs=1e-9)


def test_action_scaling_makes_surface_rms_constant(flat_screen):
    e = O.OracleAOEnv(act_dim=16, initial_screen=flat_screen)
    e.reset()
    for scale in (1e-3, 1.0, 50.0):
        e.step(scale * np.random.default_rng(0).uniform(-1, 1, 16).astype(np.float32))
        assert np.std(e.deformable_mirror.surface) == pytest.approx(0.1 * 2.2e-6, rel=1e-12)
    with np.errstate(all='ignore'):
        o, r, d, tr, info = e.step(np.zeros(16, dtype=np.float32))
    assert np.isnan(r) and np.all(np.isnan(o.astype(float)))


def test_episode_bookkeeping(flat_screen):
    e = O.OracleAOEnv(act_dim=4, timesteps_per_episode=3, initial_screen=flat_screen)
    a = np.ones(4, dtype=np.float32)
    for ep in range(2):
        e.reset()
        dones = [e.step(a)[2] for _ in range(3)]
        assert dones == [False, False, True]
    assert e.timestep == 6 and e.episode_no == 2 and e.timestep_render == 3   # global time never resets


def test_velocity_coercion():
    assert O.OracleAOEnv(atm_type='semi_dynamic', atm_vel=5, act_dim=4).velocity == 0
    assert O.OracleAOEnv(atm_type='dynamic', atm_vel=0, act_dim=4).velocity == 1


def test_von_karman_covariance_known_answers():
    cov = O.hcipy_phase_covariance_von_karman(0.2, 10.0)
    c0 = cov(np.array(0.0)) / (10.0 / 0.2) ** (5 / 3)
    assert c0 == pytest.approx(0.0863143, rel=2e-4)
    for r, want in ((5e-3, 0.883), (20e-3, 0.813), (100e-3, 0.681)):
        D = 2 * (cov(np.array(0.0)) - cov(np.array(r)))
        assert D / (6.88 * (r / 0.2) ** (5 / 3)) == pytest.approx(want, abs=2e-3)


def test_extrusion_statistics_and_mechanics():
    """AR extrusion on a 32x32 grid: column shift semantics and the structure function of
    extruded screens (acceptance test of SURVEY Appendix D)."""
    g = O.hcipy_make_pupil_grid(32, 0.5 * 32 / 240)
    cn2 = O.hcipy_Cn_squared_from_fried_parameter(0.2, 2.2e-6)
    rng = np.random.default_rng(5)
    lay = O.InfiniteAtmosphericLayer(g, cn2, 10.0, 5.0, rng)
    s0 = lay.achromatic_screen.reshape(32, 32).copy()
    lay._extrude('right')
    s1 = lay.achromatic_screen.reshape(32, 32)
    np.testing.assert_array_equal(s1[:, :-1], s0[:, 1:])         # +x drift: columns move left
    lay._extrude('left')
    np.testing.assert_array_equal(lay.achromatic_screen.reshape(32, 32)[:, 1:], s1[:, :-1])
    # statistics
    r0 = O.hcipy_fried_parameter_from_Cn_squared(cn2, 1.0)
    cov = O.hcipy_phase_covariance_von_karman(r0, 10.0)
    lag = 4
    th = 2 * (cov(np.array(0.0)) - cov(np.array(lag * g.delta[0])))
    acc = []
    for rep in range(40):
        lay.reset()
        for _ in range(64):
            lay._extrude('right')
        s = lay.achromatic_screen.reshape(32, 32)
        acc.append(np.mean((s[:, lag:] - s[:, :-lag]) ** 2))
    assert np.mean(acc) / th == pytest.approx(1.0, abs=0.12)


def test_evolve_until_counts_extrusions():
    g = O.hcipy_make_pupil_grid(16, 0.5 * 16 / 240)
    lay = O.InfiniteAtmosphericLayer(g, 1e-12, 10.0, 20.0, np.random.default_rng(0))
    n = []
    lay._extrude = lambda where: n.append(where)
    for k in range(1, 11):
        lay.t = k * 1e-3
    assert len(n) == 96 and set(n) == {'right'}                 # 10 ms * 20 m/s / (0.5/240)
