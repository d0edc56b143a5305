"""CPU: the product's set-up tables (adaptive_optics_gym_b200/tables.py) against the oracle's
independently written hcipy restatement."""
import numpy as np
import pytest

from adaptive_optics_gym_b200.tables import (AOConfig, ar_extrusion_tables, build_tables, disk_harmonic_orders,
                                             synthesize_screens, zernike_noll)
from oracle import ao_oracle as O


@pytest.mark.parametrize('act_type,K,n', [('zernike', 6, 5), ('num_actuators', 64, 2), ('zernike', 21, 3)])
def test_tables_match_oracle(act_type, K, n):
    tb = build_tables(AOConfig(act_type=act_type, num_modes=K, obs_dim=n))
    env = O.OracleAOEnv(act_type=act_type, act_dim=K, obs_dim=n, initial_screen=np.zeros(57600))
    np.testing.assert_allclose(tb['dm_modes'], env.deformable_mirror.M.T, atol=5e-15)
    np.testing.assert_array_equal(tb['aperture'], env.aperture)
    a = np.random.default_rng(0).normal(size=K)
    assert np.sqrt(a @ tb['dm_gram'] @ a) == pytest.approx(np.std(env.deformable_mirror.M @ a), rel=1e-12)
    for prop, k1, k2 in ((env.propagator_fiber, 'mft_fib_1', 'mft_fib_2'),
                         (env.propagator_fiber_subsample, 'mft_obs_1', 'mft_obs_2')):
        M1, M2, norm = prop.matrices(env.pupil_grid, 1.5e-6)
        np.testing.assert_allclose(tb[k1], M1, atol=1e-13 * np.abs(M1).max())
        np.testing.assert_allclose(tb[k2], M2, atol=1e-13)
        assert tb['mft_fib_norm'] == pytest.approx(norm, rel=1e-15)
    fg = env.propagator_fiber.output_grid
    L, beta = env.single_mode_fiber.instance(fg, 1.5e-6)
    np.testing.assert_allclose(tb['lp_modes_w'], L.T * fg.weight, atol=1e-13 * np.abs(L).max() * fg.weight)
    np.testing.assert_allclose(tb['lp_beta'], beta, rtol=1e-13)
    np.testing.assert_allclose(tb['lp_phase'], np.exp(1j * beta * 10), atol=1e-6)
    assert tb['amp_fiber'] == pytest.approx(env.wf_wfs_fiber.electric_field.real.max(), rel=1e-14)
    assert tb['amp_flux'] == pytest.approx(env.wf_sci.electric_field.real.max(), rel=1e-14)
    assert tb['sci_focal_index'] == int(np.argmax(env.unaberrated_PSF))
    assert tb['obs_weight'] == pytest.approx(env.propagator_fiber_subsample.output_grid.weight, rel=1e-15)
    # Strehl by single-pixel sum == the reference's full-plane route
    scr = O.von_karman_screen(env.pupil_grid, tb['cn2'], 10.0, np.random.default_rng(1))
    env.layer.achromatic_screen = scr
    env.reset()
    env.rew_type = 'strehl_ratio'
    env.reward_function()
    s = np.sum(env.aperture * np.exp(1j * scr / 2.2e-6))
    assert tb['strehl_scale'] * abs(s) ** 2 == pytest.approx(env.last_strehl, rel=1e-11)


def test_noll_and_disk_harmonic_ordering():
    assert [zernike_noll(j) for j in range(1, 8)] == [O.hcipy_noll_to_zernike(j) for j in range(1, 8)]
    assert zernike_noll(4) == (2, 0) and zernike_noll(2) == (1, 1) and zernike_noll(3) == (1, -1)
    assert disk_harmonic_orders(64) == O.hcipy_disk_harmonic_orders_sorted(64)
    assert disk_harmonic_orders(5) == [(1, 0), (1, -1), (1, 1), (1, -2), (1, 2)]


def test_ar_tables_match_oracle_given_same_stencil():
    g = O.hcipy_make_pupil_grid(48, 0.5 * 48 / 240)
    lay = O.InfiniteAtmosphericLayer(g, 1e-13, 10.0, 5.0, np.random.default_rng(2))
    st = np.flatnonzero(lay.stencil_left).reshape(48, 3)
    idx, A, B = ar_extrusion_tables(48, g.delta[0], 10.0, None, extra_columns=st[:, 2] % 48)
    np.testing.assert_array_equal(idx, np.flatnonzero(lay.stencil_left))
    np.testing.assert_allclose(A, lay.A_horizontal, atol=1e-6)
    np.testing.assert_allclose(B @ B.T, lay.B_horizontal @ lay.B_horizontal.T, atol=1e-4 * np.abs(B @ B.T).max())   # difference of near-equal covariances


def test_host_screen_synthesis_statistics():
    tb = build_tables(AOConfig(num_modes=4))
    scr = synthesize_screens(tb, 96, tb['cn2'], np.random.default_rng(0)).reshape(-1, 240, 240)
    r0 = O.hcipy_fried_parameter_from_Cn_squared(tb['cn2'], 1.0)
    cov = O.hcipy_phase_covariance_von_karman(r0, 10.0)
    for lag in (4, 16, 64):
        th = 2 * (cov(np.array(0.0)) - cov(np.array(lag * 0.5 / 240)))
        d = np.mean((scr[:, :, lag:] - scr[:, :, :-lag]) ** 2)
        assert d / th == pytest.approx(1.0, abs=0.15)


def test_shack_hartmann_tables_match_oracle():
    """build_sh_tables (separable Fresnel operator C E C^T, CSR lenslet pixel lists, calibration) against the
    oracle's restatement of shack_hartmann_init (2-D zero-padded FFT route, ndimage label sums)."""
    from adaptive_optics_gym_b200.tables import build_sh_tables
    cfg = AOConfig(act_type='zernike', num_modes=6, obs_dim=2, atm_type='dynamic', velocity=20.0)
    tb = build_tables(cfg)
    sh = build_sh_tables(cfg, tb)
    env = O.OracleAOEnv(atm_type='dynamic', atm_vel=20, act_type='zernike', act_dim=6, obs_dim=2, SH_operation=True,
                        initial_screen=np.zeros(57600))
    o = env.shwfs
    np.testing.assert_array_equal(sh['sh_selected'], o.estimation_subapertures)
    assert sh['sh_num_sub'] == o.estimation_subapertures.size
    np.testing.assert_allclose(sh['sh_mla_phase'], o.mla_opd * 2 * np.pi / 1.5e-6, rtol=1e-12, atol=1e-12)
    np.testing.assert_allclose(sh['sh_slopes_ref'], env.slopes_ref, atol=1e-12 * np.abs(env.slopes_ref).max() + 1e-18)
    np.testing.assert_allclose(sh['sh_response'], env.response_matrix, atol=1e-8 * np.abs(env.response_matrix).max())
    # (piston's response is rounding noise; its regularised inverse row differs at the 1e-6 level)
    np.testing.assert_allclose(sh['sh_recon'], env.reconstruction_matrix,
                               atol=1e-5 * np.abs(env.reconstruction_matrix).max())
    np.testing.assert_allclose(sh['sh_act0'], env.deformable_mirror_shack.actuators, rtol=1e-15)
    # the separable operator reproduces the camera image of an aberrated field
    scr = O.von_karman_screen(env.pupil_grid, tb['cn2'], 10.0, np.random.default_rng(3))
    env.layer.achromatic_screen = scr
    env.SH_step(poisson=False)
    E = sh['sh_amplitude'] * tb['aperture'] * np.exp(
        1j * (scr / 1.5e-6 + 2 * (2 * np.pi / 1.5e-6) * (sh['sh_act0'] @ tb['dm_modes']) + sh['sh_mla_phase']))
    C = sh['sh_fresnel']
    img = (np.abs(C @ E.reshape(240, 240) @ C.T) ** 2).ravel() * sh['sh_weight_dt']
    np.testing.assert_allclose(img, env.last_sh_image, atol=1e-10 * env.last_sh_image.max())
    # CSR lists cover exactly the pixels of the selected lenslets
    off, pix = sh['sh_pix_offsets'], sh['sh_pix_index']
    for m in (0, sh['sh_num_sub'] // 2, sh['sh_num_sub'] - 1):
        np.testing.assert_array_equal(np.sort(pix[off[m]:off[m + 1]]),
                                      np.flatnonzero(o.mla_index == o.estimation_subapertures[m]))


def test_tables_have_the_symmetries_the_fast_kernels_exploit():
    """The library picks its fast forms from table symmetries it verifies at upload (and falls back otherwise):
    conjugate-paired obs rows and real-times-phasor back-projected fibre modes (fused kernel MODE 2), conjugate-paired
    rows of the screen-synthesis matrices (real-arithmetic synthesis), centrosymmetric Fresnel operator (parity-folded
    SH step).  The reference geometry (hcipy's symmetric grids) must have all of them."""
    from adaptive_optics_gym_b200.tables import build_sh_tables
    for n in (2, 5):
        cfg = AOConfig(act_type='zernike', num_modes=6, obs_dim=n)
        tb = build_tables(cfg)
        m1o = tb['mft_obs_1']                                    # [n][Np]
        assert np.abs(m1o[::-1] - np.conj(m1o)).max() <= 1e-12 * np.abs(m1o).max()
    M1, M2, lpw = tb['mft_fib_1'], tb['mft_fib_2'], tb['lp_modes_w']
    for j in range(lpw.shape[0]):
        G = M1.T @ lpw[j].reshape(M1.shape[0], -1) @ M2.T        # fibre mode propagated back to the pupil
        k = np.argmax(np.abs(G))
        ph = G.flat[k] / np.abs(G.flat[k])
        assert np.abs((G / ph).imag).max() <= 1e-10 * np.abs(G).max()
    for key in ('scr_W1', 'scr_W2'):
        W = tb[key]
        assert W.shape[0] % 2 == 0 and W.shape[1] % 2 == 0
        assert np.abs(W[::-1] - np.conj(W)).max() <= 1e-12 * np.abs(W).max()
    sh = build_sh_tables(AOConfig(act_type='num_actuators', num_modes=64, obs_dim=2), build_tables(
        AOConfig(act_type='num_actuators', num_modes=64, obs_dim=2)))
    C = sh['sh_fresnel']
    assert np.abs(C[::-1, ::-1] - C).max() <= 1e-13 * np.abs(C).max()
