"""CPU oracle for the ``AO-v0`` env-step path -- TEST INFRASTRUCTURE ONLY.

NumPy FP64 / complex128 restatement of ``gym_AO/envs/AO_env.py`` (reference,
cited as ``AO_env.py:LINE``) and of the hcipy==0.5.1 / scikit-image==0.22.0
calls it makes.  PARITY UNPINNED (see oracle/__init__.py): hcipy is an
un-vendored dependency (``requirements.txt:1``) that is absent from this image,
so every ``hcipy_*`` function below restates hcipy's published algorithm from
its documented behaviour; the reference has no tests or golden vectors.

The oracle executes the reference's op sequence UNABRIDGED (two DM surface
GEMVs per step, separate exp passes per optical element, three full matrix
Fourier transforms including the 240x240 science plane of which one pixel is
read, the LP-mode projection and back-expansion, skimage's SSIM) -- no
algebraic shortcuts -- so that it can also serve as the timed CPU baseline.

Differences from the reference, all to make parity testable:
  * phase screens, extrusion noise and Poisson noise can be injected;
  * the global NumPy RNG is replaced by an explicit ``numpy.random.Generator``.
"""
from __future__ import annotations

import math
import numpy as np
from scipy import special


# --------------------------------------------------------------------------
# hcipy.field: grids  (Field = flat array, x fastest: flat = iy*N + ix)
# --------------------------------------------------------------------------
class CartesianGrid:
    """Regular, separable cartesian grid (hcipy CartesianGrid(RegularCoords))."""

    def __init__(self, xs, ys, weight):
        self.xs = np.asarray(xs, dtype=np.float64)      # separated coords, x
        self.ys = np.asarray(ys, dtype=np.float64)      # separated coords, y
        self.weight = float(weight)                     # scalar weights (dx*dy)
        self.dims = (self.xs.size, self.ys.size)        # (nx, ny)
        self.shape = (self.ys.size, self.xs.size)       # numpy shape [y, x]
        self.size = self.xs.size * self.ys.size

    @property
    def x(self):
        return np.tile(self.xs, self.ys.size)

    @property
    def y(self):
        return np.repeat(self.ys, self.xs.size)

    @property
    def delta(self):
        return np.array([self.xs[1] - self.xs[0], self.ys[1] - self.ys[0]])

    def scaled(self, s):
        return CartesianGrid(self.xs * s, self.ys * s, self.weight * s * s)


def hcipy_make_pupil_grid(dims, diameter):
    """hcipy.make_pupil_grid (AO_env.py:300,378,384-385): delta = D/N, samples at
    (i + 0.5) * delta - D/2 (symmetric, no sample at 0), weights delta^2."""
    delta = diameter / dims
    zero = -diameter / 2 + delta / 2
    c = zero + delta * np.arange(dims)
    return CartesianGrid(c, c, delta * delta)


def hcipy_make_focal_grid(q, num_airy, spatial_resolution):
    """hcipy.make_focal_grid (AO_env.py:314): delta = res/q, dims = 2*num_airy*q,
    zero = delta * (-dims/2 + (dims mod 2)/2) -> sample at 0 for even dims."""
    delta = spatial_resolution / q
    dims = int(2 * num_airy * q)
    zero = delta * (-dims / 2 + (dims % 2) * 0.5)
    c = zero + delta * np.arange(dims)
    return CartesianGrid(c, c, delta * delta)


def hcipy_make_circular_aperture(diameter):
    """hcipy.make_circular_aperture (AO_env.py:301): binary disc, no supersampling."""
    def func(grid):
        return ((grid.x ** 2 + grid.y ** 2) <= (diameter / 2) ** 2).astype(np.float64)
    return func


# --------------------------------------------------------------------------
# hcipy.optics.Wavefront
# --------------------------------------------------------------------------
class Wavefront:
    """hcipy.Wavefront (scalar field).  power = |E|^2 * grid.weights."""

    def __init__(self, electric_field, grid, wavelength):
        self.electric_field = np.asarray(electric_field).astype(np.complex128)
        self.grid = grid
        self.wavelength = float(wavelength)

    def copy(self):
        return Wavefront(self.electric_field.copy(), self.grid, self.wavelength)

    @property
    def wavenumber(self):
        return 2 * np.pi / self.wavelength

    @property
    def intensity(self):
        return np.abs(self.electric_field) ** 2

    @property
    def power(self):
        return self.intensity * self.grid.weight

    @property
    def total_power(self):
        return float(np.sum(self.power))

    @total_power.setter
    def total_power(self, p):
        self.electric_field *= np.sqrt(p / self.total_power)


# --------------------------------------------------------------------------
# hcipy.propagation.FraunhoferPropagator -> MatrixFourierTransform
# --------------------------------------------------------------------------
class FraunhoferPropagator:
    """hcipy.FraunhoferPropagator (AO_env.py:316,390-391).  Grid-agnostic: the
    matrix Fourier transform is built for the INCOMING wavefront's grid and
    wavelength (cached); uv = output_grid.scaled(2 pi / (f lambda));
    E_out = MFT(E_in) / (i f lambda).

    MFT: F[v,u] = sum_{y,x} E[y,x] exp(-i (u x + v y)) w  computed as M1 @ E @ M2
    with M1 = exp(-i v (x) y) * w  [Nv x Ny],  M2 = exp(-i x (x) u)  [Nx x Nu].
    """

    def __init__(self, input_grid, output_grid, focal_length=1.0):
        self.input_grid = input_grid      # declared, unused (agnostic element)
        self.output_grid = output_grid
        self.focal_length = float(focal_length)
        self._cache = {}

    def matrices(self, in_grid, wavelength):
        key = (id(in_grid), float(wavelength))
        if key not in self._cache:
            uv = self.output_grid.scaled(2 * np.pi / (self.focal_length * wavelength))
            M1 = np.exp(-1j * np.outer(uv.ys, in_grid.ys)) * in_grid.weight
            M2 = np.exp(-1j * np.outer(in_grid.xs, uv.xs))
            norm = 1.0 / (1j * self.focal_length * wavelength)
            self._cache[key] = (M1, M2, norm)
        return self._cache[key]

    def forward(self, wf):
        M1, M2, norm = self.matrices(wf.grid, wf.wavelength)
        f = wf.electric_field.reshape(wf.grid.shape)
        res = (M1 @ f @ M2).ravel() * norm
        return Wavefront(res, self.output_grid, wf.wavelength)

    __call__ = forward


# --------------------------------------------------------------------------
# hcipy.mode_basis: Zernike (Noll) and disk harmonics
# --------------------------------------------------------------------------
def hcipy_noll_to_zernike(i):
    n = int(np.sqrt(2 * i - 1) + 0.5) - 1
    if n % 2:
        m = 2 * int((2 * (i + 1) - n * (n + 1)) // 4) - 1
    else:
        m = 2 * int((2 * i + 1 - n * (n + 1)) // 4)
    return n, m * (-1) ** (i % 2)


def _zernike_radial(n, m, r):
    m = abs(m)
    R = np.zeros_like(r)
    for k in range((n - m) // 2 + 1):
        c = ((-1) ** k * math.factorial(n - k)
             / (math.factorial(k) * math.factorial((n + m) // 2 - k) * math.factorial((n - m) // 2 - k)))
        R += c * r ** (n - 2 * k)
    return R


def hcipy_zernike(n, m, D, grid):
    """hcipy.zernike: sqrt(n+1) * R_n^|m|(2r/D) * {sqrt2 cos m th | sqrt2 sin |m| th | 1},
    zero outside r > D/2 (radial_cutoff=True)."""
    r = 2 * np.hypot(grid.x, grid.y) / D
    theta = np.arctan2(grid.y, grid.x)
    if m < 0:
        az = np.sqrt(2) * np.sin(-m * theta)
    elif m == 0:
        az = np.ones_like(theta)
    else:
        az = np.sqrt(2) * np.cos(m * theta)
    z = np.sqrt(n + 1) * az * _zernike_radial(n, m, np.minimum(r, 1.0))
    return z * (r <= 1)


def hcipy_make_zernike_basis(num_modes, D, grid, starting_mode=1):
    """hcipy.make_zernike_basis (AO_env.py:346): Noll 1..K, piston first."""
    return [hcipy_zernike(*hcipy_noll_to_zernike(i), D, grid)
            for i in range(starting_mode, starting_mode + num_modes)]


def _disk_harmonic_energy(n, m, bc):
    m = abs(m)
    if bc == 'dirichlet':
        lam = special.jn_zeros(m, n)[-1]
    else:
        lam = special.jnp_zeros(m, n)[-1]
    return lam ** 2


def hcipy_disk_harmonic_orders_sorted(num_modes, bc='neumann'):
    """hcipy.get_disk_harmonic_orders_sorted: greedy energy-sorted frontier
    seeded at (n=1, m=0); emit (n,-m) then (n,m) for m != 0."""
    orders = [(1, 0)]
    energies = [_disk_harmonic_energy(1, 0, bc)]
    results = []
    while len(results) < num_modes:
        k = int(np.argmin(energies))
        order = orders[k]
        if order[1] != 0:
            results.append((order[0], -order[1]))
        results.append(order)
        del orders[k]
        del energies[k]
        for new_order in [(order[0], order[1] + 1), (order[0] + 1, order[1])]:
            if new_order not in orders:
                orders.append(new_order)
                energies.append(_disk_harmonic_energy(new_order[0], new_order[1], bc))
    return results[:num_modes]


def hcipy_disk_harmonic(n, m, D, bc, grid):
    r = 2 * np.hypot(grid.x, grid.y) / D
    theta = np.arctan2(grid.y, grid.x)
    m_negative = m < 0
    m = abs(m)
    lam = special.jn_zeros(m, n)[-1] if bc == 'dirichlet' else special.jnp_zeros(m, n)[-1]
    if m_negative:
        z = special.jv(m, lam * r) * np.sin(m * theta)
    else:
        z = special.jv(m, lam * r) * np.cos(m * theta)
    mask = hcipy_make_circular_aperture(D)(grid) > 0.5
    norm = np.sqrt(np.sum(z[mask] ** 2 * grid.weight))
    return z * mask / norm


def hcipy_make_disk_harmonic_basis(grid, num_modes, D, bc='neumann'):
    """hcipy.make_disk_harmonic_basis (AO_env.py:352)."""
    return [hcipy_disk_harmonic(n, m, D, bc, grid)
            for (n, m) in hcipy_disk_harmonic_orders_sorted(num_modes, bc)]


class DeformableMirror:
    """hcipy.DeformableMirror (AO_env.py:348,354,431): surface = M a;
    forward: E * exp(2i * surface * k)."""

    def __init__(self, modes):
        self.M = np.stack(modes, axis=1)       # [P, K] transformation matrix
        self.actuators = np.zeros(self.M.shape[1])

    def flatten(self):
        self.actuators = np.zeros(self.M.shape[1])

    @property
    def surface(self):
        return self.M @ self.actuators

    def forward(self, wf):
        wf2 = wf.copy()
        wf2.electric_field *= np.exp(2j * self.surface * wf.wavenumber)
        return wf2

    __call__ = forward


# --------------------------------------------------------------------------
# hcipy.atmosphere
# --------------------------------------------------------------------------
def hcipy_Cn_squared_from_fried_parameter(r0, wavelength):
    """AO_env.py:367."""
    k = 2 * np.pi / wavelength
    return r0 ** (-5.0 / 3) / (0.423 * k ** 2)


def hcipy_fried_parameter_from_Cn_squared(Cn_squared, wavelength):
    k = 2 * np.pi / wavelength
    return (0.423 * Cn_squared * k ** 2) ** (-3.0 / 5)


def hcipy_phase_covariance_von_karman(r0, L0):
    def func(r):
        r = r + 1e-10
        a = (L0 / r0) ** (5 / 3)
        b = special.gamma(11 / 6) / (2 ** (5 / 6) * np.pi ** (8 / 3))
        c = (24 / 5 * special.gamma(6 / 5)) ** (5 / 6)
        d = (2 * np.pi * r / L0) ** (5 / 6)
        e = special.kv(5 / 6, 2 * np.pi * r / L0)
        return a * b * c * d * e
    return func


def hcipy_inverse_tikhonov(M, rcond):
    U, S, Vt = np.linalg.svd(M, full_matrices=False)
    S_inv = S / (S ** 2 + (rcond * S.max()) ** 2)
    return (Vt.T * S_inv) @ U.T


def von_karman_screen(grid, Cn_squared, L0, rng, oversampling=16):
    """Initial achromatic screen (hcipy FiniteAtmosphericLayer(..., oversampling=16)
    .phase_for(1), used by InfiniteAtmosphericLayer._make_initial_phase_screen).

    Restated as a two-scale spectral synthesis with the von-Karman PSD
    Phi(f) = 0.0229 r0^(-5/3) (f^2 + 1/L0^2)^(-11/6) (f in cycles/m, r0 at
    lambda = 1 m): the FFT frequency grid everywhere except its central 3x3
    bins, which are replaced by an ``oversampling``-times finer grid evaluated
    as a matrix Fourier transform.  Only statistically equivalent to hcipy's
    screen (the RNG stream cannot be matched); parity tests inject screens.
    """
    N = grid.dims[0]
    delta = grid.delta[0]
    r0 = hcipy_fried_parameter_from_Cn_squared(Cn_squared, 1.0)
    f0 = 1.0 / L0

    def psd(fx, fy):
        return 0.0229 * r0 ** (-5.0 / 3) * (fx ** 2 + fy ** 2 + f0 ** 2) ** (-11.0 / 6)

    df1 = 1.0 / (N * delta)
    k1 = np.fft.fftfreq(N, d=1.0 / N)                   # integer bin index
    KX, KY = np.meshgrid(k1, k1)
    C1 = np.sqrt(psd(KX * df1, KY * df1)) * df1
    C1[(np.abs(KX) <= 1) & (np.abs(KY) <= 1)] = 0.0
    noise1 = rng.standard_normal((N, N)) + 1j * rng.standard_normal((N, N))
    # sum_k c_k exp(2 pi i f_k . x): grid coordinates are (i + 0.5 - N/2) delta
    shift = np.exp(2j * np.pi * k1 * df1 * grid.xs[0])
    spec = C1 * noise1 * shift[None, :] * shift[:, None]
    screen = (np.fft.ifft2(spec) * N * N).real

    n2 = 3 * oversampling
    df2 = df1 / oversampling
    f2 = (np.arange(n2) + 0.5 - n2 / 2) * df2
    FX, FY = np.meshgrid(f2, f2)
    C2 = np.sqrt(psd(FX, FY)) * df2
    noise2 = rng.standard_normal((n2, n2)) + 1j * rng.standard_normal((n2, n2))
    W = np.exp(2j * np.pi * np.outer(grid.xs, f2))      # [N, n2]
    screen += (W @ (C2 * noise2) @ W.T).real            # [y, x]
    return screen.ravel()


class InfiniteAtmosphericLayer:
    """hcipy.InfiniteAtmosphericLayer(grid, Cn2, L0, velocity) (AO_env.py:370) with
    stencil_length=2, use_interpolation=False.  Achromatic screen S with
    phase_for(lambda) = S / lambda; forward: E * exp(i S / lambda).

    Autoregressive extrusion (Assemat et al. 2006 as implemented by hcipy):
    new_line = A z + B xi sqrt(Cn2),  z = screen[stencil], xi ~ N(0, I),
    A = C_xz C_zz^-1 (Tikhonov, rcond 1e-10), B = U sqrt(S) from the SVD of
    C_xx - A C_zx, covariances from the von-Karman phase covariance at
    r0(Cn2=1, lambda=1).  'right'/'top' extrusions run on the 180-degree rotated
    screen.  Noise can be injected through ``noise_source`` (callable returning
    a length-N vector) for parity tests.
    """

    def __init__(self, grid, Cn_squared, L0, velocity, rng, stencil_length=2, initial_screen=None):
        self.grid = grid
        self.Cn_squared = float(Cn_squared)
        self.L0 = float(L0)
        self.velocity = np.array([velocity, 0.0], dtype=np.float64)   # scalar -> [v, 0]
        self.rng = rng
        self.stencil_length = stencil_length
        self.noise_source = None
        self._make_stencils()
        self._make_AB_matrices()
        if initial_screen is None:
            self._make_initial_phase_screen()
        else:
            self.achromatic_screen = np.array(initial_screen, dtype=np.float64).ravel()
        self.center = np.zeros(2)
        self._t = 0.0

    def _make_stencils(self):
        nx, ny = self.grid.dims
        sl = self.stencil_length
        # vertical: rows [0, sl) plus one geometric-offset pixel per column
        sb = np.zeros((ny, nx), dtype=bool)
        sb[:sl, :] = True
        for i, n in enumerate(self.rng.geometric(0.5, nx)):
            sb[(n + sl - 1) % ny, i] = True
        self.stencil_bottom = sb.ravel()
        # horizontal: columns [0, sl) plus one geometric-offset pixel per row
        st = np.zeros((ny, nx), dtype=bool)
        st[:, :sl] = True
        for i, n in enumerate(self.rng.geometric(0.5, ny)):
            st[i, (n + sl - 1) % nx] = True
        self.stencil_left = st.ravel()

    def _AB(self, stencil, new_x, new_y):
        cov_fn = hcipy_phase_covariance_von_karman(hcipy_fried_parameter_from_Cn_squared(1, 1), self.L0)
        x = np.concatenate((self.grid.x[stencil], new_x))
        y = np.concatenate((self.grid.y[stencil], new_y))
        r = np.hypot(x[None, :] - x[:, None], y[None, :] - y[:, None])
        cov = cov_fn(r)
        n = int(stencil.sum())
        cov_zz, cov_xz = cov[:n, :n], cov[n:, :n]
        cov_zx, cov_xx = cov[:n, n:], cov[n:, n:]
        A = cov_xz @ hcipy_inverse_tikhonov(cov_zz, 1e-10)
        BBt = cov_xx - A @ cov_zx
        U, S, _ = np.linalg.svd(BBt)
        B = U * np.sqrt(S)
        return A, B

    def _make_AB_matrices(self):
        g = self.grid
        d = g.delta
        self.A_vertical, self.B_vertical = self._AB(
            self.stencil_bottom, g.xs, np.full(g.dims[0], g.ys[0] - d[1]))
        self.A_horizontal, self.B_horizontal = self._AB(
            self.stencil_left, np.full(g.dims[1], g.xs[0] - d[0]), g.ys)

    def _make_initial_phase_screen(self):
        self.achromatic_screen = von_karman_screen(self.grid, self.Cn_squared, self.L0, self.rng)

    def reset(self):
        self._make_initial_phase_screen()
        self.center = np.zeros(2)
        self._t = 0.0

    def _draw(self, n):
        if self.noise_source is not None:
            return np.asarray(self.noise_source(n), dtype=np.float64)
        return self.rng.standard_normal(n)

    def _extrude(self, where):
        flipped = where in ('top', 'right')
        horizontal = where in ('left', 'right')
        screen = self.achromatic_screen[::-1] if flipped else self.achromatic_screen
        if horizontal:
            stencil, A, B = self.stencil_left, self.A_horizontal, self.B_horizontal
        else:
            stencil, A, B = self.stencil_bottom, self.A_vertical, self.B_vertical
        stencil_data = screen[stencil]
        random_data = self._draw(B.shape[1])
        new_slice = A @ stencil_data + (B @ random_data) * np.sqrt(self.Cn_squared)
        screen = screen.reshape(self.grid.shape)
        if horizontal:
            screen = np.hstack((new_slice[:, None], screen[:, :-1]))
        else:
            screen = np.vstack((new_slice[None, :], screen[:-1, :]))
        if flipped:
            self.achromatic_screen = screen[::-1, ::-1].ravel()
        else:
            self.achromatic_screen = screen.ravel()

    @property
    def t(self):
        return self._t

    @t.setter
    def t(self, t):
        self.evolve_until(t)
        self._t = t

    def evolve_until(self, t):
        d = self.grid.delta
        old_center = np.round(self.center / d).astype(int)
        self.center = self.velocity * t
        new_center = np.round(self.center / d).astype(int)
        delta = new_center - old_center
        for _ in range(abs(delta[0])):
            self._extrude('left' if delta[0] < 0 else 'right')
        for _ in range(abs(delta[1])):
            self._extrude('bottom' if delta[1] < 0 else 'top')

    def phase_for(self, wavelength):
        return self.achromatic_screen / wavelength

    def forward(self, wf):
        wf2 = wf.copy()
        wf2.electric_field *= np.exp(1j * self.phase_for(wf.wavelength))
        return wf2

    __call__ = forward


# --------------------------------------------------------------------------
# hcipy.optics.StepIndexFiber (LP modes)
# --------------------------------------------------------------------------
def _lp_eigenvalue_equation(u, m, V):
    w = np.sqrt(V ** 2 - u ** 2)
    return special.jv(m, u) / (u * special.jv(m + 1, u)) - special.kn(m, w) / (w * special.kn(m + 1, w))


def _lp_find_branch_cuts(m, V):
    """Roots u in (0, V) of the LP characteristic equation (true sign changes only:
    the poles at the zeros of J_{m+1} are rejected by a residual test)."""
    from scipy.optimize import brentq
    num_steps = 5001
    eps = 1e-9
    u = np.linspace(eps, V - eps, num_steps)
    with np.errstate(all='ignore'):
        f = _lp_eigenvalue_equation(u, m, V)
    roots = []
    for i in range(num_steps - 1):
        if np.isfinite(f[i]) and np.isfinite(f[i + 1]) and f[i] * f[i + 1] < 0:
            r = brentq(_lp_eigenvalue_equation, u[i], u[i + 1], args=(m, V), xtol=1e-15, rtol=1e-15)
            if abs(_lp_eigenvalue_equation(r, m, V)) < 1e-6:
                roots.append(r)
    if not roots:
        return None
    roots = np.array(roots)
    return roots, np.sqrt(V ** 2 - roots ** 2)


def hcipy_make_LP_modes(grid, V_number, core_radius, wavelength):
    """hcipy.make_LP_modes: m = 0,1,2,... until no solution; for each root the cos
    copy (m) then the sin copy (-m) for m > 0; each numerically normalised to
    sum(mode^2 w) = 1 on ``grid``.  Returns modes [P, J] and beta [J]."""
    R = np.hypot(grid.x, grid.y) / core_radius
    Theta = np.arctan2(grid.y, grid.x)
    k0 = 2 * np.pi / wavelength
    modes, betas = [], []
    m = 0
    while True:
        sol = _lp_find_branch_cuts(m, V_number)
        if sol is None:
            break
        for ui, wi in zip(*sol):
            mask = R < 1
            radial = np.zeros_like(R)
            radial[mask] = special.jv(m, ui * R[mask])
            radial[~mask] = special.jv(m, ui) / special.kn(m, wi) * special.kn(m, wi * R[~mask])
            for mi in ([m, -m] if m > 0 else [m]):
                az = np.cos(mi * Theta) if mi >= 0 else np.sin(mi * Theta)
                prof = radial * az
                prof = prof / np.sqrt(np.sum(prof * prof * grid.weight))
                modes.append(prof)
                betas.append(np.sqrt(k0 ** 2 - (ui / core_radius) ** 2))
        m += 1
    return np.stack(modes, axis=1), np.array(betas)


class StepIndexFiber:
    """hcipy.StepIndexFiber(core_radius, NA, fiber_length) (AO_env.py:393,471):
    out = M ((M^T (E w)) exp(i beta L))."""

    def __init__(self, core_radius, NA, fiber_length):
        self.core_radius, self.NA, self.fiber_length = core_radius, NA, fiber_length
        self._cache = {}

    def V(self, wavelength):
        return 2 * np.pi / wavelength * self.core_radius * self.NA

    def instance(self, grid, wavelength):
        key = (id(grid), float(wavelength))
        if key not in self._cache:
            self._cache[key] = hcipy_make_LP_modes(grid, self.V(wavelength), self.core_radius, wavelength)
        return self._cache[key]

    def forward(self, wf):
        M, beta = self.instance(wf.grid, wf.wavelength)
        c = M.T @ (wf.electric_field * wf.grid.weight)
        out = M @ (c * np.exp(1j * beta * self.fiber_length))
        return Wavefront(out, wf.grid, wf.wavelength)


# --------------------------------------------------------------------------
# misc hcipy / skimage
# --------------------------------------------------------------------------
def hcipy_get_strehl_from_focal(img, ref_img):
    """AO_env.py:482."""
    return img[np.argmax(ref_img)] / ref_img.max()


def skimage_ssim_1d(im1, im2, data_range, win_size=7, K1=0.01, K2=0.03):
    """skimage.metrics.structural_similarity 0.22 on 1-D float64 arrays
    (AO_env.py:495): uniform 7-window, sample covariance, crop 3, mean."""
    from scipy.ndimage import uniform_filter
    im1 = np.asarray(im1, dtype=np.float64)
    im2 = np.asarray(im2, dtype=np.float64)
    if np.any((np.asarray(im1.shape) - win_size) < 0):
        raise ValueError("win_size exceeds image extent. Either ensure that your images are "
                         "at least 7x7; or pass win_size explicitly in the function call, with "
                         "an odd value less than or equal to the smaller side of your images.")
    NP = win_size ** im1.ndim
    cov_norm = NP / (NP - 1)
    ux = uniform_filter(im1, size=win_size)
    uy = uniform_filter(im2, size=win_size)
    uxx = uniform_filter(im1 * im1, size=win_size)
    uyy = uniform_filter(im2 * im2, size=win_size)
    uxy = uniform_filter(im1 * im2, size=win_size)
    vx = cov_norm * (uxx - ux * ux)
    vy = cov_norm * (uyy - uy * uy)
    vxy = cov_norm * (uxy - ux * uy)
    C1 = (K1 * data_range) ** 2
    C2 = (K2 * data_range) ** 2
    S = ((2 * ux * uy + C1) * (2 * vxy + C2)) / ((ux ** 2 + uy ** 2 + C1) * (vx + vy + C2))
    pad = (win_size - 1) // 2
    return float(S[pad:-pad].mean(dtype=np.float64))


def hcipy_large_poisson(lam, rng, thresh=1e6):
    """hcipy.large_poisson (AO_env.py:274): Poisson below ``thresh``, rounded normal
    approximation above."""
    lam = np.asarray(lam, dtype=np.float64)
    large = lam > thresh
    small = ~large
    n = np.zeros(lam.shape)
    n[large] = np.round(lam[large] + rng.standard_normal(int(large.sum())) * np.sqrt(lam[large]))
    n[small] = rng.poisson(lam[small])
    return n


# --------------------------------------------------------------------------
# Shack-Hartmann WFS (hcipy SquareShackHartmannWavefrontSensorOptics + estimator)
# --------------------------------------------------------------------------
class ShackHartmann:
    """AO_env.py:396-465 set-up and the optics used by SH_step (AO_env.py:263-279).

    Magnifier(m): grid scaled by m, E / m (power conserving).  MLA: lenslet
    centres arange(-D_sh, D_sh, D_sh/12) per axis; every pixel belongs to its
    nearest lenslet; phase exp(-i k d^2 / (2 f)), f = f_number * pitch.
    FresnelPropagator(f): angular spectrum with a 2x zero-padded FFT,
    H = exp(i k z) exp(-i z |q|^2 / (2 k)).  NoiselessDetector: image = power * dt
    carried on ``detector_grid`` (the science focal grid, AO_env.py:412).
    Estimator: flux-weighted centroids per selected lenslet minus lenslet centre.
    """

    def __init__(self, pupil_grid, detector_grid, magnification, f_number, num_lenslets, sh_diameter):
        self.mag = magnification
        self.grid = pupil_grid.scaled(magnification)
        self.detector_grid = detector_grid
        pitch = float(sh_diameter) / num_lenslets
        c = np.arange(-sh_diameter, sh_diameter, pitch)
        self.mla_x = np.tile(c, c.size)
        self.mla_y = np.repeat(c, c.size)
        self.focal_length = f_number * pitch
        ix = np.argmin(np.abs(self.grid.xs[:, None] - c[None, :]), axis=1)
        iy = np.argmin(np.abs(self.grid.ys[:, None] - c[None, :]), axis=1)
        self.mla_index = (iy[:, None] * c.size + ix[None, :]).ravel()
        self.mla_opd = (-1 / (2 * self.focal_length)) * (
            (self.grid.x - self.mla_x[self.mla_index]) ** 2 + (self.grid.y - self.mla_y[self.mla_index]) ** 2)
        self.estimation_subapertures = np.unique(self.mla_index)
        self._tf = {}

    def transfer_function(self, wavelength):
        if wavelength not in self._tf:
            N = self.grid.dims[0]
            d = self.grid.delta[0]
            q = 2 * np.pi * np.fft.fftfreq(2 * N, d=d)
            k = 2 * np.pi / wavelength
            q2 = q[None, :] ** 2 + q[:, None] ** 2
            self._tf[wavelength] = np.exp(-0.5j * (self.focal_length / k) * q2) * np.exp(1j * k * self.focal_length)
        return self._tf[wavelength]

    def optics(self, wf):
        """magnifier -> micro-lens array -> Fresnel propagation by f."""
        N = self.grid.dims[0]
        E = wf.electric_field / self.mag
        E = E * np.exp(1j * self.mla_opd * 2 * np.pi / wf.wavelength)
        pad = np.zeros((2 * N, 2 * N), dtype=np.complex128)
        o = N // 2
        pad[o:o + N, o:o + N] = E.reshape(N, N)
        out = np.fft.ifft2(np.fft.fft2(pad) * self.transfer_function(wf.wavelength))
        return Wavefront(out[o:o + N, o:o + N].ravel(), self.grid, wf.wavelength)

    def estimate(self, image):
        from scipy import ndimage
        idx = self.estimation_subapertures
        gx, gy = self.detector_grid.x, self.detector_grid.y
        fluxes = ndimage.sum(image, self.mla_index, idx)
        sum_x = ndimage.sum(image * gx, self.mla_index, idx)
        sum_y = ndimage.sum(image * gy, self.mla_index, idx)
        cx = sum_x / fluxes
        cy = sum_y / fluxes
        return np.array((cx, cy)) - np.array((self.mla_x[idx], self.mla_y[idx]))


# --------------------------------------------------------------------------
# The environment (AO_env.py:16-503)
# --------------------------------------------------------------------------
class OracleAOEnv:
    """Line-by-line restatement of ``AOEnv`` (AO_env.py:16-503) on the classes above.

    Extra, parity-only arguments: ``seed`` (explicit Generator instead of the
    global NumPy RNG), ``initial_screen`` (inject S), ``verbose``.
    ``step(action, extrusion_noise=None)``: optional [n_ext, N] normals consumed
    in order by the extrusions of that step.
    """

    def __init__(self, atm_type='quasi_static', atm_vel=0, atm_fried=0.15, act_type='num_actuators',
                 act_dim=64, obs_dim=2, rew_type='strehl_ratio', rew_threshold=None,
                 timesteps_per_episode=20, flat_mirror_start_per_episode=True, SH_operation=False,
                 seed=0, initial_screen=None, verbose=False, num_pupil_pixels=240,
                 num_focal_pixels_fiber=128):
        self.atm_type, self.rew_type, self.act_type = atm_type, rew_type, act_type
        self.flat_mirror_start_per_episode = flat_mirror_start_per_episode
        self.rew_threshold = rew_threshold
        self.SH_operation = SH_operation
        self.rng = np.random.default_rng(seed)
        self.verbose = verbose
        self._initial_screen = initial_screen
        self.parameters_init(act_dim, atm_vel, obs_dim, timesteps_per_episode, atm_fried,
                             num_pupil_pixels, num_focal_pixels_fiber)
        aperture, pupil_grid = self.pupil_simulation()
        focal_grid = self.incoming_wavefront(aperture, pupil_grid)
        dm_modes = self.DM_function(act_type, pupil_grid)
        self.atmospheric_turbulence(pupil_grid)
        self.fiber_coupling()
        self.aperture, self.pupil_grid, self.focal_grid, self.dm_modes = aperture, pupil_grid, focal_grid, dm_modes
        if self.SH_operation:
            self.shack_hartmann_init(pupil_grid, focal_grid, aperture, dm_modes)
        self.timestep = 0
        self.episode_no = 0

    # AO_env.py:197-251
    def parameters_init(self, act_dim, velocity_value, obs_dim, timesteps_per_episode, fried_parameter,
                        num_pupil_pixels, num_focal_pixels_fiber):
        if self.atm_type in ('quasi_static', 'semi_dynamic') and velocity_value != 0:
            if self.verbose:
                print('In ' + self.atm_type + ' atmospheric condition, the velocity value should be zero.')
                print('therefore velocity value is changed to zero')
            velocity_value = 0
        elif self.atm_type == 'dynamic' and velocity_value == 0:
            if self.verbose:
                print('In ' + self.atm_type + ' atmospheric condition, the velocity value cannot be zero.')
                print('therefore velocity value is changed to 1 m/s')
            velocity_value = 1
        self.telescope_diameter = 0.5
        self.num_pupil_pixels = num_pupil_pixels
        self.wavelength_wfs = 1.5e-6
        self.wavelength_sci = 2.2e-6
        self.num_modes = act_dim
        self.delta_t = 1e-3
        self.max_steps = timesteps_per_episode
        self.velocity = velocity_value
        self.fried_parameter = fried_parameter
        self.outer_scale = 10
        self.D_pupil_fiber = 0.5
        self.num_pupil_pixels_fiber = 128
        self.num_focal_pixels_fiber = num_focal_pixels_fiber
        self.num_focal_pixels_fiber_subsample = obs_dim
        self.multimode_fiber_core_radius = 25 * 1e-6
        self.singlemode_fiber_core_radius = 4.5 * 1e-6
        self.fiber_NA = 0.14
        self.fiber_length = 10
        self.f_number = 50
        self.num_lenslets = 12
        self.sh_diameter = 5e-3
        self.stellar_magnitude = -5

    # AO_env.py:293-303
    def pupil_simulation(self):
        pupil_grid = hcipy_make_pupil_grid(self.num_pupil_pixels, self.telescope_diameter)
        aperture = hcipy_make_circular_aperture(self.telescope_diameter)(pupil_grid)
        return aperture, pupil_grid

    # AO_env.py:306-336
    def incoming_wavefront(self, aperture, pupil_grid):
        spatial_resolution = self.wavelength_sci / self.telescope_diameter
        focal_grid = hcipy_make_focal_grid(q=4, num_airy=30, spatial_resolution=spatial_resolution)
        self.propagator = FraunhoferPropagator(pupil_grid, focal_grid)
        wf = Wavefront(aperture, pupil_grid, self.wavelength_sci)
        wf.total_power = 1
        self.unaberrated_PSF = self.propagator.forward(wf).power
        zero_magnitude_flux = 3.9e10
        self.wf_wfs = Wavefront(aperture, pupil_grid, self.wavelength_wfs)
        self.wf_wfs.total_power = zero_magnitude_flux * 10 ** (-self.stellar_magnitude / 2.5)
        self.wf_wfs_fiber = Wavefront(aperture, pupil_grid, self.wavelength_wfs)
        self.wf_wfs_fiber.total_power = 1
        self.wf_sci = Wavefront(aperture, pupil_grid, self.wavelength_sci)
        self.wf_sci.total_power = zero_magnitude_flux * 10 ** (-self.stellar_magnitude / 2.5)
        return focal_grid

    # AO_env.py:339-358
    def DM_function(self, act_type, pupil_grid):
        if act_type == 'zernike':
            dm_modes = hcipy_make_zernike_basis(self.num_modes, self.telescope_diameter, pupil_grid)
        else:
            dm_modes = hcipy_make_disk_harmonic_basis(pupil_grid, self.num_modes, self.telescope_diameter, 'neumann')
        dm_modes = [mode / np.ptp(mode) for mode in dm_modes]
        self.deformable_mirror = DeformableMirror(dm_modes)
        self.deformable_mirror.flatten()
        return dm_modes

    # AO_env.py:361-370
    def atmospheric_turbulence(self, pupil_grid):
        Cn_squared = hcipy_Cn_squared_from_fried_parameter(self.fried_parameter, self.wavelength_sci)
        self.layer = InfiniteAtmosphericLayer(pupil_grid, Cn_squared, self.outer_scale, self.velocity,
                                              self.rng, initial_screen=self._initial_screen)

    # AO_env.py:373-393
    def fiber_coupling(self):
        pupil_grid_fiber = hcipy_make_pupil_grid(self.num_pupil_pixels_fiber, self.D_pupil_fiber)
        D_focus_fiber = 2.1 * self.multimode_fiber_core_radius
        focal_grid_fiber = hcipy_make_pupil_grid(self.num_focal_pixels_fiber, D_focus_fiber)
        focal_grid_fiber_subsample = hcipy_make_pupil_grid(self.num_focal_pixels_fiber_subsample, D_focus_fiber)
        focal_length = self.D_pupil_fiber / (2 * self.fiber_NA)
        self.propagator_fiber = FraunhoferPropagator(pupil_grid_fiber, focal_grid_fiber, focal_length=focal_length)
        self.propagator_fiber_subsample = FraunhoferPropagator(pupil_grid_fiber, focal_grid_fiber_subsample,
                                                               focal_length=focal_length)
        self.single_mode_fiber = StepIndexFiber(self.singlemode_fiber_core_radius, self.fiber_NA, self.fiber_length)

    # AO_env.py:396-465
    def shack_hartmann_init(self, pupil_grid, focal_grid, aperture, dm_modes):
        magnification = self.sh_diameter / self.telescope_diameter
        self.shwfs = ShackHartmann(pupil_grid, focal_grid, magnification, self.f_number,
                                   self.num_lenslets, self.sh_diameter)
        sh = self.shwfs
        wf_camera = Wavefront(aperture, pupil_grid, self.wavelength_wfs)
        image_ref = sh.optics(wf_camera).power * 1.0
        from scipy import ndimage
        fluxes = ndimage.sum(image_ref, sh.mla_index, sh.estimation_subapertures)
        flux_limit = fluxes.max() * 0.5
        sh.estimation_subapertures = sh.estimation_subapertures[fluxes > flux_limit]
        self.slopes_ref = sh.estimate(image_ref)
        self.deformable_mirror_shack = DeformableMirror(dm_modes)
        probe_amp = 0.01 * self.wavelength_wfs
        response_matrix = []
        wf_cal = Wavefront(aperture, pupil_grid, self.wavelength_wfs)
        wf_cal.total_power = 1
        for i in range(self.num_modes):
            slope = 0
            amps = [-probe_amp, probe_amp]
            for amp in amps:
                self.deformable_mirror_shack.flatten()
                self.deformable_mirror_shack.actuators[i] = amp
                dm_wf = self.deformable_mirror_shack.forward(wf_cal)
                image = sh.optics(dm_wf).power * 1.0
                slopes = sh.estimate(image)
                slope += amp * slopes / np.var(amps)
            response_matrix.append(slope.ravel())
        self.response_matrix = np.stack(response_matrix, axis=1)      # [2 N_sub, K]
        self.reconstruction_matrix = hcipy_inverse_tikhonov(self.response_matrix, rcond=1e-3)
        self.deformable_mirror_shack.flatten()
        # NB (AO_env.py:446-447): the calibration loop leaves the LAST probe on the SH
        # mirror; restate that exactly.
        self.deformable_mirror_shack.actuators[self.num_modes - 1] = probe_amp

    # AO_env.py:74-103
    def reset(self, seed=None, options=None):
        if self.atm_type == 'semi_dynamic':
            self.layer.reset()
        if self.flat_mirror_start_per_episode:
            self.deformable_mirror.flatten()
        self.timestep_render = 0
        self.layer.t = self.timestep * self.delta_t
        self.phase_screen_opd = self.layer.phase_for(self.wavelength_wfs) * (self.wavelength_wfs / (2 * np.pi)) * 1e6
        wf_wfs_after_atmos = self.layer(self.wf_wfs_fiber)
        wf_wfs_after_dm = self.deformable_mirror(wf_wfs_after_atmos)
        self.wf_wfs_after_foc = self.propagator_fiber(wf_wfs_after_dm)
        self.wf_wfs_after_foc_subsample = self.propagator_fiber_subsample(wf_wfs_after_dm)
        state = self.wf_wfs_after_foc_subsample.power
        self.last_obs_f64 = state.copy()
        return np.array(state, dtype=np.float16), {}

    # AO_env.py:106-153
    def step(self, action, extrusion_noise=None):
        trunc = False
        if self.SH_operation:
            self.deformable_mirror.actuators = np.array(action, dtype=np.float64)
        else:
            with np.errstate(all='ignore'):
                self.deformable_mirror.actuators = np.asarray(action) / (np.arange(self.num_modes) + 10)
                self.deformable_mirror.actuators = self.deformable_mirror.actuators * (
                    0.1 * self.wavelength_sci / (np.std(self.deformable_mirror.surface)))
        self.timestep += 1
        self.timestep_render += 1
        if extrusion_noise is not None:
            queue = list(np.asarray(extrusion_noise, dtype=np.float64))
            self.layer.noise_source = lambda n: queue.pop(0)
        self.layer.t = self.timestep * self.delta_t
        self.layer.noise_source = None
        self.phase_screen_opd = self.layer.phase_for(self.wavelength_wfs) * (self.wavelength_wfs / (2 * np.pi)) * 1e6
        wf_wfs_after_atmos = self.layer(self.wf_wfs_fiber)
        wf_wfs_after_dm = self.deformable_mirror(wf_wfs_after_atmos)
        self.wf_wfs_after_foc = self.propagator_fiber(wf_wfs_after_dm)
        self.wf_wfs_after_foc_subsample = self.propagator_fiber_subsample(wf_wfs_after_dm)
        next_state = self.wf_wfs_after_foc_subsample.power
        self.last_obs_f64 = next_state.copy()
        reward, rew_fiber = self.reward_function()
        if self.timestep_render == self.max_steps:
            done = True
            self.episode_no += 1
        else:
            done = False
        return np.array(next_state, dtype=np.float16), reward, done, trunc, {"power": float(rew_fiber)}

    def num_extrusions_for_next_step(self):
        """How many extrusions the NEXT step() will perform (for sizing injected noise)."""
        d = self.pupil_grid.delta[0]
        old = int(np.round(self.layer.center[0] / d))
        new = int(np.round(self.layer.velocity[0] * (self.timestep + 1) * self.delta_t / d))
        return abs(new - old)

    # AO_env.py:468-503
    def reward_function(self):
        wf_smf = self.single_mode_fiber.forward(self.wf_wfs_after_foc)
        rew_fiber = wf_smf.total_power
        if self.rew_type == 'strehl_ratio':
            self.wf_sci_focal_plane = self.propagator(self.deformable_mirror(self.layer(self.wf_sci)))
            strehl_ratio = hcipy_get_strehl_from_focal(
                self.wf_sci_focal_plane.power, self.unaberrated_PSF * self.wf_wfs.total_power) * 100
            self.last_strehl = strehl_ratio
            reward = -(100 - strehl_ratio)
        elif self.rew_type == 'smf_ssim':
            focal_power = self.wf_wfs_after_foc_subsample.power
            ref_power = np.zeros(self.num_focal_pixels_fiber_subsample ** 2)
            ref_power[int(self.num_focal_pixels_fiber_subsample ** 2 / 2)] = 2.8
            data_range = ref_power.max() - ref_power.min()
            ssim_score = skimage_ssim_1d(focal_power, ref_power, data_range=data_range)
            self.last_ssim = ssim_score
            alpha = 0.8
            reward = alpha * rew_fiber + (1 - alpha) * ssim_score
        if self.rew_threshold is not None and reward < self.rew_threshold:
            reward = -1.0
        return reward, rew_fiber

    # AO_env.py:254-290
    def SH_step(self, poisson=True, poisson_image=None):
        """``poisson_image``: inject the noisy camera image (parity); ``poisson=False``:
        skip photon noise (deterministic check)."""
        wf_wfs_after_atmos = self.layer(self.wf_wfs)
        wf_wfs_after_dm = self.deformable_mirror_shack(wf_wfs_after_atmos)
        wf_wfs_on_sh = self.shwfs.optics(wf_wfs_after_dm)
        wfs_image = wf_wfs_on_sh.power * self.delta_t
        self.last_sh_image = wfs_image.copy()
        if poisson_image is not None:
            wfs_image = np.asarray(poisson_image, dtype=np.float64)
        elif poisson:
            wfs_image = hcipy_large_poisson(wfs_image, self.rng).astype('float')
        slopes = self.shwfs.estimate(wfs_image + 1e-10)
        slopes = slopes - self.slopes_ref
        slopes = slopes.ravel()
        gain = 0.3
        leakage = 0.01
        self.deformable_mirror_shack.actuators = (1 - leakage) * self.deformable_mirror_shack.actuators \
            - gain * self.reconstruction_matrix.dot(slopes)
        action = self.deformable_mirror_shack.actuators
        return action, np.array([1])
