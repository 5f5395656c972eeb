"""Known-answer tests of the grid stages (closed forms any correct implementation of the cited shaders must give), evaluated on the
oracle (CPU) and, with the same assertions, on the CUDA path (GPU).

  divergence_compute.comp:28-86          rhs of a linear velocity field u = A x + b inside the fluid  = trace(A)
  divergence_remove.comp:19-49           u_c <- u_c - (p(g) - p(g + e_c)): a linear pressure p = k . x shifts every fluid-fluid face by +k_c
  density_projection_position_change.comp:18-51   D_c = (p(g + e_c) - p(g)) * dt = k_c * dt on fluid-fluid faces, 0 next to SOLID
  extrapolate_velocity.comp:26-90        a constant field on the valid faces is copied to every face of the first ring around them
  density_projection_gather_error.comp:41-199     a regular 2x2x2 lattice (8 particles per cell) has density 8 at every interior cell
                                                   centre: rhs = 0; a missing layer of cells next to it raises the rhs accordingly
"""
import numpy as np
import pytest

from oracle import oracle as O
from tests.util import DT, near_fluid

N = 32
A = np.array([[0.5, -1.0, 2.0], [1.5, 0.25, -0.5], [-2.0, 1.0, -0.75]])
B = np.array([4.0, -6.0, 2.5])
K = np.array([3.0, -2.0, 0.5])  # pressure gradient per cell


def block_marker(lo=8, hi=24):
    m = np.full((N, N, N), O.AIR, dtype=np.int8)
    m[0], m[-1], m[:, 0], m[:, -1], m[:, :, 0], m[:, :, -1] = 0, 0, 0, 0, 0, 0
    m[lo:hi, lo:hi, lo:hi] = O.FLUID
    return m


def face_field(c):
    z, y, x = np.meshgrid(np.arange(N), np.arange(N), np.arange(N), indexing="ij")
    q = np.stack([x, y, z], axis=-1) + 0.5
    q[..., c] += 0.5
    return (q @ A.T + B)[..., c].astype(np.float32)


def cell_pressure():
    z, y, x = np.meshgrid(np.arange(N), np.arange(N), np.arange(N), indexing="ij")
    return ((x + 0.5) * K[0] + (y + 0.5) * K[1] + (z + 0.5) * K[2]).astype(np.float32)


def fluid_fluid_faces(m, c):
    fl = m == O.FLUID
    nb = np.roll(fl, -1, axis=2 - c)
    return fl & nb


def lattice_particles(lo, hi):
    """8 particles per cell at (0.25, 0.75)^3 in the cells [lo, hi)^3."""
    c = np.arange(lo, hi)
    off = np.array([0.25, 0.75])
    ax = (c[:, None] + off[None, :]).ravel()
    z, y, x = np.meshgrid(ax, ax, ax, indexing="ij")
    return np.stack([x.ravel(), y.ravel(), z.ravel(), np.zeros(x.size)], axis=1).astype(np.float32)


class OracleBackend:
    def __init__(self, nparticles=8):
        self.f = O.OracleFluid(N, N, N, nparticles)
        self.f.set_gravity_grid([0.0, 0.0, 0.0])

    def put(self, marker=None, u=None, p=None, which=0):
        if marker is not None:
            self.f.grid(O.ARR_MARKER)[:] = marker
        if u is not None:
            for c, a in enumerate((O.ARR_UX, O.ARR_UY, O.ARR_UZ)):
                self.f.grid(a)[:] = u[c]
        if p is not None:
            self.f.grid(O.ARR_P_VEL if which == 0 else O.ARR_P_DEN)[:] = p

    def particles(self, pos):
        self.f.set_particles(pos)

    def run(self, a, b, dt=DT):
        self.f.step_stages(dt, a, b)

    def u(self):
        return [self.f.grid(a).copy() for a in (O.ARR_UX, O.ARR_UY, O.ARR_UZ)]

    def marker(self):
        return self.f.grid(O.ARR_MARKER).copy()

    def residual(self):
        return self.f.grid(O.ARR_RESIDUAL).copy()


class CudaBackend:
    def __init__(self, nparticles=8):
        import blub_b200
        from blub_b200 import fluid as F
        self.F = F
        self.f = blub_b200.HybridFluid(N, N, N, nparticles)
        self.f.set_gravity_grid([0.0, 0.0, 0.0])

    def put(self, marker=None, u=None, p=None, which=0):
        F = self.F
        if marker is not None:
            self.f.upload_grid(F.TAP_MARKER, marker)
        if u is not None:
            for c, t in enumerate((F.TAP_UX, F.TAP_UY, F.TAP_UZ)):
                self.f.upload_grid(t, np.ascontiguousarray(u[c], dtype=np.float32))
        if p is not None:
            self.f.upload_grid(F.TAP_P_VEL if which == 0 else F.TAP_P_DEN, p)

    def particles(self, pos):
        self.f.set_particles(pos)

    def run(self, a, b, dt=DT):
        self.f.step_stages(dt, a, b)

    def u(self):
        return [self.f.download_grid(t) for t in (self.F.TAP_UX, self.F.TAP_UY, self.F.TAP_UZ)]

    def marker(self):
        return self.f.download_grid(self.F.TAP_MARKER)

    def residual(self):
        return self.f.download_grid(self.F.TAP_RESIDUAL)


BACKENDS = [pytest.param(OracleBackend, id="oracle"), pytest.param(CudaBackend, id="cuda", marks=pytest.mark.gpu)]


@pytest.mark.parametrize("backend", BACKENDS)
def test_divergence_of_a_linear_field_is_its_trace(backend):
    b, m = backend(), block_marker()
    b.put(marker=m, u=[face_field(c) for c in range(3)])
    b.run(1, 2)
    rhs = b.residual()[m == O.FLUID]
    assert np.abs(rhs - np.trace(A)).max() <= 2e-4  # differences of O(100) fp32 numbers


@pytest.mark.parametrize("backend", BACKENDS)
def test_projection_subtracts_the_pressure_gradient(backend):
    b, m = backend(), block_marker()
    b.put(marker=m, u=[np.full((N, N, N), B[c], dtype=np.float32) for c in range(3)], p=cell_pressure(), which=0)
    b.run(4, 5)
    u = b.u()
    for c in range(3):
        ff = fluid_fluid_faces(m, c)
        assert ff.sum() > 3000 and np.abs(u[c][ff] - (B[c] + K[c])).max() <= 1e-4
        # faces between the fluid and the AIR around it see p = 0 outside (free surface); faces touching no FLUID cell are zeroed -- within
        # one cell of the fluid (further away nobody reads them: the CUDA passes do not visit those cells, see tests/util.py:near_fluid)
        fl = m == O.FLUID
        touches = fl | np.roll(fl, -1, axis=2 - c)
        assert (u[c][~touches & near_fluid(m)] == 0).all()


@pytest.mark.parametrize("backend", BACKENDS)
def test_position_change_is_the_pressure_gradient_times_dt(backend):
    b, m = backend(), block_marker(lo=1, hi=24)  # the block touches the SOLID walls at 0
    b.put(marker=m, p=cell_pressure(), which=1)
    b.run(11, 12)
    u = b.u()
    for c in range(3):
        ff = fluid_fluid_faces(m, c)
        assert np.abs(u[c][ff] - np.float32(K[c]) * np.float32(DT)).max() <= 2e-6
        solid = (m == O.SOLID) | (np.roll(m, -1, axis=2 - c) == O.SOLID)
        assert (u[c][solid & near_fluid(m)] == 0).all()  # Neumann: nothing moves through a wall


@pytest.mark.parametrize("backend", BACKENDS)
def test_extrapolation_copies_a_constant_field_into_the_first_ring(backend):
    b, m = backend(), block_marker()
    fl = m == O.FLUID
    rng = np.random.default_rng(0)
    u = []
    for c in range(3):
        valid = fl | np.roll(fl, -1, axis=2 - c)
        f = rng.uniform(50, 60, (N, N, N)).astype(np.float32)  # garbage on the invalid faces
        f[valid] = np.float32(B[c])
        u.append(f)
    garbage = [f.copy() for f in u]
    b.put(marker=m, u=u)
    b.run(8, 9)  # set_boundary_marker: a no-op on this marker; on the CUDA path it also rebuilds the occupancy maps the extrapolation skips by
    b.run(5, 6)
    out = b.u()
    for c in range(3):
        valid = fl | np.roll(fl, -1, axis=2 - c)
        a1, a2 = [ax for ax in range(3) if ax != 2 - c]
        ring = np.zeros_like(valid)
        for d1 in (-1, 0, 1):
            for d2 in (-1, 0, 1):
                if d1 or d2:
                    ring |= np.roll(np.roll(valid, d1, axis=a1), d2, axis=a2)
        ring &= ~valid
        assert ring.sum() > 1000
        assert (out[c][ring] == np.float32(B[c])).all()          # averages of equal numbers
        assert (out[c][valid] == np.float32(B[c])).all()         # valid faces untouched
        rest = ~ring & ~valid
        assert np.array_equal(out[c][rest], garbage[c][rest])    # nothing else is written


@pytest.mark.parametrize("backend", BACKENDS)
def test_density_of_a_regular_lattice(backend):
    lo, hi = 6, 26
    pos = lattice_particles(lo, hi)
    b = backend(nparticles=pos.shape[0])
    b.particles(pos)
    # stages 6..9: clear, advect in a zero velocity field (positions unchanged, cells marked, lists rebuilt), boundary marker, density rhs
    b.put(u=[np.zeros((N, N, N), dtype=np.float32)] * 3)
    b.run(6, 10)
    m = b.marker()
    assert (m[lo:hi, lo:hi, lo:hi] == O.FLUID).all() and (m == O.FLUID).sum() == (hi - lo) ** 3
    rhs = b.residual()
    inner = rhs[lo + 1:hi - 1, lo + 1:hi - 1, lo + 1:hi - 1]
    assert np.abs(inner).max() <= 1e-4  # density exactly 8 up to fp32 rounding of 64 weights, divided by dt
    # a cell on a flat face of the block misses one layer of neighbours: density 8 - 2 (the far layer carries weight 0.25 per axis ->
    # 8 * (1 - 0.25 / 2) = 7), but an AIR neighbour clamps it back up to the rest density: rhs = 0 there as well
    face = rhs[lo + 2:hi - 2, lo + 2:hi - 2, lo]
    assert np.abs(face).max() <= 1e-4
