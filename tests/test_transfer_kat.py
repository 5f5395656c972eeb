"""Known-answer tests of the particle <-> grid transfers: closed-form results that any correct implementation of the reference's
formulas must reproduce, evaluated on the oracle (CPU) and on the CUDA path (GPU, same assertions).

* P2G (transfer_gather_velocity.comp:63-127): every particle contributes  w * (row_c . (q - p) + v_c)  to face q.  For a velocity
  field  v(x) = S x + b  with SYMMETRIC S and rows set as the G2P pass would set them (Jacobian columns stored as rows)
  each contribution equals  w * v_c(q):  the normalised sum is v_c(q) whatever the particle positions are.
* G2P (advect_particles.comp:73-126, 184-188): trilinear interpolation is exact on linear functions, so a grid holding a linear
  field (ANY matrix A) at its staggered face centres gives every particle  v(x0)  and the columns of A as its three rows; in a
  constant field the RK4 position update is  x0 + dt * b  exactly.
"""
import numpy as np
import pytest

from oracle import oracle as O
from tests.util import DT

N = 32
S_SYM = np.array([[1.5, -0.75, 0.5], [-0.75, -2.0, 1.25], [0.5, 1.25, 0.5]], dtype=np.float64)   # 1/s
A_ANY = np.array([[0.5, -1.0, 2.0], [1.5, 0.25, -0.5], [-2.0, 1.0, -0.75]], dtype=np.float64)   # 1/s
B_VEC = np.array([4.0, -6.0, 2.5], dtype=np.float64)                                              # cells/s


def face_centres(c):
    """Position of velocity sample u_c[z, y, x]: +1 along c, +0.5 along the other axes (hybrid_fluid.glsl staggering)."""
    z, y, x = np.meshgrid(np.arange(N), np.arange(N), np.arange(N), indexing="ij")
    q = np.stack([x, y, z], axis=-1).astype(np.float64) + 0.5
    q[..., c] += 0.5
    return q


def linear_field(M, pts):
    return pts @ M.T + B_VEC


def block_particles(fluid):
    fluid.add_fluid_cube([10.0, 10.0, 10.0], [22.0, 22.0, 22.0])
    fluid.set_gravity_grid([0.0, 0.0, 0.0])


def affine_rows(pos):
    """What a G2P pass over the field S x + b leaves in the three row streams (row_c = (dv_x/dc, dv_y/dc, dv_z/dc, v_c))."""
    v = linear_field(S_SYM, pos[:, :3].astype(np.float64))
    return [np.c_[np.tile(S_SYM[:, c], (pos.shape[0], 1)), v[:, c]].astype(np.float32) for c in range(3)]


def check_p2g(marker, u):
    """u[c] must equal v_c at every face between two non-SOLID cells of which at least one is FLUID."""
    checked = 0
    for c in range(3):
        ax = 2 - c
        nb = np.roll(marker, -1, axis=ax)
        sel = ((marker == O.FLUID) | (nb == O.FLUID)) & (marker != O.SOLID) & (nb != O.SOLID)
        idx = [slice(None)] * 3
        idx[ax] = slice(0, N - 1)
        sel[tuple(idx)] &= True
        idx[ax] = slice(N - 1, N)
        sel[tuple(idx)] = False  # no +c neighbour
        want = linear_field(S_SYM, face_centres(c))[..., c]
        err = np.abs(u[c][sel] - want[sel]).max()
        assert err <= 2e-4 * np.abs(want[sel]).max() + 1e-4, (c, err)
        checked += int(sel.sum())
    assert checked > 3 * 12 ** 3


def test_oracle_p2g_reproduces_a_symmetric_affine_field():
    f = O.OracleFluid(N, N, N, 20000)
    block_particles(f)
    pos = f.particles().copy()
    f.set_particles(pos, *affine_rows(pos))
    f.step_stages(DT, 0, 1)
    check_p2g(f.grid(O.ARR_MARKER), [f.grid(a) for a in (O.ARR_UX, O.ARR_UY, O.ARR_UZ)])


def g2p_setup(seed=3, count=4000):
    rng = np.random.default_rng(seed)
    pos = np.c_[rng.uniform(8.0, 24.0, (count, 3)), np.zeros(count)].astype(np.float32)
    marker = np.full((N, N, N), O.AIR, dtype=np.int8)
    marker[0], marker[-1], marker[:, 0], marker[:, -1], marker[:, :, 0], marker[:, :, -1] = 0, 0, 0, 0, 0, 0
    cells = np.floor(pos[:, :3]).astype(int)
    marker[cells[:, 2], cells[:, 1], cells[:, 0]] = O.FLUID
    return pos, marker


def check_g2p_linear(pos0, rows):
    v0 = linear_field(A_ANY, pos0[:, :3].astype(np.float64))
    scale = np.abs(v0).max()
    for c in range(3):  # row_c = (column c of A, v_c(x0))
        assert np.abs(rows[c][:, 3] - v0[:, c]).max() <= 1e-5 * scale + 1e-4, c
        assert np.abs(rows[c][:, :3] - A_ANY[:, c]).max() <= 2e-3, (c, np.abs(rows[c][:, :3] - A_ANY[:, c]).max())


def test_oracle_g2p_is_exact_on_linear_fields():
    pos, marker = g2p_setup()
    f = O.OracleFluid(N, N, N, pos.shape[0])
    f.set_gravity_grid([0.0, 0.0, 0.0])
    f.set_particles(pos)
    f.grid(O.ARR_MARKER)[:] = marker
    for c, a in enumerate((O.ARR_UX, O.ARR_UY, O.ARR_UZ)):
        f.grid(a)[:] = linear_field(A_ANY, face_centres(c))[..., c].astype(np.float32)
    f.step_stages(1e-3, 6, 8)  # transfer_clear + advect
    check_g2p_linear(pos, [f.particles(a) for a in (O.ARR_ROWX, O.ARR_ROWY, O.ARR_ROWZ)])


def test_oracle_advection_in_a_constant_field():
    pos, marker = g2p_setup(seed=4)
    f = O.OracleFluid(N, N, N, pos.shape[0])
    f.set_gravity_grid([0.0, 0.0, 0.0])
    f.set_particles(pos)
    f.grid(O.ARR_MARKER)[:] = marker
    for c, a in enumerate((O.ARR_UX, O.ARR_UY, O.ARR_UZ)):
        f.grid(a)[:] = np.float32(B_VEC[c])
    f.step_stages(DT, 6, 8)
    want = pos[:, :3].astype(np.float64) + DT * B_VEC
    assert np.abs(f.particles()[:, :3] - want).max() <= 4e-6 * 24
    for c, a in enumerate((O.ARR_ROWX, O.ARR_ROWY, O.ARR_ROWZ)):
        r = f.particles(a)
        assert np.abs(r[:, :3]).max() == 0.0 and np.abs(r[:, 3] - B_VEC[c]).max() <= 1e-6 * 6


# ------------------------------------------------------------------------------------------------ the same on the CUDA path
@pytest.mark.gpu
def test_cuda_p2g_reproduces_a_symmetric_affine_field():
    import blub_b200
    from blub_b200 import fluid as F

    gpu = blub_b200.HybridFluid(N, N, N, 20000)
    block_particles(gpu)
    pos = gpu.download_particles().copy()
    gpu.set_particles(pos, *affine_rows(pos))
    gpu.step_stages(DT, 0, 1)
    check_p2g(gpu.download_grid(F.TAP_MARKER), [gpu.download_grid(t) for t in (F.TAP_UX, F.TAP_UY, F.TAP_UZ)])


@pytest.mark.gpu
def test_cuda_g2p_is_exact_on_linear_fields_and_advects_exactly_in_a_constant_one():
    import blub_b200
    from blub_b200 import fluid as F

    for linear in (True, False):
        pos, marker = g2p_setup(seed=3 if linear else 4)
        gpu = blub_b200.HybridFluid(N, N, N, pos.shape[0])
        gpu.set_gravity_grid([0.0, 0.0, 0.0])
        gpu.set_particles(pos)
        gpu.upload_grid(F.TAP_MARKER, marker)
        for c, t in enumerate((F.TAP_UX, F.TAP_UY, F.TAP_UZ)):
            field = linear_field(A_ANY, face_centres(c))[..., c] if linear else np.full((N, N, N), B_VEC[c])
            gpu.upload_grid(t, field.astype(np.float32))
        gpu.step_stages(1e-3 if linear else DT, 6, 8)
        rows = [gpu.download_particles(t) for t in (F.TAP_VX, F.TAP_VY, F.TAP_VZ)]
        if linear:
            check_g2p_linear(pos, rows)
        else:
            want = pos[:, :3].astype(np.float64) + DT * B_VEC
            assert np.abs(gpu.download_particles()[:, :3] - want).max() <= 4e-6 * 24
            for c in range(3):
                assert np.abs(rows[c][:, :3]).max() <= 1e-5 and np.abs(rows[c][:, 3] - B_VEC[c]).max() <= 1e-5
