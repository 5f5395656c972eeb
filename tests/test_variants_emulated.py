"""Index logic of the less obvious kernels, settled on a CPU before any GPU minute is spent: the bit-mask extrapolation is emulated word by
word and must equal the oracle's pass bit for bit; the gather P2G is emulated thread by thread (slots, lane shifts, shared-memory
combination, halos); the warp-level segmented run reduction of the scatter form is emulated lane by lane with the exact shuffle / ballot
semantics of the kernel."""
import numpy as np

from oracle import oracle as O
from tests.util import DT


M32 = 0xFFFFFFFF


def test_bit_mask_extrapolation_equals_the_reference_pass():
    """extrapolate_kernel (fluid_kernels.cu) works on 32-cell FLUID words: validity words by shifts / ORs, candidates by bit tests.  Emulated
    word for word here (ragged rows: nx = 40 -> a full word and an 8-cell word) and compared with the oracle's per-cell pass, bit for bit."""
    nx, ny, nz = 40, 24, 24
    rng = np.random.default_rng(2)
    m = np.full((nz, ny, nx), O.AIR, np.int8)
    m[rng.random(m.shape) < 0.1] = O.FLUID
    m[rng.random(m.shape) < 0.03] = O.SOLID
    m[0], m[-1], m[:, 0], m[:, -1], m[:, :, 0], m[:, :, -1] = 0, 0, 0, 0, 0, 0
    u = [rng.uniform(-5, 5, m.shape).astype(np.float32) for _ in range(3)]
    f = O.OracleFluid(nx, ny, nz, 8)
    f.grid(O.ARR_MARKER)[:] = m
    for c, a in enumerate((O.ARR_UX, O.ARR_UY, O.ARR_UZ)):
        f.grid(a)[:] = u[c]
    f.step_stages(DT, 5, 6)
    want = [f.grid(a).copy() for a in (O.ARR_UX, O.ARR_UY, O.ARR_UZ)]
    # fluid_bits_kernel
    wpr = (nx + 31) // 32
    words = np.zeros((nz, ny, wpr), np.uint64)
    for x in range(nx):
        words[:, :, x // 32] |= (m[:, :, x] == O.FLUID).astype(np.uint64) << np.uint64(x % 32)

    def fbits(xw, y, z):
        if xw < 0 or xw >= wpr or y < 0 or y >= ny or z < 0 or z >= nz:
            return 0
        return int(words[z, y, xw])

    def valid_word(c, xw, y, z):
        fw = fbits(xw, y, z)
        if c == 0:
            return (fw | (fw >> 1) | (fbits(xw + 1, y, z) << 31)) & M32
        if c == 1:
            return fw | fbits(xw, y + 1, z)
        return fw | fbits(xw, y, z + 1)

    out = [x.copy() for x in u]
    strides = (1, nx, nx * ny)
    for z in range(nz):
        for y in range(ny):
            for xw in range(wpr):
                cells = min(32, nx - xw * 32)
                cells_mask = M32 if cells == 32 else (1 << cells) - 1
                fluid = fbits(xw, y, z)
                for c in range(3):
                    nb = [[0] * 3 for _ in range(3)]
                    if c == 0:
                        for ob in (-1, 0, 1):
                            for oa in (-1, 0, 1):
                                nb[ob + 1][oa + 1] = valid_word(0, xw, y + oa, z + ob)
                    else:
                        for ob in (-1, 0, 1):
                            yy, zz = (y, z + ob) if c == 1 else (y + ob, z)
                            mm, l, r = valid_word(c, xw, yy, zz), valid_word(c, xw - 1, yy, zz), valid_word(c, xw + 1, yy, zz)
                            nb[ob + 1][0] = ((mm << 1) | (l >> 31)) & M32
                            nb[ob + 1][1] = mm
                            nb[ob + 1][2] = ((mm >> 1) | (r << 31)) & M32
                    own, anyv = nb[1][1], 0
                    for ob in range(3):
                        for oa in range(3):
                            if oa != 1 or ob != 1:
                                anyv |= nb[ob][oa]
                    todo = cells_mask & ~fluid & ~own & anyv & M32
                    sa = strides[1] if c == 0 else 1
                    sb = strides[1] if c == 2 else strides[2]
                    flat_in, flat_out = u[c].reshape(-1), out[c].reshape(-1)
                    row = (z * ny + y) * nx + xw * 32
                    k = 0
                    while todo:
                        if todo & 1:
                            numv, avg = np.float32(0), np.float32(0)
                            for ob in (-1, 0, 1):
                                for oa in (-1, 0, 1):
                                    if (oa or ob) and (nb[ob + 1][oa + 1] >> k) & 1:
                                        numv += np.float32(1)
                                        avg = np.float32(avg + flat_in[row + k + oa * sa + ob * sb])
                            flat_out[row + k] = np.float32(avg / numv)
                        todo >>= 1
                        k += 1
    assert sum(int((want[c] != u[c]).sum()) for c in range(3)) > 10000
    for c in range(3):
        assert np.array_equal(out[c], want[c])


def test_gather_p2g_index_logic_equals_a_direct_scatter():
    """p2g_gather_kernel (transfer_kernels.cu) emulated thread by thread in Python -- blocks of 32 lanes x (GW + 2) warps marching along y,
    18 face slots per cell, x-combination by lane shifts, z-combination through the shared buffer, halo lanes / warps / rows dropped -- against
    a plain per-particle scatter of the same (face, particle) pairs in float64."""
    GW, GLY, GXS = 6, 16, 30
    nx, ny, nz = 40, 24, 24
    rng = np.random.default_rng(4)
    n = 2500
    pos = np.c_[rng.uniform(1.0, nx - 1.0, n), rng.uniform(1.0, ny - 1.0, n), rng.uniform(1.0, nz - 1.0, n)]
    pos[:20, 0], pos[20:40, 1], pos[40:60, 2] = nx - 1.0, 1.0, nz - 1.0  # on the clamp planes
    rows = rng.normal(0, 2.0, (3, n, 4))
    cell = (np.minimum(pos[:, 2].astype(int), nz - 1) * ny + np.minimum(pos[:, 1].astype(int), ny - 1)) * nx + np.minimum(pos[:, 0].astype(int), nx - 1)
    order = np.argsort(cell, kind="stable")
    cell_start = np.searchsorted(cell[order], np.arange(nx * ny * nz + 1))
    sat = lambda v: min(max(v, 0.0), 1.0)
    for axis in range(3):
        off = [0.5, 0.5, 0.5]
        off[axis] = 1.0
        # direct scatter (transfer_build_linkedlist.comp:21-23 + transfer_gather_velocity.comp:23-31)
        want = np.zeros((nz, ny, nx, 2))
        for i in range(n):
            d = [int(pos[i, k] - off[k]) for k in range(3)]
            for oz in (0, 1):
                for oy in (0, 1):
                    for ox in (0, 1):
                        fc = [d[0] + ox, d[1] + oy, d[2] + oz]
                        t = [fc[k] + off[k] - pos[i, k] for k in range(3)]
                        w = sat(1 - abs(t[0])) * sat(1 - abs(t[1])) * sat(1 - abs(t[2]))
                        if w > 0:
                            want[fc[2], fc[1], fc[0]] += (w * (rows[axis, i, :3] @ t + rows[axis, i, 3]), w)
        got = np.full((nz, ny, nx, 2), np.nan)
        nf = [2 if k == axis else 3 for k in range(3)]
        for bz in range((nz + GW - 1) // GW):
            for by in range((ny + GLY - 1) // GLY):
                for bx in range((nx + GXS - 1) // GXS):
                    x0, y0, z0 = bx * GXS, by * GLY, bz * GW
                    acc = np.zeros((GW + 2, 32, nf[1], nf[0], nf[2], 2))  # [warp][lane][fy][fx][fz]
                    y_end = min(y0 + GLY, ny)
                    for y in range(y0 - 1, y0 + GLY + 1):
                        for wz in range(GW + 2):
                            for lane in range(32):
                                cx, cz = x0 - 1 + lane, z0 - 1 + wz
                                if not (0 <= cx < nx and 0 <= cz < nz and 0 <= y < ny):
                                    continue
                                ci = (cz * ny + y) * nx + cx
                                for j in range(cell_start[ci], cell_start[ci + 1]):
                                    i = order[j]
                                    cc = (cx, y, cz)
                                    t = [[(cc[k] + f - 1) + off[k] - pos[i, k] for f in range(nf[k])] for k in range(3)]
                                    w = [[sat(1 - abs(v)) for v in t[k]] for k in range(3)]
                                    for fy in range(nf[1]):
                                        for fx in range(nf[0]):
                                            for fz in range(nf[2]):
                                                ww = w[0][fx] * w[1][fy] * w[2][fz]
                                                vv = rows[axis, i, 0] * t[0][fx] + rows[axis, i, 1] * t[1][fy] + rows[axis, i, 2] * t[2][fz] + rows[axis, i, 3]
                                                acc[wz, lane, fy, fx, fz] += (ww * vv, ww)
                        yr = y - 1
                        if y0 <= yr < y_end:
                            q = np.zeros((3, GW + 2, 32, 2))
                            for wz in range(GW + 2):
                                for lane in range(32):
                                    for fz in range(nf[2]):
                                        a = acc[wz, :, 0, :, fz]  # [lane][fx]
                                        if axis == 0:
                                            s = a[lane, 1] + (a[lane + 1, 0] if lane < 31 else a[lane, 0])
                                        else:
                                            s = (a[lane - 1, 2] if lane > 0 else a[lane, 2]) + a[lane, 1] + (a[lane + 1, 0] if lane < 31 else a[lane, 0])
                                        q[fz, wz, lane] = s
                            for wz in range(1, GW + 1):
                                for lane in range(1, GXS + 1):
                                    cx, cz = x0 - 1 + lane, z0 - 1 + wz
                                    if cx >= nx or cz >= nz:
                                        continue
                                    s = q[1, wz, lane] + q[0, wz + 1, lane] if axis == 2 else q[2, wz - 1, lane] + q[1, wz, lane] + q[0, wz + 1, lane]
                                    assert np.isnan(got[cz, yr, cx, 0])  # every face is stored by exactly one thread
                                    got[cz, yr, cx] = s
                        acc[:, :, :-1] = acc[:, :, 1:].copy()
                        acc[:, :, -1] = 0
        assert not np.isnan(got).any()
        assert np.abs(got - want).max() <= 1e-9 * max(1.0, np.abs(want).max())


def emulate_segmented_run_sum(keys, v):
    """segmented_run_sum<NV> of transfer_kernels.cu for one 32-lane warp (NV = 1): runs of equal keys cut into groups of four lanes."""
    lane = np.arange(32)
    v = v.astype(np.float64).copy()
    prev = np.r_[keys[0], keys[:-1]]                      # __shfl_up_sync(key, 1); lane 0 keeps its own
    start = (lane == 0) | (prev != keys)
    starts = sum(1 << i for i in range(32) if start[i])   # __ballot_sync
    if starts == 0xFFFFFFFF:
        return np.ones(32, bool), v
    pos, end = np.empty(32, int), np.empty(32, int)
    for l in range(32):
        upto = starts & (0xFFFFFFFF if l == 31 else ((2 << l) - 1))
        pos[l] = l - (upto.bit_length() - 1)              # 31 - __clz(upto)
        above = 0 if l == 31 else starts & ~(((2 << l) - 1) & 0xFFFFFFFF) & 0xFFFFFFFF
        end[l] = ((above & -above).bit_length() - 1) if above else 32   # __ffs(above) - 1
    for o in (1, 2):
        t = np.array([v[l + o] if l + o < 32 else v[l] for l in range(32)])  # __shfl_down_sync: out-of-range lanes read themselves
        v = np.where(((pos & 3) + o < 4) & (lane + o < end), v + t, v)
    return (pos & 3) == 0, v


def test_segmented_run_reduction_sums_every_group_of_four_into_its_first_lane():
    rng = np.random.default_rng(0)
    for trial in range(500):
        n_runs = int(rng.integers(1, 33))
        cuts = np.sort(rng.choice(np.arange(1, 32), size=n_runs - 1, replace=False)) if n_runs > 1 else np.array([], int)
        keys, start = np.zeros(32, int), 0
        for k, c in enumerate(list(cuts) + [32]):
            keys[start:c] = int(rng.integers(0, 5)) if trial % 2 else k  # odd trials: adjacent runs may share a key and merge
            start = c
        v = rng.random(32)
        head, out = emulate_segmented_run_sum(keys, v)
        l = 0
        while l < 32:
            r = l
            while r + 1 < 32 and keys[r + 1] == keys[l]:
                r += 1
            for g in range(l, r + 1, 4):  # groups of four inside the run [l, r]
                ge = min(g + 3, r)
                assert head[g] and not head[g + 1:ge + 1].any()
                assert abs(out[g] - v[g:ge + 1].sum()) < 1e-12
            l = r + 1
        assert abs(out[head].sum() - v.sum()) < 1e-9  # every contribution is issued exactly once


def emulate_matched_group_sum(keys, v):
    """matched_group_sum<NV> of transfer_kernels.cu for one 32-lane warp (NV = 1): ALL lanes with the same key form a peer group
    (match.any), cut into groups of four by rank; tree (v0 + v1) + (v2 + v3), the lowest lane of a group issues."""
    v = v.astype(np.float64).copy()
    peers = [sum(1 << j for j in range(32) if keys[j] == keys[l]) for l in range(32)]          # __match_any_sync
    if all(peers[l] == (1 << l) for l in range(32)):                                          # __all_sync
        return np.ones(32, bool), v
    rank = np.array([bin(peers[l] & ((1 << l) - 1)).count("1") for l in range(32)])
    n1 = np.empty(32, int)
    for l in range(32):
        above = 0 if l == 31 else peers[l] & ~((2 << l) - 1) & 0xFFFFFFFF
        n1[l] = ((above & -above).bit_length() - 1) if above else -1                           # __ffs(above) - 1
    n2 = np.array([n1[n1[l]] if n1[l] >= 0 else n1[l] for l in range(32)])                     # __shfl_sync(n1, n1 >= 0 ? n1 : lane)
    take1, take2 = ((rank & 1) == 0) & (n1 >= 0), ((rank & 3) == 0) & (n2 >= 0)
    t = np.array([v[n1[l]] if n1[l] >= 0 else v[l] for l in range(32)])
    v = np.where(take1, v + t, v)
    t = np.array([v[n2[l]] if n2[l] >= 0 else v[l] for l in range(32)])
    v = np.where(take2, v + t, v)
    return (rank & 3) == 0, v


def test_matched_group_reduction_sums_every_four_peers_into_their_lowest_lane():
    rng = np.random.default_rng(1)
    for trial in range(600):
        if trial % 3 == 0:
            keys = rng.integers(0, int(rng.integers(1, 40)), size=32)          # peers anywhere in the warp
        elif trial % 3 == 1:
            keys = np.sort(rng.integers(0, 6, size=32)) * 2 + rng.integers(0, 2, size=32)  # cell-sorted warp, two dual cells per primal cell
        else:
            keys = -1 - np.arange(32)                                          # invalid lanes: unique keys, nobody has a peer
            keys[: int(rng.integers(0, 33))] = int(rng.integers(0, 3))
        v = rng.random(32)
        head, out = emulate_matched_group_sum(keys, v)
        for key in np.unique(keys):
            lanes = np.flatnonzero(keys == key)
            for g in range(0, len(lanes), 4):
                grp = lanes[g:g + 4]
                assert head[grp[0]] and not head[grp[1:]].any()
                assert abs(out[grp[0]] - v[grp].sum()) < 1e-12
        assert abs(out[head].sum() - v.sum()) < 1e-9  # every contribution is issued exactly once
        # adjacent runs are a special case of peer groups: on run-structured keys without repeats both reductions issue the same sums
        if trial % 3 == 2:
            head_a, out_a = emulate_segmented_run_sum(keys, v)
            assert (head_a == head).all() and np.abs(out_a[head] - out[head]).max() < 1e-12


def test_peer_groups_issue_far_fewer_reductions_than_adjacent_runs():
    """The design argument for matched_group_sum, on the oracle's dam break (32^3, 40 k particles) in the array order the CUDA path keeps
    (stable sort by primal cell every 8 steps): peer groups of four save >= 45 % of the P2G reductions at every step, adjacent runs <= 30 %."""
    import os
    import sys

    sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tools"))
    import red_stats as R

    from tests import util

    f = util.oracle_from_scene("dam_small")
    f.set_rebin_frequency(0)
    dims = O.load_scene(util.scene_path("dam_small"))["fluid"]["grid_dimension"]
    nx, ny = dims["x"], dims["y"]
    order = np.arange(f.num_particles)
    for step in range(10):
        p = f.particles()[:, :3].copy()
        if step % 8 == 0:
            order = R.sort_by_primal_cell(p, order, nx, ny)
        v, d = R.reductions_per_particle(p[order], nx, ny), R.reductions_per_particle(p[order], nx, ny, comps=(3,))
        assert v["none"] == 24.0 and d["none"] == 8.0
        assert v["match4"] <= 0.55 * 24.0 and v["adjacent4"] >= 0.70 * 24.0, (step, v)
        assert d["match4"] <= 0.70 * 8.0 and d["match4"] <= 0.80 * d["adjacent4"], (step, d)
        f.step(DT)
