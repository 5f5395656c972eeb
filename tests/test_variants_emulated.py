"""The opt-in kernel variants (BLUB_EXTRAPOLATE=bytes, BLUB_SCATTER=aggregate) have not run on a GPU yet.  What CAN be settled
on a CPU is that their algorithms are right: the byte-mask extrapolation is emulated in NumPy and must equal the oracle's pass bit for
bit; the warp-level segmented run reduction is emulated lane by lane with the exact shuffle / ballot semantics of the kernel."""
import numpy as np

from oracle import oracle as O
from tests.util import DT


def test_byte_mask_extrapolation_equals_the_reference_pass():
    nx, ny, nz = 32, 24, 24
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
    # face_valid_kernel
    fl = m == O.FLUID
    valid = np.zeros(m.shape, np.uint8)
    for c in range(3):
        nb = np.zeros_like(fl)
        src, dst = [slice(None)] * 3, [slice(None)] * 3
        src[2 - c], dst[2 - c] = slice(1, None), slice(0, -1)
        nb[tuple(dst)] = fl[tuple(src)]
        valid |= (fl | nb).astype(np.uint8) << c
    # extrapolate_bytes_kernel, reading the INPUT field only (the pass never reads a face it writes)
    out = [x.copy() for x in u]
    zz, yy, xx = np.nonzero((valid & 7) != 7)
    for z, y, x in zip(zz, yy, xx):
        for c in range(3):
            if valid[z, y, x] & (1 << c):
                continue
            a, b = (1 if c == 0 else 0), (1 if c == 2 else 2)
            numv, avg = np.float32(0), np.float32(0)
            for ob in (-1, 0, 1):
                for oa in (-1, 0, 1):
                    if oa == 0 and ob == 0:
                        continue
                    h = [x, y, z]
                    h[a] += oa
                    h[b] += ob
                    if min(h) < 0 or h[0] >= nx or h[1] >= ny or h[2] >= nz:
                        continue
                    if valid[h[2], h[1], h[0]] & (1 << c):
                        numv += np.float32(1)
                        avg = np.float32(avg + u[c][h[2], h[1], h[0]])
            if numv > 0:
                out[c][z, y, x] = np.float32(avg / numv)
    assert sum(int((want[c] != u[c]).sum()) for c in range(3)) > 10000
    for c in range(3):
        assert np.array_equal(out[c], want[c])


def emulate_segmented_run_sum(keys, v):
    """segmented_run_sum<NV> of fluid_kernels.cu for one 32-lane warp (NV = 1)."""
    lane = np.arange(32)
    v = v.astype(np.float64).copy()
    prev = np.r_[keys[0], keys[:-1]]                      # __shfl_up_sync(key, 1); lane 0 keeps its own
    head = (lane == 0) | (prev != keys)
    heads = sum(1 << i for i in range(32) if head[i])     # __ballot_sync
    end = np.empty(32, int)
    for l in range(32):
        above = 0 if l == 31 else heads & ~(((2 << l) - 1) & 0xFFFFFFFF) & 0xFFFFFFFF
        end[l] = ((above & -above).bit_length() - 1) if above else 32   # __ffs(above) - 1
    o = 1
    while o < 32:
        t = np.array([v[l + o] if l + o < 32 else v[l] for l in range(32)])  # __shfl_down_sync: out-of-range lanes read themselves
        v = np.where(lane + o < end, v + t, v)
        o <<= 1
    return head, v


def test_segmented_run_reduction_sums_every_run_into_its_first_lane():
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
            assert head[l] and not head[l + 1:r + 1].any()
            assert abs(out[l] - v[l:r + 1].sum()) < 1e-12
            l = r + 1
