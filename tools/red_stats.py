"""How many reductions (RED lane-ops) does the P2G / density scatter issue per particle under each warp aggregation scheme?

Pure NumPy on particle positions in ARRAY ORDER (no GPU, no oracle import: positions come from a file or from the caller --
tests/test_variants_emulated.py feeds positions of the CPU oracle's dam break, re-sorted by primal cell every 8 steps as the CUDA path does).

    python tools/red_stats.py snapshots.npz NX NY        # arrays of shape [n, 3] (grid units), one line per array

Schemes: none = one reduction per (particle, face); adjacent4 = runs of adjacent lanes with the same dual cell, cut into groups of four
(segmented_run_sum); match4 / match8 / matchall = all lanes of the warp with the same dual cell (matched_group_sum), groups of 4 / 8 / any size.
"""
import sys

import numpy as np


def dual_keys(p, c, nx, ny):
    """dual cell of component c (0..2: velocity components, transfer_build_linkedlist.comp:21-23; 3: density, offset 0.5 everywhere)"""
    off = np.full(3, 0.5, np.float32)
    if c < 3:
        off[c] = 1.0
    d = (p.astype(np.float32) - off).astype(np.int32)
    return d[:, 0] + nx * (d[:, 1] + ny * d[:, 2])


def reductions_per_particle(p, nx, ny, comps=(0, 1, 2)):
    """p: [n, 3] positions in array order.  Returns {scheme: reductions per particle} summed over `comps` (8 faces per issuing lane)."""
    m = len(p) // 32 * 32
    idx = np.arange(32)[None, :]
    tot = dict(none=0, adjacent4=0, match4=0, match8=0, matchall=0)
    for c in comps:
        k = dual_keys(p[:m], c, nx, ny).reshape(-1, 32)
        start = np.ones_like(k, bool)
        start[:, 1:] = k[:, 1:] != k[:, :-1]
        pos = idx - np.maximum.accumulate(np.where(start, idx, 0), axis=1)           # position inside the adjacent run
        ks = np.sort(k, axis=1)                                                      # rank among ALL equal keys of the warp
        st = np.ones_like(ks, bool)
        st[:, 1:] = ks[:, 1:] != ks[:, :-1]
        r = idx - np.maximum.accumulate(np.where(st, idx, 0), axis=1)
        tot["none"] += k.size
        tot["adjacent4"] += int(((pos & 3) == 0).sum())
        tot["match4"] += int(((r & 3) == 0).sum())
        tot["match8"] += int(((r & 7) == 0).sum())
        tot["matchall"] += int(st.sum())
    return {a: 8.0 * b / m for a, b in tot.items()}


def sort_by_primal_cell(p, order, nx, ny):
    """the stable counting sort of the binning stage: x-fastest cell order, ascending previous index inside a cell"""
    ci = p.astype(np.int32)
    cell = ci[:, 0] + nx * (ci[:, 1] + ny * ci[:, 2])
    return order[np.argsort(cell[order], kind="stable")]


if __name__ == "__main__":
    data, nx, ny = np.load(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
    for key in data.files:
        v, d = reductions_per_particle(data[key], nx, ny), reductions_per_particle(data[key], nx, ny, comps=(3,))
        print(key, "P2G", " ".join(f"{a} {b:.2f}" for a, b in v.items()), "| density", " ".join(f"{a} {b:.2f}" for a, b in d.items()))
