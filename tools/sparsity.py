"""How sparse is the FLUID set the pressure solver sees?  (decides the skip granularity of the PCG kernels)

    python tools/sparsity.py [scene] [checkpoint steps ...]
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import blub_b200  # noqa: E402
from blub_b200 import fluid as F  # noqa: E402

scene = sys.argv[1] if len(sys.argv) > 1 else "dam_256"
checkpoints = [int(a) for a in sys.argv[2:]] or [3, 60, 110]
f = blub_b200.HybridFluid.from_scene(os.path.join(ROOT, "tests", "golden", "scenes", scene + ".json"))
done = 0
print("steps  fluid%  tile(128x8x4)%  row(128x1x4)%  col(4x1x4)%  quad(4x1x1)%  sector(8x1x1)%  tile(32x8x4)%  fluid+nbr%")
for cp in checkpoints:
    while done < cp:
        f.step(F.DT_120HZ)
        done += 1
    f.synchronize()
    m = f.download_grid(F.TAP_MARKER) == 1  # [z, y, x]
    nz, ny, nx = m.shape
    fl = m.mean()
    t = m.reshape(nz // 4, 4, ny // 8, 8, nx // 128, 128).any(axis=(1, 3, 5)).mean()
    row = m.reshape(nz // 4, 4, ny, nx // 128, 128).any(axis=(1, 4)).mean()
    col = m.reshape(nz // 4, 4, ny, nx // 4, 4).any(axis=(1, 4)).mean()
    quad = m.reshape(nz, ny, nx // 4, 4).any(axis=3).mean()
    sec = m.reshape(nz, ny, nx // 8, 8).any(axis=3).mean()
    t32 = m.reshape(nz // 4, 4, ny // 8, 8, nx // 32, 32).any(axis=(1, 3, 5)).mean()
    nb = m.copy()
    nb[1:] |= m[:-1]; nb[:-1] |= m[1:]; nb[:, 1:] |= m[:, :-1]; nb[:, :-1] |= m[:, 1:]; nb[:, :, 1:] |= m[:, :, :-1]; nb[:, :, :-1] |= m[:, :, 1:]
    print(f"{cp:5d} {100*fl:7.2f} {100*t:12.2f} {100*row:14.2f} {100*col:12.2f} {100*quad:12.2f} {100*sec:13.2f} {100*t32:14.2f} {100*nb.mean():10.2f}")
