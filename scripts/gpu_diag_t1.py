# Diagnostic: one step of the scalar tile kernel against the direct kernel, which populations / cells differ
import sys
import numpy as np
sys.path.insert(0, "tests")
sys.path.insert(0, ".")
from common import native_run, tile_case

for lattice, shape, walls in [("D3Q19", (2, 64, 16), False), ("D3Q19", (3, 4, 512), False), ("D3Q27", (2, 64, 16), False), ("D3Q19", (4, 8, 128), True)]:
    for policy in ("FP32FP32",):
        g = tile_case(lattice, shape, 1, 13, walls, policy)
        ref, _, _ = native_run(g, cells_per_thread=1)
        for v in (501, 502):
            f, _, _ = native_run(g, cells_per_thread=v)
            bad = f != ref
            print(lattice, shape, walls, policy, v, "differ:", int(bad.sum()), "of", f.size)
            if bad.any():
                print("  per population:", bad.reshape(bad.shape[0], -1).sum(1).tolist())
                l = int(np.argmax(bad.reshape(bad.shape[0], -1).sum(1)))
                idx = np.argwhere(bad[l])
                print("  population", l, "first bad cells:", idx[:6].tolist(), "last:", idx[-3:].tolist())
                print("  per x:", bad[l].sum((1, 2)).tolist()[:8], "per y:", bad[l].sum((0, 2)).tolist()[:16], "per z:", bad[l].sum((0, 1)).tolist()[:16])
                i = tuple(idx[0])
                print("  got", f[l][i], "want", ref[l][i], "init there", g["f_init"][l][i])
                # is the wrong value some other element of the input?
                hit = np.argwhere(g["f_init"] == f[l][i])
                print("  the value got is f_init at", hit[:4].tolist())
