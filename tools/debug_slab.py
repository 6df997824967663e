import sys, pathlib
ROOT = pathlib.Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / "tests"))
import numpy as np
import fluidsolvergpu_b200 as fsg
from test_parity_gpu import _slab_scene
for fast in (True, False):
  for world in (2, 3, 5):
    cfg, state = _slab_scene(fsg, fast)
    n = state["pos"].shape[0]
    cuts = fsg.slab_cuts(fsg.slab.layer_hist_from_positions(cfg, state["pos"]), world)
    cfg.capacity = n
    G = cfg.grid
    with fsg.SlabGroup(cfg, world, cuts, capacity=2 * n + 64) as g, fsg.FluidSolver(cfg) as s:
        g.upload(state)
        for k in range(8):
            cur = fsg.by_index(g.download())
            s.upload({f: cur[f] for f in cur if f != "cell"})
            g.step(1); s.step(1)
            a, b = fsg.by_index(g.download()), fsg.by_index(s.download())
            d = np.abs(a["dens"].astype(np.float64) - b["dens"])
            bad = np.flatnonzero(d > 1e-2)
            if bad.size:
                print("fast", fast, "world", world, "cuts", cuts, "step", k, "bad", bad.size, "counts", [sl.last_counts for sl in g.slabs])
                for i in bad[:6]:
                    p0 = cur["pos"][i]; c0 = cur["cell"][i]
                    print("  idx", i, "ddens", d[i], "pos before", p0, "cell before", c0, "ix,iy,iz", c0 // (G*G), (c0 // G) % G, c0 % G, "vel", cur["vel"][i], "bnd", cur["boundary"][i])
                if bad.size >= 2:
                    i, j = bad[0], bad[1]
                    print("  dist(bad0,bad1)", np.linalg.norm(cur["pos"][i].astype(np.float64) - cur["pos"][j]))
                break
        else:
            print("fast", fast, "world", world, "ok")
