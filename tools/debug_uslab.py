import sys, pathlib
sys.path.insert(0, str(pathlib.Path(__file__).resolve().parents[1]))
import numpy as np
import fluidsolvergpu_b200 as fsg
state = fsg.scenes.random_unidyn_scene(6000, 11, boundary_frac=0.1)
state["vel"][state["boundary"] == 0, 0] += np.float32(4.0)
n = state["pos"].shape[0]
cfg = fsg.FluidSolver.unidyn_config(capacity=n)
world = 2
cuts = fsg.slab_cuts(fsg.slab.layer_hist_from_positions(cfg, state["pos"]), world)
print("cuts", cuts, "n", n)
G = cfg.grid
with fsg.SlabGroup(cfg, world, cuts, capacity=2 * n + 64) as g:
    g.upload(state)
    for k in range(6):
        for r, sl in enumerate(g.slabs):
            d = sl.download_slots()
            cell = d["cell"]
            live = cell < G ** 3
            ix = cell[live] // (G * G)
            print("step", k, "rank", r, "slots", len(cell), "live", int(live.sum()), "layers", (ix.min(), ix.max()) if ix.size else None,
                  "parked", int((cell == G ** 3).sum()), "dead", int((cell > G ** 3).sum()))
        g.step(1)
        try:
            print("check", g.check())
        except Exception as e:
            print("check failed:", e)
            for r, sl in enumerate(g.slabs):
                d = sl.download_slots()
                cell = d["cell"]; live = cell < G ** 3
                ix = cell[live] // (G * G)
                print(" rank", r, "layers after", (ix.min(), ix.max()), np.bincount(ix, minlength=G))
                bad = live & ((cell // (G * G) < cuts[r][0] - 1) | (cell // (G * G) > cuts[r][1]))
                print(" bad", int(bad.sum()), d["pos"][bad][:5], d["vel"][bad][:5], d["boundary"][bad][:5], d["index"][bad][:5])
            break
