"""Two slabs of the 256^3 plume in ONE process (peer-memory protocol with plain pointers), a few steps: run under
ncu --metrics gpu__time_duration.sum to see what the slab phases cost kernel by kernel."""
import sys, pathlib
sys.path.insert(0, str(pathlib.Path(__file__).resolve().parents[1]))
import fluidsolvergpu_b200 as fsg

G = int(sys.argv[1]) if len(sys.argv) > 1 else 256
cfg = fsg.scenes.plume_config(G)
hist = fsg.slab.plume_layer_hist(cfg, 0.05)
cuts = fsg.slab_cuts(hist, 2)
owned = [int(hist[a:b].sum()) for a, b in cuts]
cap = int(max(owned) * 1.05) + 3 * int(hist.max()) + 65536
cap_m, cap_g = fsg.slab.message_caps(hist, cuts)
with fsg.SlabGroup(cfg, 2, cuts, capacity=cap, cap_m=cap_m, cap_g=cap_g, peer=True) as g:
    g.scene_plume(0.05, 0.005, 20261018)
    g.step(4)
    print("ok", g.check())
