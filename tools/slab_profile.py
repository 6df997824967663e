"""W slabs of the G^3 plume in ONE process (peer-memory protocol with plain pointers; sorted-ghost pipeline unless `classic`), a few
steps: run under ncu --metrics gpu__time_duration.sum to see what the slab phases cost kernel by kernel.
Usage: python tools/slab_profile.py [G] [W] [steps] [classic]"""
import sys, pathlib
sys.path.insert(0, str(pathlib.Path(__file__).resolve().parents[1]))
import fluidsolvergpu_b200 as fsg

G = int(sys.argv[1]) if len(sys.argv) > 1 else 256
W = int(sys.argv[2]) if len(sys.argv) > 2 else 2
steps = int(sys.argv[3]) if len(sys.argv) > 3 else 4
classic = len(sys.argv) > 4 and sys.argv[4] == "classic"
cfg = fsg.scenes.plume_config(G)
hist = fsg.slab.plume_layer_hist(cfg, 0.05)
cuts = fsg.slab_cuts(hist, W)
owned = [int(hist[a:b].sum()) for a, b in cuts]
cap_m, cap_g = fsg.slab.message_caps(hist, cuts)
cap = int(max(owned) * 1.05) + 3 * int(hist.max()) + 65536 if classic else int(max(owned) * 1.03) + 2 * cap_g + 65536
with fsg.SlabGroup(cfg, W, cuts, capacity=cap, cap_m=cap_m, cap_g=cap_g, peer=True, classic=classic) as g:
    g.scene_plume(0.05, 0.005, 20261018)
    g.step(steps)
    print("ok mode", g.slabs[0].mode, g.check())
