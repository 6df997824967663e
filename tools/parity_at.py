"""The oracle-sampled parity check of bench.py (one device step vs one CPU-oracle step from the same bits on the 27-bin neighbourhoods
of random home bins) at an arbitrary grid — e.g. 1024^3, where the default bench line skips it (two 35 GB downloads).
Usage: python tools/parity_at.py GRID [bins]"""
import json, pathlib, sys, time
ROOT = pathlib.Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
import fluidsolvergpu_b200 as fsg
import bench

G = int(sys.argv[1]) if len(sys.argv) > 1 else 256
bins = int(sys.argv[2]) if len(sys.argv) > 2 else 256
cfg = fsg.scenes.plume_config(G)
cfg.capacity = fsg.scenes.plume_count(cfg, bench.SPACING)
t0 = time.time()
with fsg.FluidSolver(cfg) as s:
    n = s.scene_plume(bench.SPACING, bench.JITTER, bench.SEED)
    s.step(3)
    rec = bench.parity_sample(s, cfg, bins)
rec["particles"] = n
rec["seconds"] = round(time.time() - t0, 1)
print(json.dumps(rec))
