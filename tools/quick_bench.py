"""Ad-hoc resident-state timing of fsg_step on the plume scene (development aid; bench.py is the contract)."""
import sys, time, pathlib, json
ROOT = pathlib.Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
import torch
import fluidsolvergpu_b200 as fsg

for G in [int(a) for a in sys.argv[1:]] or [64, 128, 256]:
    cfg = fsg.scenes.plume_config(G)
    cfg.capacity = fsg.scenes.plume_count(cfg)
    with fsg.FluidSolver(cfg) as s:
        n = s.scene_plume()
        s.step(2)
        st = s.pair_stats_one_step()
        stream = torch.cuda.ExternalStream(s.stream())
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        K = 10
        s.set_profiling(True)
        with torch.cuda.stream(stream):
            e0.record(stream)
            s.step(K, sync=False)
            e1.record(stream)
        s.sync()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / K
        ph = s.phase_ms()
        ph = {k: round(v / max(1, ph["steps"]), 4) for k, v in ph.items() if k != "steps"}
        print(json.dumps(dict(G=G, n=n, ms_per_step=ms, particle_steps_per_s=n / ms * 1e3, cell_updates_per_s=G**3 / ms * 1e3,
                              pairs_tested=st["pairs_tested"], pairs_in_range=st["pairs_in_range"], occupied=st["occupied_bins"],
                              hbm_frac=272 * n / (ms * 1e-3) / 6544.7e9, phases=ph)))
