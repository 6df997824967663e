"""Ad-hoc timing of the unidyn path on the 128^3 plume (pure fluid): development aid for ncu launch lists."""
import sys, pathlib, json
ROOT = pathlib.Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
import numpy as np
import torch
import fluidsolvergpu_b200 as fsg
from fluidsolvergpu_b200 import scenes

G = int(sys.argv[1]) if len(sys.argv) > 1 else 128
cfg = scenes.plume_config(G)
s3 = scenes.plume_scene(cfg, 0.05, 0.005, 20261018)
n3 = s3["pos"].shape[0]
s3["solid"], s3["fluid"] = np.zeros(n3, np.float32), np.ones(n3, np.float32)
ucfg = fsg.FluidSolver.unidyn_config(capacity=n3, grid=G, origin=cfg.origin, unidyn_open_box=1)
with fsg.FluidSolver(ucfg) as s:
    s.upload(s3)
    s.step(5)
    stream = torch.cuda.ExternalStream(s.stream())
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    s.step(10, sync=False)
    e1.record(stream)
    s.sync()
    torch.cuda.synchronize()
    print(json.dumps(dict(G=G, n=n3, ms_per_step=e0.elapsed_time(e1) / 10)))
