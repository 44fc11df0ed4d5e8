"""Debug aid: which rays of tests/test_dvr_gpu.py::test_warp_per_ray_voxel_mismatch_rate differ from the C oracle."""
import json, os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import dvr_ref
from vidar_b200 import render
rng = np.random.default_rng(123)
Z, Y, X, M = 8, 50, 50, 6000
sigma = rng.uniform(0, 1, (1, 3, Z, Y, X)).astype(np.float32)
origin = np.array([[[25.0, 25.0, 4.0], [24.5, 25.5, 3.5], [24.37, 25.61, 3.52]]], np.float32)
points = (rng.uniform(0, 1, (1, M, 3)) * np.array([60, 60, 10]) - np.array([5, 5, 1])).astype(np.float32)
points[0, ::7] = np.round(points[0, ::7])
points[0, ::11] = np.round(points[0, ::11] * 2) / 2
tindex = rng.integers(0, 3, (1, M)).astype(np.float32)
dev = torch.device("cuda:0")
s, o, p, t = (torch.from_numpy(a).to(dev) for a in (sigma, origin, points, tindex))
out = {"eps": os.environ.get("VIDAR_DVR_TIE_EPS", "default")}
for name, f_gpu, f_ref in (("fwd_test", lambda: render.dvr.render_forward(s, o, p, t, [3, Z, Y, X], "test")[0],
                            lambda: dvr_ref.render_forward(sigma, origin, points, tindex, None, "test")[0]),
                           ("render_l1", lambda: render.dvr.render(s, o, p, t, "l1")[0],
                            lambda: dvr_ref.render(sigma, origin, points, tindex, "l1")[0])):
    a, b = f_gpu().cpu().numpy()[0], f_ref()[0]
    bad = np.flatnonzero(np.abs(a - b) > 1e-5 * np.maximum(1.0, np.abs(b)))
    out[name] = {"n_bad": int(bad.size), "rays": [dict(i=int(i), frame=int(tindex[0, i]), point=points[0, i].tolist(),
                                                        gpu=float(a[i]), ref=float(b[i])) for i in bad[:12]]}
print(json.dumps(out))
