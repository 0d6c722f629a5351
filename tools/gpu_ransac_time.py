"""Times tsar_fit_region_planes on a C4-shaped reference view (one large textureless facet, > 50 000 reliable
pixels -> capped at 49 999 points as in the reference) against the scalar C restatement on the host."""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import __graft_entry__ as ge  # noqa: E402
import parity_common as pc  # noqa: E402

pkg = ge.load_package()
L = pkg._lib
from oracle import cpu_binding as cb  # noqa: E402
from tsar_mvs_b200.engine import cameras_to_struct  # noqa: E402

cfg = dict(W=1920, H=1080, n_images=2, V=1, fx=1160.0, radius=5.0, arc_deg=10.0)
scene = pkg.scene.make_scene(cfg)
params, mine, _ = pc.make_engines(pkg, scene, variants=())
H, W = scene["H"], scene["W"]
rng = np.random.RandomState(9)
disp = (scene["cam_f"] / scene["gt_depth"]).astype(np.float32)
disp *= (1 + 0.0005 * rng.normal(size=disp.shape)).astype(np.float32)
out = rng.rand(H, W) < 0.2
disp[out] *= rng.uniform(0.8, 1.2, out.sum()).astype(np.float32)
scale = (rng.rand(H, W) < 0.6).astype(np.float32)
text = scene["region_text"].copy()
size = np.array([(scene["labels"] == r).sum() / 16.0 for r in range(len(text))], np.float32)
per = mine.lib.tsar_ransac_rand_per_region()
rnd = rng.randint(0, 2 ** 31 - 1, size=(len(text), per)).astype(np.uint32)
mine.upload(L.F_DEPTH, disp); mine.upload(L.F_SCALE, scale); mine.upload(L.F_CANNY, scene["canny"])
p0 = np.tile(np.array([0, 0, 1, -1], np.float32), (len(text), 1))
res = {}
for n_fit in (1, 3):
    t = text.copy()
    t[:] = 1.0
    order = np.argsort(-size)
    t[order[:n_fit]] = -1.0
    npts = [int(((scene["canny"] == r) & (scale == 1)).sum()) for r in order[:n_fit]]
    mine.fit_region_planes(t, size, rnd, p0)     # warm-up
    t0 = time.perf_counter()
    fitted = mine.fit_region_planes(t, size, rnd, p0)
    dt = time.perf_counter() - t0
    res[f"gpu_ms_{n_fit}_regions"] = dt * 1e3
    res[f"points_{n_fit}_regions"] = npts
    print(n_fit, "regions", npts, "points:", dt * 1e3, "ms", flush=True)
cams = cameras_to_struct(scene["cams"])
r = int(order[0])
t0 = time.perf_counter()
want, used = cb.fit_region_plane(pkg._lib.TsarCamera, cams[0], scene["cam_f"], disp, scale, scene["canny"], r, size[r], rnd[r], p0[r])
res["cpu_ms_1_region"] = (time.perf_counter() - t0) * 1e3
res["cpu_points"] = int(used)
res["bit_exact_vs_cpu"] = bool(np.array_equal(fitted[r], want))
mine.close()

# ---- a C2 view with the weak-texture detector's REAL region table (thousands of regions, a few of them weak) ----------
import torch  # noqa: E402
from tsar_mvs_b200 import texture as tx  # noqa: E402
sc2 = pkg.scene.make_scene("C2", backend="torch", device="cuda:0")
imgs = [im.contiguous() for im in sc2["images"]]
ref_u8 = imgs[0].cpu().numpy().astype(np.uint8)
t0 = time.perf_counter()
det = tx.detect(ref_u8)
res["c2_detector_host_ms_first_call"] = (time.perf_counter() - t0) * 1e3     # includes importing OpenCV
t0 = time.perf_counter()
det = tx.detect(ref_u8)
res["c2_detector_host_ms"] = (time.perf_counter() - t0) * 1e3
res["c2_regions"] = int(len(det["text"]))
res["c2_weak_regions"] = int((det["text"] == -1).sum())
eng = pkg.DepthmapEngine(0)
eng.set_views_device([t.data_ptr() for t in imgs], sc2["W"], sc2["H"], cameras_to_struct(sc2["cams"]), sc2["subset"], cam_f=sc2["cam_f"])
eng.set_params(pkg.make_params(box=11, iterations=8, min_disparity=sc2["min_disparity"], max_disparity=sc2["max_disparity"]))
eng.set_labels_quarter(det["labels_q"])
eng.init_planes(1); eng.iterate(8, 1); eng.lrdiff(); eng.getview()
confid = eng.download(L.F_CONFID)
eng.upload(L.F_SCALE, (confid > 0.8).astype(np.float32))
p0 = np.tile(np.array([0, 0, 1, -1], np.float32), (len(det["text"]), 1))
eng.fit_region_planes(det["text"], det["size"], None, p0, seed=5)       # warm-up (allocates the persistent scratch)
n0 = eng.launch_count()
t0 = time.perf_counter()
planes = eng.fit_region_planes(det["text"], det["size"], None, p0, seed=5)
res["c2_fit_ms_all_weak_regions_seeded_stream"] = (time.perf_counter() - t0) * 1e3
res["c2_fit_kernel_launches"] = int(eng.launch_count() - n0)
lab_full = tx.expand_labels(det["labels_q"], sc2["W"], sc2["H"])
res["c2_weak_region_reliable_points"] = [int(((lab_full == r) & (confid > 0.8)).sum()) for r in np.nonzero(det["text"] == -1)[0]]
eng.close()
print(json.dumps(res))
json.dump(res, open(os.path.join(ROOT, "gpurun_out", "ransac_time.json"), "w"), indent=1)
