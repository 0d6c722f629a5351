"""Times gipuma_WMF (4 levels) and gipuma_WMF_Final (6 levels) -- ours vs the reference kernels -- on a
1920x1080 view (development tooling; results under profiles/)."""
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
rb = pc.ref_binding()
cfg = dict(W=1920, H=1080, n_images=3, V=2, fx=1160.0, radius=5.0, arc_deg=10.0)
scene = pkg.scene.make_scene(cfg)
params, mine, refs = pc.make_engines(pkg, scene, iterations=2, variants=("snapshot",))
ref = refs["snapshot"]
ref.init_planes(7); ref.iterate(2, 7); ref.lrdiff(); ref.getview()
n0, c0, d0 = ref.download(rb.F_NORM4), ref.download(rb.F_COST), ref.download(rb.F_DEPTH)
reliable = (c0 < 0.25).astype(np.float32)
mine.load_planes(n0, c0); mine.upload(L.F_DEPTH, d0)
for e in (mine, ref):
    e.set_regions(scene["region_text"], scene["region_norm4"])
mine.upload(L.F_CANNY, scene["canny"]); ref.upload(rb.F_CANNY, scene["canny"])
res = {"reliable_fraction": float(reliable.mean())}


def timed(fn, sync):
    sync()
    t0 = time.perf_counter()
    fn()
    sync()
    return (time.perf_counter() - t0) * 1e3


for name, n_it, call_m, call_r in (("wmf", 4, mine.wmf, ref.wmf), ("wmf_final", 6, mine.wmf_final, ref.wmf_final)):
    for rep in range(2):   # second repetition is the warm one
        mine.upload(L.F_SCALE, reliable); ref.upload(rb.F_SCALE, reliable)
        tm = [timed(lambda it=it: call_m(it), mine.sync) for it in range(n_it)]
        tr = [timed(lambda it=it: call_r(it), lambda: ref.download(rb.F_SCALE)) for it in range(n_it)]
    res[name] = {"ours_ms": tm, "reference_ms": tr, "ours_total": sum(tm), "reference_total": sum(tr)}
    print(name, json.dumps(res[name]), flush=True)
json.dump(res, open(os.path.join(ROOT, "gpurun_out", "wmf_time.json"), "w"), indent=1)
mine.close(); ref.close()
