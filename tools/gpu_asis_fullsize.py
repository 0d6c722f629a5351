"""Agreement of our output with the reference build AS WRITTEN (racy) at benchmark size: C2-shaped scene, 8
iterations, same seeds.  Reports the north-star tolerance figures (1e-3 relative depth, 1 degree) over all pixels and
over the textured facets, next to the as-is build's own run-to-run agreement and the race-free twin."""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import __graft_entry__ as ge  # noqa: E402
import parity_common as pc  # noqa: E402

cfgname = sys.argv[1] if len(sys.argv) > 1 else "C2"
pkg = ge.load_package()
L = pkg._lib
rb = pc.ref_binding()
scene = pkg.scene.make_scene(cfgname, backend="torch", device="cuda:0")
scene["images"] = [im.cpu().numpy() for im in scene["images"]]
params, mine, refs = pc.make_engines(pkg, scene, iterations=8, variants=("asis", "snapshot"))
seed = 20240601
mine.depthmap(seed); o_m = mine.download(L.F_NORM4); mine.close()
refs["snapshot"].depthmap(seed, iters=8); o_s = refs["snapshot"].download(rb.F_NORM4); refs["snapshot"].close()
refs["asis"].depthmap(seed, iters=8); o_a = refs["asis"].download(rb.F_NORM4)
refs["asis"].depthmap(seed, iters=8); o_a2 = refs["asis"].download(rb.F_NORM4); refs["asis"].close()
tex = scene["region_text"][scene["labels"]] > 0


def masked(a, b, m):
    x, y = a.copy(), b.copy()
    return pc.output_agreement(x[m][None], y[m][None])


res = {"config": cfgname, "ours_vs_snapshot_bit_exact": pc.frac_bit_exact(o_m, o_s),
       "ours_vs_asis": pc.output_agreement(o_m, o_a), "asis_vs_asis": pc.output_agreement(o_a2, o_a),
       "ours_vs_asis_textured": masked(o_m, o_a, tex), "asis_vs_asis_textured": masked(o_a2, o_a, tex),
       "gt_ours": pc.gt_agreement(o_m, scene), "gt_asis": pc.gt_agreement(o_a, scene), "textured_fraction": float(tex.mean())}
print(json.dumps(res, indent=1))
json.dump(res, open(os.path.join(ROOT, "gpurun_out", f"asis_fullsize_{cfgname}.json"), "w"), indent=1)
