"""How far the reference build AS WRITTEN is from itself, and how far we are from it (VERDICT r01 task 9).

The reference's propagation launch races on same-colour pixels, so K runs of the unmodified build with
identical seeds give K different depthmaps.  This runs it K times at a config (default C2, 8 iterations), and reports
the distribution (min / median / max) over the K(K-1)/2 as-is pairs and over the K ours-vs-as-is pairs of the
north-star tolerance figures (1e-3 relative depth, 1 degree), for depth and normals separately and for textured and
untextured pixels separately.  Ours is checked bit for bit against the race-free twin in the same run.

    python tools/gpu_asis_distribution.py [C2] [K]   ->  gpurun_out/r02_asis_distribution_<cfg>.json
"""
import itertools
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
K = int(sys.argv[2]) if len(sys.argv) > 2 else 8
SEED = 20240601
pkg = ge.load_package()
L = pkg._lib
rb = pc.ref_binding()
import torch  # noqa: E402

scene = pkg.scene.make_scene(cfgname, backend="torch", device="cuda:0")
scene["images"] = [im.cpu().numpy() for im in scene["images"]]
torch.cuda.empty_cache()
params, mine, refs = pc.make_engines(pkg, scene, iterations=8, variants=("asis", "snapshot"))
mine.depthmap(SEED); o_m = mine.download(L.F_NORM4); mine.close()
refs["snapshot"].depthmap(SEED, iters=8); o_s = refs["snapshot"].download(rb.F_NORM4); refs["snapshot"].close()
runs = []
for k in range(K):
    refs["asis"].depthmap(SEED, iters=8)
    runs.append(refs["asis"].download(rb.F_NORM4).copy())
refs["asis"].close()
tex = torch.from_numpy(scene["region_text"][scene["labels"]] > 0).cuda()
dev = [torch.from_numpy(r).cuda() for r in runs]
d_m = torch.from_numpy(o_m).cuda()


def agreement(a, b):
    """Per-pixel gates of pc.output_agreement, evaluated on the device in float64."""
    za, zb = a[..., 3].double(), b[..., 3].double()
    na, nb = a[..., :3].double(), b[..., :3].double()
    valid = zb != 0
    rel = (za - zb).abs() / torch.where(valid, zb.abs(), torch.ones_like(zb))
    cosang = (na * nb).sum(-1) / (na.norm(dim=-1) * nb.norm(dim=-1)).clamp_min(1e-30)
    ang = torch.rad2deg(torch.arccos(cosang.clamp(-1, 1)))
    both_invalid = (~valid) & (za == 0)
    d_ok = ((rel <= pc.REL_DEPTH_TOL) & valid) | both_invalid
    a_ok = ((ang <= pc.ANGLE_TOL_DEG) & valid) | both_invalid
    bits = (a.view(torch.int32) == b.view(torch.int32)).all(-1)
    out = {}
    for name, m in (("all", None), ("textured", tex), ("untextured", ~tex)):
        sel = (lambda t: t) if m is None else (lambda t: t[m])
        out[name] = dict(both=float(sel(d_ok & a_ok).double().mean()), depth=float(sel(d_ok).double().mean()),
                         normal=float(sel(a_ok).double().mean()), bit_exact=float(sel(bits).double().mean()))
    return out


def spread(rows):
    out = {}
    for region in ("all", "textured", "untextured"):
        out[region] = {}
        for key in ("both", "depth", "normal", "bit_exact"):
            v = np.array([r[region][key] for r in rows])
            out[region][key] = dict(min=float(v.min()), median=float(np.median(v)), max=float(v.max()))
    return out


pairs = [agreement(dev[i], dev[j]) for i, j in itertools.combinations(range(K), 2)]
ours = [agreement(d_m, dev[i]) for i in range(K)]
gt = [pc.gt_agreement(r, scene)["frac_within_1pct_textured"] for r in runs]
res = {"config": cfgname, "iterations": 8, "runs_of_the_reference_as_written": K, "pairs": len(pairs),
       "tolerance": {"relative_depth": pc.REL_DEPTH_TOL, "normal_angle_deg": pc.ANGLE_TOL_DEG},
       "ours_vs_race_free_twin_bit_exact": pc.frac_bit_exact(o_m, o_s),
       "asis_vs_asis": spread(pairs), "ours_vs_asis": spread(ours),
       "textured_fraction": float(tex.double().mean()),
       "within_1pct_of_ground_truth_textured": {"ours": pc.gt_agreement(o_m, scene)["frac_within_1pct_textured"],
                                                "asis_min": min(gt), "asis_max": max(gt)}}
print(json.dumps(res, indent=1))
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
json.dump(res, open(os.path.join(ROOT, "gpurun_out", f"r02_asis_distribution_{cfgname}.json"), "w"), indent=1)
