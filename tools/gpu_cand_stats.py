"""Device-side counter of executed cost evaluations and duplicate propagation candidates, per checkerboard launch
(VERDICT r01 task 2).  Runs the per-view sequence of a config with the half-steps launched one by
one (bit-identical to the fused sequence) and asks tsar_dbg_candidate_stats before every propagation launch.

    python tools/gpu_cand_stats.py [C2] [iters]    ->  gpurun_out/cand_stats_<cfg>.json
"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as ge  # noqa: E402

cfgname = sys.argv[1] if len(sys.argv) > 1 else "C2"
iters = int(sys.argv[2]) if len(sys.argv) > 2 else 8
SEED = 20240601
pkg = ge.load_package()
L = pkg._lib
from tsar_mvs_b200.engine import cameras_to_struct  # noqa: E402

scene = pkg.scene.make_scene(cfgname, backend="torch", device="cuda:0")
imgs = [im.contiguous() for im in scene["images"]]
cfg = pkg.scene.CONFIGS[cfgname]
W, H, V = cfg["W"], cfg["H"], cfg["V"]
params = pkg.make_params(box=11, iterations=iters, min_disparity=scene["min_disparity"], max_disparity=scene["max_disparity"])
eng = pkg.DepthmapEngine(0)
eng.set_views_device([t.data_ptr() for t in imgs], W, H, cameras_to_struct(scene["cams"]), scene["subset"], cam_f=scene["cam_f"])
eng.set_params(params)
eng.init_planes(SEED)
R = 0
dz = scene["max_disparity"] * 0.5
while dz >= 0.01:
    R += 1
    dz /= 10.0
rows = []
for it in range(iters):
    for col in range(2):
        st = eng.candidate_stats(col)
        st.update(iteration=it, colour=col)
        rows.append(st)
        eng.launch(L.BLACK_SPATIAL if col == 0 else L.RED_SPATIAL)
        eng.launch(L.BLACK_REFINE if col == 0 else L.RED_REFINE, SEED + 1 + 2 * it + col)
closed_form = eng.eval_count(iters)
eng.close()
tot = {k: sum(r[k] for r in rows) for k in rows[0] if isinstance(rows[0][k], int) and k not in ("iteration", "colour")}
executed = V * (W * H + tot["in_depth_range"] + R * tot["pixels"])
with_skip = V * (W * H + tot["distinct"] + R * tot["pixels"])
res = {
    "config": cfgname, "W": W, "H": H, "V": V, "iterations": iters, "refinement_rounds": R,
    "closed_form_evaluations": closed_form,
    "closed_form_check": V * (W * H + tot["behind_border_guards"] + R * tot["pixels"]),
    "executed_evaluations_as_written": executed,
    "executed_evaluations_skipping_duplicates": with_skip,
    "propagation_candidates": {k: tot[k] for k in ("behind_border_guards", "in_depth_range", "dup_of_own_plane", "dup_of_earlier_candidate", "distinct",
                                                     "distinct_without_own_rule")},
    "warp_rounds_propagation": {k: tot[k] for k in ("warp_rounds_as_written", "warp_rounds_lane_lists", "warp_rounds_packed",
                                                      "warp_rounds_lane_lists_without_own_rule", "warp_rounds_packed_without_own_rule", "warps")},
    "warp_rounds_refinement": R * tot["warps"],
    "per_launch": rows,
}
wr = res["warp_rounds_propagation"]
ref_rounds = res["warp_rounds_refinement"]
res["launch_time_model"] = {
    "as_written": 1.0,
    "lane_lists": (wr["warp_rounds_lane_lists"] + ref_rounds) / (wr["warp_rounds_as_written"] + ref_rounds),
    "packed": (wr["warp_rounds_packed"] + ref_rounds) / (wr["warp_rounds_as_written"] + ref_rounds),
    "lane_lists_without_own_rule": (wr["warp_rounds_lane_lists_without_own_rule"] + ref_rounds) / (wr["warp_rounds_as_written"] + ref_rounds),
    "packed_without_own_rule": (wr["warp_rounds_packed_without_own_rule"] + ref_rounds) / (wr["warp_rounds_as_written"] + ref_rounds),
}
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
json.dump(res, open(os.path.join(ROOT, "gpurun_out", f"cand_stats_{cfgname}.json"), "w"), indent=1)
print(json.dumps({k: v for k, v in res.items() if k != "per_launch"}, indent=1))
for r in rows:
    print(r["iteration"], r["colour"], "range", r["in_depth_range"], "dup_own", r["dup_of_own_plane"], "dup_prev", r["dup_of_earlier_candidate"],
          "rounds", r["warp_rounds_as_written"], r["warp_rounds_lane_lists"], r["warp_rounds_packed"], "hist", r["pixels_by_distinct"])
