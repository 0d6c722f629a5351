#!/usr/bin/env python3
"""(The launch-shape variants c g h i e f need the experimental pm_inst_w11<x>.cu translation units of the commit that
introduced this note; the default build only has the u8/f32 texel variants.)
Times the launch-shape variants of the 11x11 PatchMatch kernels on a bench-sized scene and checks that
they all give identical output (development tooling)."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch
from tests import parity_common as pc

cfg = sys.argv[1] if len(sys.argv) > 1 else "C2"
variants = sys.argv[2].split(",") if len(sys.argv) > 2 else ["u8", "f32"]   # u8 | f32 | w11 launch shapes: c g h i e f
pkg = pc.load_pkg()
L = pkg._lib
scene = pkg.scene.make_scene(cfg, backend="torch", device="cuda:0")
imgs = [im.contiguous() for im in scene["images"]]
from tsar_mvs_b200.engine import cameras_to_struct
cams = cameras_to_struct(scene["cams"])
params = pkg.make_params(box=11, iterations=8, min_disparity=scene["min_disparity"], max_disparity=scene["max_disparity"])
res, base = {}, None
for v in variants:
    os.environ["TSAR_B200_NO_U8"] = "1" if v == "f32" else "0"
    os.environ["TSAR_B200_W11_VARIANT"] = v if v in ("c", "g", "h", "i", "e", "f", "q", "r") else ""
    eng = pkg.DepthmapEngine(0)
    eng.set_views_device([t.data_ptr() for t in imgs], scene["W"], scene["H"], cams, scene["subset"], cam_f=scene["cam_f"])
    eng.set_params(params)
    ms = [eng.depthmap(20240601) for _ in range(3)]
    eng.profile(True); eng.depthmap(20240601); chk_ms, n = eng.profile_read()
    out = eng.download(L.F_NORM4)
    if base is None:
        base = out
    res[v] = dict(ms=ms, checker_ms_avg=chk_ms / n, same_as_first=pc.frac_bit_exact(out, base),
                  gt=pc.gt_agreement(out, scene))
    print(v, json.dumps(res[v]), flush=True)
    eng.close()
json.dump(res, open(os.path.join(ROOT, "gpurun_out", f"variants_{cfg}.json"), "w"), indent=1)
