#!/usr/bin/env python3
"""Calibration: can 8-bit / 16-bit / fp16 texel storage reproduce the fp32-texel bilinear samples exactly?"""
import os, sys, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
from tests import parity_common as pc
pkg = pc.load_pkg()
scene = pkg.scene.make_scene("C1")
params, mine, _ = pc.make_engines(pkg, scene, variants=())
rng = np.random.RandomState(5)
n = 400000
W, H = scene["W"], scene["H"]
xy = np.stack([rng.uniform(-3, W + 3, n), rng.uniform(-3, H + 3, n)], axis=1).astype(np.float32)
out = np.empty((4, n), np.float32); rate = np.empty(4, np.float32)
mine._ck(mine.lib.tsar_dbg_tex_formats(mine.h, 1, n, xy.ctypes.data, out.ctypes.data, rate.ctypes.data), "texfmt")
f32, u8, u16, f16 = out
res = dict(rate_gsamples=dict(f32=float(rate[0]), u8=float(rate[1]), u16=float(rate[2]), f16=float(rate[3])))
res["f16_equal"] = float((f16 == f32).mean())
for name, v, scale in (("u8", u8, 255.0 * 256.0), ("u16", u16, 65535.0 / 257.0 * 256.0)):
    snapped = np.rint(v.astype(np.float32) * np.float32(scale)).astype(np.float32) * np.float32(1 / 256.0)
    res[name + "_snapped_equal"] = float((snapped == f32).mean())
    res[name + "_max_abs"] = float(np.abs(snapped - f32).max())
print(json.dumps(res, indent=1))
np.savez_compressed(os.path.join(ROOT, "gpurun_out", "texfmt.npz"), xy=xy, out=out, img=scene["images"][1])
json.dump(res, open(os.path.join(ROOT, "gpurun_out", "texfmt.json"), "w"), indent=1)
