#!/usr/bin/env python3
"""Pins the tie rule of the texture model (DESIGN.md "Texture model"): the bilinear weights have 8 fractional bits; random
coordinates almost never land exactly between two weight steps, so the calibration with random samples does not say
whether the unit rounds such ties up, to even, or truncates.  Samples a view at coordinates i + 0.5 + k/512 (every
half step of the weight grid, k odd = exact tie) and further exact dyadic offsets (k/1024, k/4096) in x and y, fp32
texels and 8-bit texels, and compares with the model under the three candidate rules.

    python tools/gpu_tex_ties.py   ->  gpurun_out/r02_texture_tie_rule.json
"""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from tests import parity_common as pc  # noqa: E402

pkg = pc.load_pkg()
scene = pkg.scene.make_scene("C1")
params, mine, _ = pc.make_engines(pkg, scene, variants=())
img = np.ascontiguousarray(scene["images"][1], np.float32)
H, W = img.shape
rng = np.random.RandomState(17)


def model(x, y, rule):
    """oracle/oracle_cpu.c oc_tex with a selectable rounding of the 8-bit weights."""
    xb = (x.astype(np.float32) - np.float32(0.5)).astype(np.float32)
    yb = (y.astype(np.float32) - np.float32(0.5)).astype(np.float32)
    fi, fj = np.floor(xb), np.floor(yb)
    fx, fy = (xb - fi).astype(np.float64) * 256.0, (yb - fj).astype(np.float64) * 256.0
    if rule == "half_up":
        ax, ay = np.floor(fx + 0.5), np.floor(fy + 0.5)
    elif rule == "half_even":
        ax, ay = np.rint(fx), np.rint(fy)
    else:
        ax, ay = np.floor(fx), np.floor(fy)
    ax, ay = ax.astype(np.int64), ay.astype(np.int64)
    i, j = fi.astype(np.int64), fj.astype(np.int64)
    w11 = (ax * ay + 128) >> 8
    w10, w01 = ax - w11, ay - w11
    w00 = 256 - ax - ay + w11
    i0, i1 = np.clip(i, 0, W - 1), np.clip(i + 1, 0, W - 1)
    j0, j1 = np.clip(j, 0, H - 1), np.clip(j + 1, 0, H - 1)
    r = w00 * img[j0, i0].astype(np.int64) + w10 * img[j0, i1].astype(np.int64) + w01 * img[j1, i0].astype(np.int64) + w11 * img[j1, i1].astype(np.int64)
    return (r.astype(np.float32) * np.float32(1 / 256.0)).astype(np.float32)


def sample(x, y):
    n = len(x)
    xy = np.stack([x, y], axis=1).astype(np.float32)
    out = np.empty((4, n), np.float32)
    rate = np.empty(4, np.float32)
    mine._ck(mine.lib.tsar_dbg_tex_formats(mine.h, 1, n, xy.ctypes.data, out.ctypes.data, rate.ctypes.data), "texfmt")
    f32 = out[0]
    u8 = np.rint(out[1].astype(np.float32) * np.float32(255.0 * 256.0)).astype(np.float32) * np.float32(1 / 256.0)
    return f32, u8


res = {"image": f"{W}x{H} synthetic view (8-bit valued)", "sets": {}}
n = 300000
for name, den in (("k/512 (every tie and every weight step)", 512), ("k/1024", 1024), ("k/4096", 4096), ("k/65536", 65536)):
    ix = rng.randint(-2, W + 2, n)
    iy = rng.randint(-2, H + 2, n)
    kx, ky = rng.randint(0, den, n), rng.randint(0, den, n)
    x = (ix + 0.5 + kx / den).astype(np.float32)
    y = (iy + 0.5 + ky / den).astype(np.float32)
    f32, u8 = sample(x, y)
    ties = ((kx * 512) % den == 0) & (((kx * 512) // den) % 2 == 1) | ((ky * 512) % den == 0) & (((ky * 512) // den) % 2 == 1)
    row = {"samples": n, "samples_on_a_tie": int(ties.sum()), "u8_equals_f32_texels": float((u8 == f32).mean())}
    for rule in ("half_up", "half_even", "truncate"):
        m = model(x, y, rule)
        row[rule] = {"all": float((m == f32).mean()), "ties_only": float((m[ties] == f32[ties]).mean()) if ties.any() else None}
    res["sets"][name] = row
# ties where half-up and half-even differ: even weight step below the tie
best = max(("half_up", "half_even", "truncate"), key=lambda r: min(v[r]["all"] for v in res["sets"].values()))
res["rule_that_matches_everywhere"] = best if all(v[best]["all"] == 1.0 for v in res["sets"].values()) else None
print(json.dumps(res, indent=1))
json.dump(res, open(os.path.join(ROOT, "gpurun_out", "r02_texture_tie_rule.json"), "w"), indent=1)
mine.close()

# golden vectors for the CPU suite: hardware samples of the tiny scene's view 1 on the weight grid and on its exact ties
tiny = pkg.scene.make_scene("tiny")
_, eng, _ = pc.make_engines(pkg, tiny, variants=())
th, tw = tiny["images"][1].shape
m = 4000
gx = (rng.randint(-2, tw + 2, m) + 0.5 + rng.randint(0, 512, m) / 512.0).astype(np.float32)
gy = (rng.randint(-2, th + 2, m) + 0.5 + rng.randint(0, 512, m) / 512.0).astype(np.float32)
gxy = np.stack([gx, gy], axis=1).astype(np.float32)
gout = np.empty(m, np.float32)
eng._ck(eng.lib.tsar_dbg_tex_sample(eng.h, 1, m, gxy.ctypes.data, gout.ctypes.data), "tex")
os.makedirs(os.path.join(ROOT, "gpurun_out", "golden"), exist_ok=True)
np.savez_compressed(os.path.join(ROOT, "gpurun_out", "golden", "tiny_tex_ties.npz"), xy=gxy, out=gout)
eng.close()
