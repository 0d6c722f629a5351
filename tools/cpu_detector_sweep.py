#!/usr/bin/env python3
"""One-off differential stress of the weak-texture detector (texture.detect = tsar_weak_* host functions + cv2) against the
reference's whole texture() (main.cpp:365-596 in oracle/_ref/libtsar_ref_host.so), on the CPU: rendered views, noise,
piecewise-constant mosaics with large flat cells (many weak regions, Hough lines), gradients, saturated images, sizes that
are not multiples of 4 and down to 64 x 48.  The full-resolution label map and every region statistic must be identical.

    python tools/cpu_detector_sweep.py [N]   ->  profiles/r02_detector_sweep.json
"""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as g  # noqa: E402

pkg = g.load_package()
from oracle import ref_host_binding as rh  # noqa: E402
from tsar_mvs_b200 import scene, texture as tx  # noqa: E402

N = int(sys.argv[1]) if len(sys.argv) > 1 else 60
rng = np.random.RandomState(4711)
rows, bad = [], 0
t0 = time.time()
for trial in range(N):
    W, H = int(rng.randint(64, 1500)), int(rng.randint(48, 1100))
    kind = ("render", "noise", "mosaic", "gradient", "saturated")[trial % 5]
    if kind == "render":
        cams = scene.make_cameras(W, H, 2, 0.5625 * W, 4.0, 10.0)
        img = scene.Scene(W, H, 0.5625 * W, 4.0, seed=int(rng.randint(1, 9999))).render(cams[0], W, H)[0].astype(np.uint8)
        if rng.rand() < 0.5:
            img[H // 8:H - H // 6, W // 10:W - W // 7] = int(rng.randint(0, 256))
    elif kind == "noise":
        img = rng.randint(0, 256, (H, W)).astype(np.uint8)
    elif kind == "mosaic":
        cell = int(rng.choice([24, 60, 150, 400]))
        blk = rng.randint(0, 256, (H // cell + 1, W // cell + 1)).astype(np.uint8)
        img = np.repeat(np.repeat(blk, cell, 0), cell, 1)[:H, :W].copy()
        img[rng.rand(H, W) < 0.002] = 255                          # speckles: tiny components
    elif kind == "gradient":
        yy, xx = np.mgrid[0:H, 0:W]
        img = ((xx * rng.uniform(0.05, 0.6) + yy * rng.uniform(0.05, 0.6)) % 256).astype(np.uint8)
    else:
        img = np.full((H, W), 255, np.uint8)
        img[H // 3:H // 3 + 5] = 0
        img[:, W // 2:W // 2 + 3] = 17
    ref = rh.texture(img)
    det = tx.detect(img)
    ok = np.array_equal(tx.expand_labels(det["labels_q"], W, H), ref["canny"]) and all(np.array_equal(det[k], ref[k]) for k in ("text", "cenxi", "cenyi", "size"))
    row = dict(trial=trial, W=W, H=H, image=kind, regions=int(len(ref["text"])), weak_regions=int((ref["text"] == -1).sum()), identical=bool(ok))
    bad += 0 if ok else 1
    rows.append(row)
    print(("ok  " if ok else "FAIL"), json.dumps(row), flush=True)
res = dict(trials=N, identical_trials=N - bad, seconds=round(time.time() - t0, 1), rows=rows)
json.dump(res, open(os.path.join(ROOT, "profiles", "r02_detector_sweep.json"), "w"), indent=1)
print(f"{N - bad} of {N} images identical")
sys.exit(1 if bad else 0)
