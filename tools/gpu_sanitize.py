"""Exercises every kernel of libtsar_b200.so on small inputs (our library only, no reference build), optionally
under a memory checker where the pool allows one:
    compute-sanitizer --tool memcheck --error-exitcode 9 python tools/gpu_sanitize.py
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as ge  # noqa: E402

pkg = ge.load_package()
L = pkg._lib
from tsar_mvs_b200.engine import cameras_to_struct  # noqa: E402


def engine_for(scene, **kw):
    params = pkg.make_params(min_disparity=scene["min_disparity"], max_disparity=scene["max_disparity"], **kw)
    e = pkg.DepthmapEngine(0)
    e.set_views(scene["images"], cameras_to_struct(scene["cams"]), scene["subset"], cam_f=scene["cam_f"])
    e.set_params(params)
    return e


def main():
    rng = np.random.RandomState(3)
    odd = dict(W=67, H=33, n_images=3, V=2, fx=150.0, radius=1.0, arc_deg=14.0)
    for cfg, kws in (("tiny", [dict(box=11, iterations=2), dict(box=19, iterations=1), dict(box=7, iterations=1, n_best=3),
                               dict(box=12, iterations=1, n_best=2, cost_comb=0)]),
                     (odd, [dict(box=11, iterations=2), dict(box=5, iterations=1)])):
        scene = pkg.scene.make_scene(cfg, with_colour=True)
        H, W = scene["H"], scene["W"]
        for kw in kws:
            e = engine_for(scene, **kw)
            e.depthmap(11)                                   # rng rows, init, checker (fused), lrdiff, getview, compute_disp
            out = e.download(L.F_NORM4)
            assert np.isfinite(out).all()
            os.environ["TSAR_B200_UNFUSED"] = "1"
            e2 = engine_for(scene, **kw)
            e2.depthmap(11)                                  # unfused launches + merge kernel
            del os.environ["TSAR_B200_UNFUSED"]
            assert np.array_equal(out, e2.download(L.F_NORM4))
            e2.close()
            # explicit-plane evaluation
            n = 500
            xy = np.stack([rng.randint(0, W, n), rng.randint(0, H, n)], 1).astype(np.int32)
            nrm = rng.normal(size=(n, 3)); nrm[:, 2] = -np.abs(nrm[:, 2]) - 0.3
            nrm /= np.linalg.norm(nrm, axis=1, keepdims=True)
            pl = np.concatenate([nrm, rng.uniform(0.7, 1.4, (n, 1))], 1).astype(np.float32)
            e.eval_planes(xy, pl)
            # glue + depth completion
            e.init_planes(5); e.iterate(1, 5); e.lrdiff(); e.getview()
            e.set_regions(scene["region_text"], scene["region_norm4"])
            e.upload(L.F_CANNY, scene["canny"])
            e.update_scale_2(); e.update_scale(); e.compute_disp()
            wn = e.download(L.F_NORM4)
            dsp = np.where(wn[..., 3] > 0, scene["cam_f"] / np.maximum(wn[..., 3], 1e-6), 1.0).astype(np.float32)
            e.upload(L.F_DEPTH, dsp); e.get_disp()
            # weighted-median stages
            e.upload(L.F_SCALE, (rng.rand(H, W) < 0.6).astype(np.float32))
            for it in range(4):
                e.wmf(it)
            e.upload(L.F_SCALE, (rng.rand(H, W) < 0.6).astype(np.float32))
            for it in range(3):
                e.wmf_final(it)
            # region plane fit
            text = scene["region_text"].copy(); text[1] = -1.0
            size = np.array([(scene["labels"] == r).sum() / 16.0 for r in range(len(text))], np.float32)
            per = e.lib.tsar_ransac_rand_per_region()
            rnd = rng.randint(0, 2 ** 31 - 1, size=(len(text), per)).astype(np.uint32)
            e.upload(L.F_DEPTH, (scene["cam_f"] / scene["gt_depth"]).astype(np.float32))
            e.upload(L.F_SCALE, (rng.rand(H, W) < 0.5).astype(np.float32))
            p0 = np.tile(np.array([0, 0, 1, -1], np.float32), (len(text), 1))
            e.fit_region_planes(text, size, rnd, p0)
            e.fit_region_planes(text, size, None, p0, seed=5)          # device-generated random stream
            # reliable-pixel sources of the flows, quarter-resolution labels, candidate counters
            e.scale_from_confidence(0.8)
            png = np.zeros((H, W, 3), np.uint8); png[::2] = 255; png[1::4, :, 1] = 255
            e.scale_from_weak_png(png)
            e.set_labels_quarter(rng.randint(0, 5, size=((H + 3) // 4, (W + 3) // 4)).astype(np.int32))
            e.init_planes(7)
            for colour in (0, 1):
                st = e.candidate_stats(colour)
                assert st["distinct"] <= st["in_depth_range"] <= st["behind_border_guards"]
            # gSLICr, all modes
            bgr = scene["bgr"]
            bgrx = np.concatenate([bgr, np.zeros(bgr.shape[:2] + (1,), np.uint8)], -1)
            for enforce, correct in ((False, False), (True, False), (False, True)):
                e.slic(bgrx, spixel_size=8, enforce_connectivity=enforce, correct_reduction=correct)
            # end-to-end host call
            e.depthmap_host(scene["images"], scene["cams"], scene["subset"], e.params, 3, cam_f=scene["cam_f"])
            e.close()
            print("ok", cfg if isinstance(cfg, str) else "odd", kw, flush=True)
    print("SANITIZE-SCRIPT-DONE")


if __name__ == "__main__":
    main()
