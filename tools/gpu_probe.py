#!/usr/bin/env python3
"""GPU probe (development tooling): runs the reference-kernel oracle and the product library side by
side on synthetic scenes and writes everything we need to read offline into gpurun_out/.
Usage (on the GPU box):  python tools/gpu_probe.py [stage ...]
"""
import ctypes as C
import json
import os
import sys
import time
import traceback

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from tests import parity_common as pc  # noqa: E402

OUT = os.path.join(ROOT, "gpurun_out")
os.makedirs(OUT, exist_ok=True)
RES = {}


def log(*a):
    print(*a, flush=True)


def stage(fn):
    name = fn.__name__
    t = time.time()
    try:
        RES[name] = fn()
    except Exception as e:  # keep going: every stage is independent
        RES[name] = {"error": repr(e), "trace": traceback.format_exc()}
        log(f"[{name}] FAILED: {e!r}")
        log(traceback.format_exc())
    RES[name + "_s"] = round(time.time() - t, 2)
    with open(os.path.join(OUT, "probe.json"), "w") as f:
        json.dump(RES, f, indent=1, default=float)


pkg = pc.load_pkg()
rb = pc.ref_binding()
L = pkg._lib
SEED = 20240601


def s_peaks():
    scene = pkg.scene.make_scene("small")
    params, mine, _ = pc.make_engines(pkg, scene, variants=())
    out = (C.c_float * 3)()
    mine._ck(mine.lib.tsar_dbg_peaks(mine.h, out), "peaks")
    mine.close()
    r = dict(ffma_tflops=out[0], mufu_gops=out[1], tex_gsamples=out[2])
    log("[peaks]", r)
    return r


def s_texcal():
    """Raw texture samples for calibrating a software bilinear model."""
    scene = pkg.scene.make_scene("tiny")
    params, mine, _ = pc.make_engines(pkg, scene, variants=())
    rng = np.random.RandomState(3)
    n = 200000
    W, H = scene["W"], scene["H"]
    xy = np.stack([rng.uniform(-3, W + 3, n), rng.uniform(-3, H + 3, n)], axis=1).astype(np.float32)
    # a block of samples on a fine grid inside one texel to read the weight quantisation
    k = 4096
    xy[:k, 0] = 10.5 + np.arange(k) / k
    xy[:k, 1] = 7.5
    xy[k:2 * k, 0] = 20.5
    xy[k:2 * k, 1] = 9.5 + np.arange(k) / k
    out = np.empty(n, np.float32)
    mine._ck(mine.lib.tsar_dbg_tex_sample(mine.h, 1, n, xy.ctypes.data, out.ctypes.data), "tex_sample")
    np.savez_compressed(os.path.join(OUT, "texcal.npz"), xy=xy, out=out, img=scene["images"][1])
    mine.close()
    return dict(n=n)


def s_eval():
    res = {}
    for cfg, box, nb in (("small", 11, 1), ("small", 11, 2), ("small", 19, 1), ("small", 7, 3)):
        scene = pkg.scene.make_scene(cfg)
        params, mine, refs = pc.make_engines(pkg, scene, box=box, n_best=nb, variants=("asis",))
        ref = refs["asis"]
        xy, planes = pc.random_planes(scene, 50000)
        c_m, b_m, r_m = mine.eval_planes(xy, planes, wrapper_rounding=True)
        c_r, b_r, r_r = ref.eval_planes(xy, planes)
        key = f"{cfg}_box{box}_nbest{nb}"
        d = np.abs(c_m.astype(np.float64) - c_r)
        res[key] = dict(cost_bit_exact=pc.frac_bit_exact(c_m, c_r), beview_equal=float((b_m == b_r).mean()),
                        ratio_bit_exact=pc.frac_bit_exact(r_m, r_r), max_abs=float(d.max()), mean_cost=float(c_r.mean()),
                        frac_valid=float((c_r < 2).mean()))
        log("[eval]", key, res[key])
        if key == "small_box11_nbest1":
            bad = np.nonzero(~pc.bits_equal(c_m, c_r))[0][:2000]
            np.savez_compressed(os.path.join(OUT, "eval_small.npz"), xy=xy, planes=planes, c_m=c_m, c_r=c_r, b_m=b_m,
                                b_r=b_r, r_m=r_m, r_r=r_r, bad=bad)
        mine.close(); ref.close()
    return res


def s_init():
    scene = pkg.scene.make_scene("small")
    params, mine, refs = pc.make_engines(pkg, scene, variants=("asis",))
    ref = refs["asis"]
    mine.init_planes(SEED); ref.init_planes(SEED)
    n_m, c_m = mine.download(L.F_NORM4), mine.download(L.F_COST)
    n_r, c_r = ref.download(rb.F_NORM4), ref.download(rb.F_COST)
    r = dict(norm4_bit_exact=pc.frac_bit_exact(n_m, n_r), cost_bit_exact=pc.frac_bit_exact(c_m, c_r),
             xyz_bit_exact=pc.frac_bit_exact(n_m[..., :3], n_r[..., :3]), w_bit_exact=pc.frac_bit_exact(n_m[..., 3], n_r[..., 3]))
    np.savez_compressed(os.path.join(OUT, "init_small.npz"), n_m=n_m, n_r=n_r, c_m=c_m, c_r=c_r)
    log("[init]", r)
    mine.close(); ref.close()
    return r


def s_launch():
    """Single half-steps from the reference's init state: refine (deterministic in the reference) and
    spatial propagation (vs the snapshot oracle bit-exactly, vs the as-is oracle statistically)."""
    scene = pkg.scene.make_scene("small")
    params, mine, refs = pc.make_engines(pkg, scene, variants=("asis", "snapshot"))
    asis, snap = refs["asis"], refs["snapshot"]
    asis.init_planes(SEED)
    n0, c0 = asis.download(rb.F_NORM4), asis.download(rb.F_COST)
    res = {}

    def put(eng, fld_n, fld_c):
        eng.upload(fld_n, n0); eng.upload(fld_c, c0)

    # --- refine
    for kind, nm in ((L.BLACK_REFINE, "black_refine"), (L.RED_REFINE, "red_refine")):
        put(asis, rb.F_NORM4, rb.F_COST); asis.launch(kind, SEED + 1)
        mine.load_planes(n0, c0); mine.launch(kind, SEED + 1)
        a_n, a_c = asis.download(rb.F_NORM4), asis.download(rb.F_COST)
        m_n, m_c = mine.download(L.F_NORM4), mine.download(L.F_COST)
        res[nm] = dict(norm4_bit_exact=pc.frac_bit_exact(m_n, a_n), cost_bit_exact=pc.frac_bit_exact(m_c, a_c),
                       changed=float((~pc.bits_equal(a_c, c0)).mean()),
                       beview_eq=float((mine.download(L.F_BEVIEW) == asis.download(rb.F_BEVIEW)).mean()),
                       ratio_bit_exact=pc.frac_bit_exact(mine.download(L.F_RATIO), asis.download(rb.F_RATIO)))
        log("[launch]", nm, res[nm])
        if nm == "black_refine":
            np.savez_compressed(os.path.join(OUT, "refine_small.npz"), n0=n0, c0=c0, a_n=a_n, a_c=a_c, m_n=m_n, m_c=m_c)
    # --- spatial propagation
    for kind, nm in ((L.BLACK_SPATIAL, "black_spatial"), (L.RED_SPATIAL, "red_spatial")):
        put(snap, rb.F_NORM4, rb.F_COST); snap.launch(kind)
        put(asis, rb.F_NORM4, rb.F_COST); asis.launch(kind)
        mine.load_planes(n0, c0); mine.launch(kind)
        s_n, s_c = snap.download(rb.F_NORM4), snap.download(rb.F_COST)
        a_n, a_c = asis.download(rb.F_NORM4), asis.download(rb.F_COST)
        m_n, m_c = mine.download(L.F_NORM4), mine.download(L.F_COST)
        res[nm] = dict(vs_snapshot_norm4=pc.frac_bit_exact(m_n, s_n), vs_snapshot_cost=pc.frac_bit_exact(m_c, s_c),
                       vs_asis_norm4=pc.frac_bit_exact(m_n, a_n), snapshot_vs_asis_norm4=pc.frac_bit_exact(s_n, a_n),
                       changed=float((~pc.bits_equal(s_c, c0)).mean()))
        log("[launch]", nm, res[nm])
        if nm == "black_spatial":
            np.savez_compressed(os.path.join(OUT, "spatial_small.npz"), n0=n0, c0=c0, s_n=s_n, s_c=s_c, m_n=m_n, m_c=m_c, a_n=a_n, a_c=a_c)
    for e in (mine, asis, snap):
        e.close()
    return res


def s_full():
    """Whole north-star sequence: ours vs snapshot oracle vs as-is oracle (twice: noise floor)."""
    res = {}
    for cfg in ("small",):
        scene = pkg.scene.make_scene(cfg)
        params, mine, refs = pc.make_engines(pkg, scene, variants=("asis", "snapshot"))
        asis, snap = refs["asis"], refs["snapshot"]
        ms_m = mine.depthmap(SEED)
        o_m = mine.download(L.F_NORM4); conf_m = mine.download(L.F_CONFID); lr_m = mine.download(L.F_LRDIFF)
        ms_s = snap.depthmap(SEED); o_s = snap.download(rb.F_NORM4); conf_s = snap.download(rb.F_CONFID); lr_s = snap.download(rb.F_LRDIFF)
        ms_a = asis.depthmap(SEED); o_a = asis.download(rb.F_NORM4)
        ms_a2 = asis.depthmap(SEED); o_a2 = asis.download(rb.F_NORM4)
        res[cfg] = dict(ms_mine=ms_m, ms_snapshot=ms_s, ms_asis=ms_a, ms_asis2=ms_a2,
                        mine_vs_snapshot=pc.output_agreement(o_m, o_s), mine_vs_asis=pc.output_agreement(o_m, o_a),
                        snapshot_vs_asis=pc.output_agreement(o_s, o_a), asis_vs_asis=pc.output_agreement(o_a2, o_a),
                        gt_mine=pc.gt_agreement(o_m, scene), gt_asis=pc.gt_agreement(o_a, scene),
                        confid_max_abs=float(np.abs(conf_m - conf_s).max()), lrdiff_bit_exact=pc.frac_bit_exact(lr_m, lr_s),
                        confid_bit_exact=pc.frac_bit_exact(conf_m, conf_s), launches=mine.launch_count())
        log("[full]", cfg, json.dumps(res[cfg], default=float))
        np.savez_compressed(os.path.join(OUT, f"full_{cfg}.npz"), o_m=o_m, o_s=o_s, o_a=o_a, o_a2=o_a2, gt=scene["gt_depth"], labels=scene["labels"])
        # unfused path must be identical to the fused one
        os.environ["TSAR_B200_UNFUSED"] = "1"
        params, mine2, _ = pc.make_engines(pkg, scene, variants=())
        del os.environ["TSAR_B200_UNFUSED"]
        mine2.depthmap(SEED)
        res[cfg]["fused_vs_unfused_bit_exact"] = pc.frac_bit_exact(mine2.download(L.F_NORM4), o_m)
        log("[full] fused_vs_unfused", res[cfg]["fused_vs_unfused_bit_exact"])
        for e in (mine, mine2, asis, snap):
            e.close()
    return res


def s_glue():
    scene = pkg.scene.make_scene("small")
    params, mine, refs = pc.make_engines(pkg, scene, variants=("snapshot",))
    ref = refs["snapshot"]
    ref.init_planes(SEED); ref.iterate(2, SEED)
    n0, c0 = ref.download(rb.F_NORM4), ref.download(rb.F_COST)
    bv, ratio = ref.download(rb.F_BEVIEW), ref.download(rb.F_RATIO)
    mine.load_planes(n0, c0); mine.upload(L.F_BEVIEW, bv); mine.upload(L.F_RATIO, ratio)
    res = {}
    ref.lrdiff(); mine.lrdiff()
    res["lrdiff"] = dict(bit_exact=pc.frac_bit_exact(mine.download(L.F_LRDIFF), ref.download(rb.F_LRDIFF)),
                         max_abs=float(np.abs(mine.download(L.F_LRDIFF) - ref.download(rb.F_LRDIFF)).max()))
    mine.upload(L.F_LRDIFF, ref.download(rb.F_LRDIFF))
    ref.getview(); mine.getview()
    res["getview"] = dict(confid=pc.frac_bit_exact(mine.download(L.F_CONFID), ref.download(rb.F_CONFID)),
                          depth=pc.frac_bit_exact(mine.download(L.F_DEPTH), ref.download(rb.F_DEPTH)))
    # regions / depth completion
    for e in (mine, ref):
        e.set_regions(scene["region_text"], scene["region_norm4"])
    mine.upload(L.F_CANNY, scene["canny"]); ref.upload(rb.F_CANNY, scene["canny"])
    ref.update_scale_2(); mine.update_scale_2()
    res["update_scale_2"] = dict(fakedepth=pc.frac_bit_exact(mine.download(L.F_FAKEDEPTH), ref.download(rb.F_FAKEDEPTH)))
    ref.update_scale(); mine.update_scale()
    res["update_scale"] = dict(norm4=pc.frac_bit_exact(mine.download(L.F_NORM4), ref.download(rb.F_NORM4)),
                               cost=pc.frac_bit_exact(mine.download(L.F_COST), ref.download(rb.F_COST)),
                               scale=pc.frac_bit_exact(mine.download(L.F_SCALE), ref.download(rb.F_SCALE)),
                               depth=pc.frac_bit_exact(mine.download(L.F_DEPTH), ref.download(rb.F_DEPTH)))
    ref.compute_disp(); mine.compute_disp()
    res["compute_disp"] = dict(norm4=pc.frac_bit_exact(mine.download(L.F_NORM4), ref.download(rb.F_NORM4)))
    # get_disp on imported world normals + disparities
    wn = ref.download(rb.F_NORM4).copy()
    dsp = np.where(wn[..., 3] > 0, scene["cam_f"] / np.maximum(wn[..., 3], 1e-6), 1.0).astype(np.float32)
    for e, f_n, f_d in ((mine, L.F_NORM4, L.F_DEPTH), (ref, rb.F_NORM4, rb.F_DEPTH)):
        e.upload(f_n, wn); e.upload(f_d, dsp)
    ref.get_disp(); mine.get_disp()
    res["get_disp"] = dict(norm4=pc.frac_bit_exact(mine.download(L.F_NORM4), ref.download(rb.F_NORM4)))
    log("[glue]", json.dumps(res, default=float))
    mine.close(); ref.close()
    return res


def s_time_c1():
    scene = pkg.scene.make_scene("C1")
    params, mine, refs = pc.make_engines(pkg, scene, variants=("asis",))
    ref = refs["asis"]
    t_m = [mine.depthmap(SEED) for _ in range(3)]
    t_r = [ref.depthmap(SEED, prefetch=(i > 0)) for i in range(3)]
    o_m, o_r = mine.download(L.F_NORM4), ref.download(rb.F_NORM4)
    n_ev = mine.eval_count(params.iterations)
    r = dict(ms_mine=t_m, ms_ref=t_r, evals=n_ev, gevals_per_s_mine=n_ev / (min(t_m) * 1e-3) / 1e9,
             gevals_per_s_ref=n_ev / (min(t_r) * 1e-3) / 1e9, agreement=pc.output_agreement(o_m, o_r),
             gt_mine=pc.gt_agreement(o_m, scene), gt_ref=pc.gt_agreement(o_r, scene))
    log("[time_c1]", json.dumps(r, default=float))
    mine.close(); ref.close()
    return r


def s_slic():
    res = {}
    for cfg, size in (("small", 20), ("C1", 20)):
        scene = pkg.scene.make_scene(cfg, with_colour=True)
        bgrx = np.concatenate([scene["bgr"], np.zeros(scene["bgr"].shape[:2] + (1,), np.uint8)], axis=-1)
        eng = pkg.DepthmapEngine(0)
        for rep in range(2):
            t = time.time(); mine = eng.slic(bgrx, spixel_size=size); t_m = time.time() - t
        ref, ms_r = rb.ref_slic(bgrx, spixel_size=size)
        ref, ms_r = rb.ref_slic(bgrx, spixel_size=size)
        full = eng.slic(bgrx, spixel_size=size, correct_reduction=True)
        res[cfg] = dict(labels_equal=float((mine == ref).mean()), n_labels=int(len(np.unique(ref))), wall_ms_mine=t_m * 1e3,
                        ms_ref=ms_r, full_vs_parity=float((full == mine).mean()))
        np.savez_compressed(os.path.join(OUT, f"slic_{cfg}.npz"), mine=mine, ref=ref, full=full)
        log("[slic]", cfg, res[cfg])
        eng.close()
    return res


STAGES = dict(slic=s_slic, peaks=s_peaks, texcal=s_texcal, eval=s_eval, init=s_init, launch=s_launch, full=s_full, glue=s_glue,
              time_c1=s_time_c1)

if __name__ == "__main__":
    todo = sys.argv[1:] or list(STAGES)
    for name in todo:
        log(f"===== {name}")
        stage(STAGES[name])
    log(json.dumps(RES, indent=1, default=float)[:6000])
