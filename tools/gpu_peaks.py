"""Measured issue-rate peaks of this B200 for the PatchMatch roofline (MEASURED_PEAKS.json has only HBM and bf16):
FP32 FFMA, MUFU and bilinear texture fetch, with the SM clock sampled while the microbenchmarks run.

    python tools/gpu_peaks.py  ->  gpurun_out/r02_fp32_peaks.json   (copied to profiles/r02_fp32_peaks.json)
"""
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as ge  # noqa: E402

pkg = ge.load_package()
scene = pkg.scene.make_scene("small")
from tsar_mvs_b200.engine import cameras_to_struct  # noqa: E402

eng = pkg.DepthmapEngine(0)
eng.set_views(scene["images"], cameras_to_struct(scene["cams"]), scene["subset"], cam_f=scene["cam_f"])
eng.set_params(pkg.make_params(min_disparity=scene["min_disparity"], max_disparity=scene["max_disparity"]))
rows = []
stop = False


def poll():
    q = "clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"
    while not stop:
        try:
            out = subprocess.run(["nvidia-smi", "--id=0", f"--query-gpu={q}", "--format=csv,noheader,nounits"], capture_output=True, text=True, timeout=5).stdout
            rows.append([c.strip() for c in out.strip().split(",")])
        except Exception:
            pass
        time.sleep(0.05)


th = threading.Thread(target=poll, daemon=True)
th.start()
runs = []
t_end = time.time() + 4.0
while time.time() < t_end:
    runs.append(eng.peaks())
stop = True
th.join()
eng.close()
best = [max(r[k] for r in runs) for k in range(3)]
sm = sorted(float(r[0]) for r in rows if r and r[0].replace(".", "").isdigit())
name = subprocess.run(["nvidia-smi", "--id=0", "--query-gpu=name,driver_version", "--format=csv,noheader"], capture_output=True, text=True).stdout.strip()
res = {
    "gpu": name, "how": "tsar_dbg_peaks (csrc/debug_kernels.cuh): 8 independent FFMA chains / 8 MUFU.RSQ chains / 4 bilinear fp32 fetches per "
                        "thread, 148*8 CTAs of 256 threads, CUDA events, best of all repetitions within 4 s",
    "fp32_ffma_tflops": best[0], "mufu_gops": best[1], "tex_bilinear_gsamples": best[2],
    "nominal": {"fp32_ffma_tflops": 148 * 128 * 2 * 1.965e9 / 1e12, "mufu_gops": 148 * 16 * 1.965, "tex_bilinear_gsamples": 148 * 4 * 1.965,
                "note": "148 SMs x (128 FFMA lanes x 2 flop | 16 MUFU lanes | 4 bilinear samples) per clock at 1965 MHz"},
    "repetitions": len(runs),
    "clocks_under_load": {"samples": len(sm), "sm_mhz_median": sm[len(sm) // 2] if sm else None, "sm_mhz_min": sm[0] if sm else None,
                          "sm_mhz_max": sm[-1] if sm else None, "sm_max_mhz": float(rows[0][1]) if rows else None,
                          "power_w_max": max(float(r[2]) for r in rows) if rows else None,
                          "any_slowdown_reason_active": any("Active" == c for r in rows for c in r[3:]) if rows else None},
}
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
json.dump(res, open(os.path.join(ROOT, "gpurun_out", "r02_fp32_peaks.json"), "w"), indent=1)
print(json.dumps(res, indent=1))
