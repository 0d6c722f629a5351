"""Command line of the reference (`gipuma <ref image> <other images...> flags`, main.cpp:708-1009, run scripts
scripts/*.sh) on top of libtsar_b200.so.

    python -m tsar_cli 00000000.jpg 00000001.jpg ... -mslp_folder data/TRAIN/pipes/ -images_folder data/TRAIN/pipes/images/ \
        -krt_file x -output_folder results/pipes/ -no_display --cam_scale=1 --iterations=8 --blocksize=11 \
        --cost_gamma=10 --cost_comb=best_n --n_best=1 --min_angle= --max_angle=

Kept from the reference: positional arguments are image names, the first one is the reference view; `--key=value`
numeric flags, `-key value` path flags, `-color_processing` / `-view_selection` booleans; unknown flags only warn
(main.cpp:941-945) -- the scripts pass an unknown `-no_display` and empty `--min_angle=`/`--max_angle=`;
`-krt_file` merely selects the MVSNet-style reader, its value is ignored (fileIoUtils.h:111-118); cameras come from
`<mslp>/cams/<8-digit stem>_cam.txt`, the source views from `<mslp>/pair.txt` (main.cpp:1351-1376, camera id =
atoi(name[4:8])); results go to `<mslp>/APD/<stem>/TSAR_disp.dmb` and `TSAR_normals.dmb` (main.cpp:1807-1861).

Flows:
  * default (north-star): random initialisation + red/black PatchMatch + L/R check + confidence + output layout;
  * `-import_apd`: the shipped flow -- planes imported from `<mslp>/APD/<stem>/depths_geom.dmb` + `normals.dmb`
    (main.cpp:1459-1490) then gipuma_get_disp / getview / output.
Textureless-region completion: the weak-texture detector of main.cpp:365-596 (texture.py; cv2 for pyrDown /
HoughLinesP / line as the reference uses OpenCV) labels the reference view; regions it flags get a plane from
tsar_fit_region_planes and are completed by update_scale(_2).  `-no_weak_texture` skips it; `-regions_file labels.npy`
(+ `-regions_text text.npy`) supplies a precomputed label map instead.
`-write_ply` also writes `TSAR_model.ply` (displayUtils.h:77-158).
Extra: `--synthetic=<C1|C2|small|tiny>` first writes a synthetic dataset in the reference's folder layout.

`-all_views` replaces the per-view process loop of the run scripts (scripts/pipes.sh:30-49): every image of the
dataset becomes the reference view once.  All images are decoded and uploaded ONCE and stay resident in HBM; each
reference view binds itself + its pair.txt neighbours from that pool (device-to-device), two contexts on two streams
pipeline view r+1 behind view r, results are written by the worker threads.  Under torchrun the views are sharded
round-robin over the ranks (shard.py), one GPU per rank, no collective.
"""
import os
import sys

import numpy as np

from . import dmb, scene
from .engine import DepthmapEngine, make_params

NUMERIC = {"blocksize", "iterations", "n_best", "cost_gamma", "depth_min", "depth_max", "max_views", "min_angle", "max_angle",
           "cam_scale", "cost_tau_color", "cost_tau_gradient", "cost_alpha", "disp_tol", "normal_tol", "census_epsilon",
           "self_similarity_n", "good_factor", "num_img_processed", "seed", "synthetic", "device"}
PATHS = {"images_folder", "mslp_folder", "krt_file", "output_folder", "p_folder", "camera_folder", "calib_file", "pmvs_folder",
         "bounding_folder", "regions_file", "regions_text", "regions_size"}
BOOLS = {"color_processing", "view_selection", "import_apd", "no_display", "all_views", "no_weak_texture", "write_ply"}


def parse_args(argv):
    """getParametersFromCommandLine (main.cpp:708-1009), tolerant in the same places."""
    opt = dict(images=[], blocksize=19, iterations=8, n_best=2, cost_comb=1, cam_scale=1.0, depth_min=-1.0, depth_max=-1.0,
               seed=20240601, device=0, images_folder="", mslp_folder="", output_folder="", synthetic=None,
               color_processing=False, import_apd=False, all_views=False, no_weak_texture=False, write_ply=False, regions_file=None, regions_text=None, regions_size=None)
    i = 0
    while i < len(argv):
        a = argv[i]
        if a.startswith("--"):
            key, _, val = a[2:].partition("=")
            if key == "cost_comb":
                opt["cost_comb"] = {"all": 0, "best_n": 1, "angle": 2, "good": 3}.get(val, opt["cost_comb"])
            elif key == "synthetic":
                opt["synthetic"] = val
            elif key in NUMERIC:
                if val != "":  # the scripts pass `--min_angle=` with an empty value
                    opt[key] = float(val) if "." in val or key in ("cam_scale", "depth_min", "depth_max") else int(val)
            else:
                print(f"[tsar_cli] warning: unknown option {a}", file=sys.stderr)
        elif a.startswith("-") and len(a) > 1 and not a[1].isdigit():
            key = a[1:]
            if key in BOOLS:
                opt[key] = True
            elif key in PATHS:
                i += 1
                opt[key] = argv[i] if i < len(argv) else ""
            else:
                print(f"[tsar_cli] warning: unknown option {a}", file=sys.stderr)
        else:
            opt["images"].append(a)
        i += 1
    return opt


# ---- dataset I/O -------------------------------------------------------------------------------------------
def read_cam_txt(path):
    """MVSNet-style camera file as readKRtFileMiddlebury reads it (fileIoUtils.h:111-163): 'extrinsic', 4x4 [R|t],
    'intrinsic', 3x3 K, then depth_min interval depth_num depth_max."""
    tok = open(path).read().split()
    assert tok[0] == "extrinsic", path
    E = np.array(tok[1:17], float).reshape(4, 4)
    assert tok[17] == "intrinsic", path
    K = np.array(tok[18:27], float).reshape(3, 3)
    tail = [float(v) for v in tok[27:31]]
    depth_min, depth_max = tail[0], tail[3]
    return K, E[:3, :3], E[:3, 3], depth_min, depth_max


def write_cam_txt(path, K, R, t, depth_min, depth_max, interval=0.0, depth_num=0):
    E = np.eye(4)
    E[:3, :3], E[:3, 3] = R, t
    with open(path, "w") as f:
        f.write("extrinsic\n" + "\n".join(" ".join(repr(float(v)) for v in row) for row in E) + "\n\nintrinsic\n")
        f.write("\n".join(" ".join(repr(float(v)) for v in row) for row in np.asarray(K, float)) + "\n\n")
        f.write(f"{depth_min!r} {interval!r} {depth_num} {depth_max!r}\n")


def read_pair_subset(path, camera_id):
    """Source views of camera `camera_id` as indices into the command-line image list (main.cpp:1351-1376): the list is
    [reference, all other images in order], so image k maps to k+1 if k < camera_id else k."""
    lines = open(path).read().split("\n")
    line = lines[2 * camera_id + 2].split()  # 1 header line, then per camera: id line + neighbour line
    n = int(line[0])
    subset = []
    for j in range(n):
        k = int(line[1 + 2 * j])
        subset.append(k if k > camera_id else k + 1)
    return subset


def write_pair_txt(path, neighbours):
    with open(path, "w") as f:
        f.write(f"{len(neighbours)}\n")
        for i, nb in enumerate(neighbours):
            f.write(f"{i}\n{len(nb)} " + " ".join(f"{k} 1.0" for k in nb) + "\n")


def _imread_gray(path, colour=False):
    """Grey image (cv::imread IMREAD_GRAYSCALE, main.cpp:1302).  With -color_processing the reference uploads the
    BGRA image instead and its kernels sample the first component (tex2D<float>, gipuma.cu:247,262,265): blue."""
    if path.endswith(".npy"):
        return np.load(path).astype(np.float32)
    import cv2  # image decoding only
    im = cv2.imread(path, cv2.IMREAD_COLOR if colour else cv2.IMREAD_GRAYSCALE)
    if im is None:
        raise FileNotFoundError(path)
    return (im[..., 0] if colour else im).astype(np.float32)


def write_synthetic_dataset(cfg_name, root):
    """Writes a synthetic scene in the reference's folder layout: <root>/images/%08d.png, <root>/cams/%08d_cam.txt,
    <root>/pair.txt.  Returns (scene dict, list of image names)."""
    import cv2
    sc = scene.make_scene(cfg_name)
    os.makedirs(os.path.join(root, "images"), exist_ok=True)
    os.makedirs(os.path.join(root, "cams"), exist_ok=True)
    names = []
    n = len(sc["images"])
    for i, (im, cam) in enumerate(zip(sc["images"], sc["cams"])):
        name = f"{i:08d}.png"
        cv2.imwrite(os.path.join(root, "images", name), im.astype(np.uint8))
        R, C = cam["_R_world"], cam["_C_world"]
        write_cam_txt(os.path.join(root, "cams", f"{i:08d}_cam.txt"), cam["K"], R, -R @ C, cam["depthMin"], cam["depthMax"])
        names.append(name)
    write_pair_txt(os.path.join(root, "pair.txt"), [[k for k in range(n) if k != i] for i in range(n)])
    return sc, names


# ---- the driver ---------------------------------------------------------------------------------------------
def run(argv):
    opt = parse_args(argv)
    mslp = opt["mslp_folder"]
    if opt["synthetic"]:
        mslp = mslp or "./synthetic_" + opt["synthetic"] + "/"
        _, names = write_synthetic_dataset(opt["synthetic"], mslp)
        opt["images"] = opt["images"] or names
        opt["images_folder"] = os.path.join(mslp, "images/")
    if opt["all_views"]:
        return run_all_views(opt, mslp)
    if not opt["images"]:
        print(__doc__)
        return 2
    names = opt["images"]
    stem = names[0][:8]                         # main.cpp:1462: numind = imgname.substr(0, 8)
    camera_id = int(names[0][4:8])              # main.cpp:1347-1349: atoi(name.substr(4, 8)) (sic: 4 characters)
    images = [_imread_gray(os.path.join(opt["images_folder"], n), opt["color_processing"]) for n in names]
    Ks, Rs, ts, dmin, dmax = [], [], [], None, None
    for i, n in enumerate(names):
        K, R, t, a, b = read_cam_txt(os.path.join(mslp, "cams", f"{n[:8]}_cam.txt"))
        K = K.copy()
        K[:2] /= opt["cam_scale"]               # scaleK (cameraGeometryUtils.h:141-151)
        Ks.append(K); Rs.append(R); ts.append(t)
        if i == 0:
            dmin, dmax = a, b                   # depth range of the reference view (fileIoUtils.h:145-153)
    if opt["depth_min"] > 0:
        dmin = opt["depth_min"]
    if opt["depth_max"] > 0:
        dmax = opt["depth_max"]
    cams = scene.cameras_from_krt(Ks, Rs, ts, dmin, dmax)
    subset = read_pair_subset(os.path.join(mslp, "pair.txt"), camera_id)
    subset = [s for s in subset if s < len(names)]
    f = float(np.float32(cams[0]["f"]))
    params = make_params(box=opt["blocksize"], iterations=opt["iterations"], n_best=opt["n_best"], cost_comb=opt["cost_comb"],
                         min_disparity=float(np.float32(f / np.float32(dmax))), max_disparity=float(np.float32(f / np.float32(dmin))))
    eng = DepthmapEngine(int(opt["device"]))
    eng.set_views(images, cams, subset, cam_f=f)
    eng.set_params(params)
    from . import _lib as L
    out_dir = os.path.join(mslp, "APD", stem)
    H, W = images[0].shape
    if opt["import_apd"]:                       # shipped flow, main.cpp:1459-1490
        depth = dmb.read_dmb(os.path.join(out_dir, "depths_geom.dmb"))
        normal = dmb.read_dmb(os.path.join(out_dir, "normals.dmb"))
        n4 = np.concatenate([normal, np.zeros((H, W, 1), np.float32)], axis=-1)
        eng.upload(L.F_NORM4, n4)
        eng.upload(L.F_COST, np.ones((H, W), np.float32))
        eng.upload(L.F_DEPTH, (np.float32(f) / depth).astype(np.float32))
        eng.get_disp()
    else:
        eng.init_planes(int(opt["seed"]))
        eng.iterate(opt["iterations"], int(opt["seed"]))
        eng.lrdiff()
    eng.getview()
    confid = eng.download(L.F_CONFID)
    text = size = None
    if opt["regions_file"]:
        labels = np.load(opt["regions_file"]).astype(np.float32)
        text = np.load(opt["regions_text"]).astype(np.float32)
        size = np.load(opt["regions_size"]).astype(np.float32) if opt["regions_size"] else \
            np.array([(labels == r).sum() / 16.0 for r in range(len(text))], np.float32)
        eng.upload(L.F_CANNY, labels)
    elif not opt["no_weak_texture"]:            # texture() runs first in the reference's main (main.cpp:1896)
        from . import texture
        det = texture.detect(images[0].astype(np.uint8))
        n_weak = int((det["text"] == -1).sum())
        print(f"[tsar_cli] weak-texture detector: {len(det['text']) - 1} regions, {n_weak} weakly textured")
        if n_weak:
            text, size = det["text"], det["size"]
            eng.set_labels_quarter(det["labels_q"])
    if text is not None:
        eng.upload(L.F_SCALE, (confid > 0.8).astype(np.float32))       # reliable pixels (stands in for APD's weak.png)
        rng = np.random.RandomState(int(opt["seed"]) & 0x7fffffff)
        rnd = rng.randint(0, 2 ** 31 - 1, size=(len(text), eng.lib.tsar_ransac_rand_per_region())).astype(np.uint32)
        planes = eng.fit_region_planes(text, size, rnd, np.tile(np.array([0, 0, 1, -1], np.float32), (len(text), 1)))
        eng.set_regions(text, planes)
        eng.update_scale_2()
        eng.update_scale()
    eng.compute_disp()
    out = eng.download(L.F_NORM4)
    dmb.write_outputs(out_dir, out)
    dmb.write_dmb(os.path.join(out_dir, "TSAR_confidence.dmb"), confid)   # computed but never written by the reference
    if opt["write_ply"]:                        # main.cpp:1836-1843 (always on there; 27 B per pixel, so opt-in here)
        dmb.write_model_ply(os.path.join(out_dir, "TSAR_model.ply"), out[..., 3], out[..., :3], images[0], Ks[0], Rs[0], ts[0])
    print(f"[tsar_cli] {stem}: {W}x{H}, {len(subset)} source views -> {out_dir}/TSAR_disp.dmb, TSAR_normals.dmb")
    eng.close()
    return 0


def read_pair_neighbours(path, camera_id):
    """Camera ids of the source views of `camera_id` in pair.txt order (main.cpp:1351-1376)."""
    lines = open(path).read().split("\n")
    line = lines[2 * camera_id + 2].split()
    return [int(line[1 + 2 * j]) for j in range(int(line[0]))]


def run_all_views(opt, mslp):
    """Persistent multi-view driver (SURVEY section 8 rows e/f4): see the module docstring."""
    import concurrent.futures as cf
    import time

    import torch

    from . import _lib as L
    from .shard import views_for_rank
    rank, world = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))
    ndev = torch.cuda.device_count()
    if ndev == 0:
        raise RuntimeError("tsar_cli -all_views needs a CUDA device (no CPU path)")
    dev = int(os.environ.get("LOCAL_RANK", opt["device"])) % ndev
    folder = opt["images_folder"] or os.path.join(mslp, "images/")
    names = opt["images"] or sorted(n for n in os.listdir(folder) if n.lower().endswith((".jpg", ".jpeg", ".png", ".npy")))
    ids = [int(n[4:8]) for n in names]                          # camera id of each image (main.cpp:1347-1349)
    by_id = {c: k for k, c in enumerate(ids)}
    krt = [read_cam_txt(os.path.join(mslp, "cams", f"{n[:8]}_cam.txt")) for n in names]
    t0 = time.perf_counter()
    pool = [torch.from_numpy(_imread_gray(os.path.join(folder, n), opt["color_processing"])).to(f"cuda:{dev}") for n in names]   # resident, once
    H, W = pool[0].shape
    mine = views_for_rank(len(names), rank, world)
    lanes = []
    for _ in range(min(2, max(1, len(mine)))):
        st = torch.cuda.Stream(device=dev)
        lanes.append(dict(eng=DepthmapEngine(dev, stream=st.cuda_stream), stream=st))
    done = []

    def process(k, lane):
        eng = lane["eng"]
        r = mine[k]
        nb = [by_id[c] for c in read_pair_neighbours(os.path.join(mslp, "pair.txt"), ids[r]) if c in by_id]
        sel = [r] + nb
        Ks = []
        for i in sel:
            K = krt[i][0].copy()
            K[:2] /= opt["cam_scale"]
            Ks.append(K)
        dmin = opt["depth_min"] if opt["depth_min"] > 0 else krt[r][3]
        dmax = opt["depth_max"] if opt["depth_max"] > 0 else krt[r][4]
        cams = scene.cameras_from_krt(Ks, [krt[i][1] for i in sel], [krt[i][2] for i in sel], dmin, dmax)
        f = float(np.float32(cams[0]["f"]))
        params = make_params(box=opt["blocksize"], iterations=opt["iterations"], n_best=opt["n_best"], cost_comb=opt["cost_comb"],
                             min_disparity=float(np.float32(f / np.float32(dmax))), max_disparity=float(np.float32(f / np.float32(dmin))))
        from .engine import cameras_to_struct
        eng.set_views_device([pool[i].data_ptr() for i in sel], W, H, cameras_to_struct(cams), list(range(1, len(sel))), cam_f=f)
        eng.set_params(params)
        eng.init_planes(int(opt["seed"]) + r)
        eng.iterate(opt["iterations"], int(opt["seed"]) + r)
        eng.lrdiff(); eng.getview(); eng.compute_disp()
        out, confid = eng.download(L.F_NORM4), eng.download(L.F_CONFID)
        out_dir = os.path.join(mslp, "APD", names[r][:8])
        dmb.write_outputs(out_dir, out)
        dmb.write_dmb(os.path.join(out_dir, "TSAR_confidence.dmb"), confid)
        done.append(r)
        return r

    import itertools
    import threading
    ticket, ticket_lock = itertools.count(), threading.Lock()

    def lane_worker(j):     # one host thread per lane (a context is never shared); views are taken from a shared queue,
        while True:         # so lanes stay balanced when the views' neighbour counts differ
            with ticket_lock:
                k = next(ticket)
            if k >= len(mine):
                return
            process(k, lanes[j])

    with cf.ThreadPoolExecutor(len(lanes)) as ex:
        for fu in [ex.submit(lane_worker, j) for j in range(len(lanes))]:
            fu.result()
    torch.cuda.synchronize(dev)
    dt = time.perf_counter() - t0
    for lane in lanes:
        lane["eng"].close()
    print(f"[tsar_cli] rank {rank}/{world}: {len(done)} of {len(names)} reference views ({W}x{H}) in {dt:.2f} s "
          f"({len(done) / dt:.2f} depthmaps/s incl. decoding and .dmb writing) -> {os.path.join(mslp, 'APD')}")
    return 0


if __name__ == "__main__":
    sys.exit(run(sys.argv[1:]))
