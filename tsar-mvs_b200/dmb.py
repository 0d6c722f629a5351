"""`.dmb` depth / normal maps as the reference writes them (fileIoUtils.h:333-381): int32 type (1 = float),
int32 height, int32 width, int32 channels, then h*w*channels float32 row-major.  Fusion.exe and the TSAR
scripts consume TSAR_disp.dmb (1 channel) and TSAR_normals.dmb (3 channels) (main.cpp:1807-1861)."""
import numpy as np


def write_dmb(path, arr):
    a = np.ascontiguousarray(arr, np.float32)
    if a.ndim == 2:
        a = a[..., None]
    h, w, nb = a.shape
    with open(path, "wb") as f:
        np.array([1, h, w, nb], np.int32).tofile(f)
        a.tofile(f)


def read_dmb(path):
    with open(path, "rb") as f:
        t, h, w, nb = np.fromfile(f, np.int32, 4)
        if t != 1:
            raise ValueError(f"{path}: unsupported dmb type {t} (only float32 = 1 is written by the reference)")
        data = np.fromfile(f, np.float32, int(h) * int(w) * int(nb))
    if data.size != int(h) * int(w) * int(nb):
        raise ValueError(f"{path}: truncated dmb ({data.size} of {int(h) * int(w) * int(nb)} values)")
    a = data.reshape(int(h), int(w), int(nb))
    return a[..., 0] if nb == 1 else a


def write_outputs(folder, out_norm4):
    """TSAR_disp.dmb + TSAR_normals.dmb from the gipuma_compute_disp layout (xyz = world normal, w = depth)."""
    import os
    os.makedirs(folder, exist_ok=True)
    write_dmb(os.path.join(folder, "TSAR_disp.dmb"), out_norm4[..., 3])
    write_dmb(os.path.join(folder, "TSAR_normals.dmb"), out_norm4[..., :3])


def write_model_ply(path, depth, normals, gray, K, R, t):
    """TSAR_model.ply as storePlyFileBinary writes it (displayUtils.h:77-158, call site main.cpp:1836-1843): binary
    little-endian, one vertex per pixel -- float x y z (get3Dpoint with the UNtransformed camera P = K[R|t],
    cameraGeometryUtils.h:53-65: X = M^-1 (depth*(x,y,1) - p4)), float nx ny nz, uchar grey x3 -- pixels visited
    column by column (x outer, y inner; the reference's OpenMP loop writes them in a thread-dependent order, this is
    its sequential order).  Non-finite points become (0,0,0)."""
    depth = np.asarray(depth, np.float32)
    H, W = depth.shape
    P = (np.asarray(K, np.float64) @ np.concatenate([np.asarray(R, np.float64), np.asarray(t, np.float64).reshape(3, 1)], 1)).astype(np.float32)
    Minv = np.linalg.inv(P[:, :3].astype(np.float64)).astype(np.float32)
    xs, ys = np.meshgrid(np.arange(W, dtype=np.float32), np.arange(H, dtype=np.float32), indexing="ij")   # [x][y]
    d = depth.T
    rhs = np.stack([d * xs - P[0, 3], d * ys - P[1, 3], d - P[2, 3]], -1).astype(np.float32)
    with np.errstate(invalid="ignore", over="ignore"):
        X = (rhs @ Minv.T).astype(np.float32)
    bad = ~np.isfinite(X).all(-1)
    X[bad] = 0.0
    rec = np.zeros(W * H, dtype=[("p", "<f4", 3), ("n", "<f4", 3), ("c", "u1", 3)])
    rec["p"] = X.reshape(-1, 3)
    rec["n"] = np.asarray(normals, np.float32).transpose(1, 0, 2).reshape(-1, 3)
    rec["c"] = np.asarray(gray).astype(np.uint8).T.reshape(-1, 1)
    with open(path, "wb") as f:
        f.write(("ply\nformat binary_little_endian 1.0\n" + f"element vertex {H * W}\n" +
                 "property float x\nproperty float y\nproperty float z\nproperty float nx\nproperty float ny\nproperty float nz\n"
                 "property uchar red\nproperty uchar green\nproperty uchar blue\nend_header\n").encode())
        rec.tofile(f)


def read_model_ply(path):
    """Reader for the file above (tests / hand-over to a fusion stage): returns (points [n][3], normals, grey)."""
    with open(path, "rb") as f:
        n = None
        while True:
            line = f.readline().decode().strip()
            if line.startswith("element vertex"):
                n = int(line.split()[-1])
            if line == "end_header":
                break
        rec = np.fromfile(f, dtype=[("p", "<f4", 3), ("n", "<f4", 3), ("c", "u1", 3)], count=n)
    return rec["p"], rec["n"], rec["c"]
