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
