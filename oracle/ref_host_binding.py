"""ctypes binding of oracle/_ref/libtsar_ref_host.so -- TEST INFRASTRUCTURE ONLY.

The library is the reference's own HOST code for SURVEY section 8 rows f1 / f3 (main.cpp:147-164, 214-362, 365-596,
1520-1730), cut out of the checkout and compiled by oracle/build_ref.sh around oracle/ref_host_driver.cpp.  The OpenCV
library calls inside texture() are served by cv2 through the callbacks below; rand() reads the stream given to
set_rand().  Only tests/ may import this."""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB = os.path.join(_HERE, "_ref", "libtsar_ref_host.so")
_lib = None
_keep = []


def available():
    return os.path.exists(LIB)


def _u8(ptr, rows, cols, ch=1):
    n = rows * cols * ch
    a = np.ctypeslib.as_array(C.cast(ptr, C.POINTER(C.c_ubyte)), shape=(n,))
    return a.reshape((rows, cols) if ch == 1 else (rows, cols, ch))


def load():
    """Loads the library and installs the cv2-backed callbacks."""
    global _lib
    if _lib is not None:
        return _lib
    import cv2
    lib = C.CDLL(LIB)
    vp, i, d = C.c_void_p, C.c_int, C.c_double
    PYR = C.CFUNCTYPE(None, vp, i, i, vp, i, i)
    THR = C.CFUNCTYPE(None, vp, i, i, d, d)
    CVT = C.CFUNCTYPE(None, vp, i, i, vp)
    HOU = C.CFUNCTYPE(i, vp, i, i, d, d, i, d, d, C.POINTER(i), i)
    LIN = C.CFUNCTYPE(None, vp, i, i, i, i, i, i, i, i)

    def pyrdown(src, rows, cols, dst, drows, dcols):
        _u8(dst, drows, dcols)[:] = cv2.pyrDown(_u8(src, rows, cols), dstsize=(dcols, drows))

    def threshold(img, rows, cols, thresh, maxval):
        a = _u8(img, rows, cols)
        a[:] = cv2.threshold(a, thresh, maxval, cv2.THRESH_BINARY)[1]

    def cvtcolor(bgr, rows, cols, gray):
        _u8(gray, rows, cols)[:] = cv2.cvtColor(_u8(bgr, rows, cols, 3), cv2.COLOR_BGR2GRAY)

    def hough(gray, rows, cols, rho, theta, thr, min_len, max_gap, out, cap):
        lines = cv2.HoughLinesP(_u8(gray, rows, cols), rho, theta, thr, minLineLength=min_len, maxLineGap=max_gap)
        n = 0 if lines is None else min(len(lines), cap)
        for k in range(n):
            for q in range(4):
                out[4 * k + q] = int(lines[k][0][q])
        return n

    def line(img, rows, cols, x1, y1, x2, y2, value, thickness):
        cv2.line(_u8(img, rows, cols), (x1, y1), (x2, y2), int(value), int(thickness))

    cbs = (PYR(pyrdown), THR(threshold), CVT(cvtcolor), HOU(hough), LIN(line))
    _keep.extend(cbs)
    lib.ref_host_set_callbacks.argtypes = [PYR, THR, CVT, HOU, LIN]
    lib.ref_host_set_callbacks(*cbs)
    lib.ref_host_set_rand.argtypes = [vp, C.c_size_t]
    lib.ref_host_rand_used.restype = C.c_size_t
    lib.ref_host_roberts.argtypes = [vp, i, i, vp]
    lib.ref_host_connect.argtypes = [vp, i, i, vp, vp, i, C.POINTER(i), vp, i]
    lib.ref_host_texture.argtypes = [vp, i, i, vp, C.POINTER(i), vp, vp, vp, vp, i]
    lib.ref_host_fit_regions.argtypes = [i, i, vp, vp, vp, i, vp, vp, vp, vp, C.c_float, C.c_float, C.c_float, vp, vp, vp]
    _lib = lib
    return lib


def roberts(gray):
    """roberts() main.cpp:214-240 (no threshold)."""
    g = np.ascontiguousarray(gray, np.uint8)
    out = np.empty_like(g)
    load().ref_host_roberts(g.ctypes.data, g.shape[0], g.shape[1], out.ctypes.data)
    return out


def connect(edges):
    """Connect() main.cpp:242-362 -> (labels, label_count, weak labels)."""
    e = np.ascontiguousarray(edges, np.uint8)
    h, w = e.shape
    labels = np.empty((h, w), np.int32)
    cnt = np.empty(h * w + 1, np.int32)
    weak = np.empty(h * w + 1, np.int32)
    n = C.c_int(0)
    nw = load().ref_host_connect(e.ctypes.data, h, w, labels.ctypes.data, cnt.ctypes.data, len(cnt), C.byref(n), weak.ctypes.data, len(weak))
    return labels, cnt[:n.value].copy(), weak[:nw].copy()


def texture(gray_full):
    """texture() main.cpp:365-596 on a full-resolution grey image -> dict(canny, text, cenxi, cenyi, size)."""
    g = np.ascontiguousarray(gray_full, np.uint8)
    h, w = g.shape
    canny = np.empty((h, w), np.float32)
    cap = (h // 4) * (w // 4) + 2
    text, size = np.empty(cap, np.float32), np.empty(cap, np.float32)
    cx, cy = np.empty(cap, np.int32), np.empty(cap, np.int32)
    n = C.c_int(0)
    rc = load().ref_host_texture(g.ctypes.data, h, w, canny.ctypes.data, C.byref(n), text.ctypes.data, cx.ctypes.data, cy.ctypes.data,
                                 size.ctypes.data, cap)
    if rc != 0:
        raise RuntimeError("ref_host_texture: region table larger than expected")
    k = n.value
    return dict(canny=canny, text=text[:k].copy(), cenxi=cx[:k].copy(), cenyi=cy[:k].copy(), size=size[:k].copy())


def fit_regions(cam0, cam_f, depth, scale, canny, text, size, rnd_stream, planes0, cenxi=None, cenyi=None):
    """The per-region RANSAC block main.cpp:1520-1730.  cam0: dict/struct with f, baseline, P_col34, M_inv; rnd_stream: the
    rand() values in call order (46 000 per region with text == -1, regions in index order).  Returns the region planes."""
    lib = load()
    depth, scale, canny = (np.ascontiguousarray(a, np.float32) for a in (depth, scale, canny))
    H, W = depth.shape
    text = np.ascontiguousarray(text, np.float32)
    size = np.ascontiguousarray(size, np.float32)
    n = len(text)
    cx = np.ascontiguousarray(cenxi if cenxi is not None else np.zeros(n), np.int32)
    cy = np.ascontiguousarray(cenyi if cenyi is not None else np.zeros(n), np.int32)
    stream = np.ascontiguousarray(rnd_stream, np.uint32).ravel()
    planes = np.ascontiguousarray(planes0, np.float32).reshape(n, 4).copy()
    pc = np.ascontiguousarray(np.asarray(cam0["P_col34"], np.float32).reshape(3))
    mi = np.ascontiguousarray(np.asarray(cam0["M_inv"], np.float32).reshape(9))
    lib.ref_host_set_rand(stream.ctypes.data, len(stream))
    lib.ref_host_fit_regions(W, H, depth.ctypes.data, scale.ctypes.data, canny.ctypes.data, n, text.ctypes.data, size.ctypes.data,
                             cx.ctypes.data, cy.ctypes.data, float(np.float32(cam_f)), float(np.float32(cam0["f"])),
                             float(np.float32(cam0["baseline"])), pc.ctypes.data, mi.ctypes.data, planes.ctypes.data)
    used = lib.ref_host_rand_used()
    lib.ref_host_set_rand(None, 0)
    return planes, used
