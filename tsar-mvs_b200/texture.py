"""Weak-texture region detector: host-side mirror of `texture()` (main.cpp:365-596).

Produces what the depth-completion kernels consume: the per-pixel region label map `lines->canny` and the per-region
tables `cannylines->text` (1 textured / -1 weakly textured), `cenxi`, `cenyi`, `size`.  The scan-order dependent
stages (Roberts edges, two-pass labelling with the reference's own equivalence table, region statistics) are the C
functions `tsar_weak_*` of libtsar_b200.so; the OpenCV stages the reference calls (pyrDown, HoughLinesP, line) are
called through cv2 here, as the reference's host program calls them through the C++ API.
"""
import ctypes as C

import numpy as np

from . import _lib as L

ROBTHR, HOUTHR, MIN_LINE_LENGTH, MAX_LINE_GAP, WEAKTEXTNUM, SIZERAT = 4, 110, 160, 18, 5000, 2   # main.cpp:59-64


def _ck(rc, what):
    if rc != 0:
        raise RuntimeError(f"{what} failed ({rc})")


def edges(gray, thr=ROBTHR):
    gray = np.ascontiguousarray(gray, np.uint8)
    out = np.empty_like(gray)
    _ck(L.load().tsar_weak_edges(gray.ctypes.data, gray.shape[1], gray.shape[0], int(thr), out.ctypes.data), "tsar_weak_edges")
    return out


def connect(edge_map):
    """Connect(): returns (labels int32 [h][w], label_count int32 [n_labels])."""
    e = np.ascontiguousarray(edge_map, np.uint8)
    h, w = e.shape
    labels = np.empty((h, w), np.int32)
    cap = 1024
    lib = L.load()
    while True:
        cnt = np.empty(cap, np.int32)
        n = C.c_int(0)
        rc = lib.tsar_weak_connect(e.ctypes.data, w, h, labels.ctypes.data, cnt.ctypes.data, cap, C.byref(n))
        if rc == 0:
            return labels, cnt[:n.value].copy()
        if n.value <= cap:
            _ck(rc, "tsar_weak_connect")
        cap = n.value


def boundary(labels, label):
    labels = np.ascontiguousarray(labels, np.int32)
    out = np.empty(labels.shape, np.uint8)
    _ck(L.load().tsar_weak_boundary(labels.ctypes.data, labels.shape[1], labels.shape[0], int(label), out.ctypes.data), "tsar_weak_boundary")
    return out


def close_border(edge_map):
    e = np.ascontiguousarray(edge_map, np.uint8).copy()
    _ck(L.load().tsar_weak_close_border(e.ctypes.data, e.shape[1], e.shape[0]), "tsar_weak_close_border")
    return e


def regions(labels, label_count, min_pixels=WEAKTEXTNUM, size_ratio=SIZERAT):
    labels = np.ascontiguousarray(labels, np.int32)
    cnt = np.ascontiguousarray(label_count, np.int32)
    n = len(cnt)
    text, size = np.empty(n, np.float32), np.empty(n, np.float32)
    cx, cy = np.empty(n, np.int32), np.empty(n, np.int32)
    _ck(L.load().tsar_weak_regions(labels.ctypes.data, labels.shape[1], labels.shape[0], cnt.ctypes.data, n, int(min_pixels), int(size_ratio),
                                   text.ctypes.data, cx.ctypes.data, cy.ctypes.data, size.ctypes.data), "tsar_weak_regions")
    return text, cx, cy, size


def hough_close(edge_map, labels, label_count, min_pixels=WEAKTEXTNUM, hough=(HOUTHR, MIN_LINE_LENGTH, MAX_LINE_GAP)):
    """main.cpp:389-436: for every weak label of the first labelling, straight boundary segments found by
    HoughLinesP are drawn into the edge map (closes gaps in the region's outline)."""
    import cv2  # the reference's own third-party dependency for this stage
    e = edge_map.copy()
    for lab in range(1, len(label_count)):
        if label_count[lab] <= min_pixels:
            continue
        lines = cv2.HoughLinesP(boundary(labels, lab), 1, np.pi / 180, hough[0], minLineLength=hough[1], maxLineGap=hough[2])
        for ln in (lines if lines is not None else []):
            x1, y1, x2, y2 = (int(v) for v in ln[0])
            cv2.line(e, (x1, y1), (x2, y2), 255, 1)
    return e


def quarter_gray(gray_full):
    """main.cpp:375-379: two pyrDown with explicit halved sizes."""
    import cv2
    g = np.ascontiguousarray(gray_full, np.uint8)
    d2 = cv2.pyrDown(g, dstsize=(g.shape[1] // 2, g.shape[0] // 2))
    return cv2.pyrDown(d2, dstsize=(d2.shape[1] // 2, d2.shape[0] // 2))


def detect(gray_full, min_pixels=WEAKTEXTNUM, thr=ROBTHR, hough=(HOUTHR, MIN_LINE_LENGTH, MAX_LINE_GAP), quarter=None):
    """texture(): full-resolution 8-bit grey image -> dict(labels_q, label_count, text, cenxi, cenyi, size).
    `labels_q` is the quarter-resolution label map; DepthmapEngine.set_labels_quarter expands it to canny[] on the
    device, `expand_labels` does the same on the host."""
    q = quarter_gray(gray_full) if quarter is None else np.ascontiguousarray(quarter, np.uint8)
    e0 = edges(q, thr)
    lab0, cnt0 = connect(e0)
    e1 = close_border(hough_close(e0, lab0, cnt0, min_pixels, hough))
    lab, cnt = connect(e1)
    text, cx, cy, size = regions(lab, cnt, min_pixels)
    return dict(labels_q=lab, label_count=cnt, text=text, cenxi=cx, cenyi=cy, size=size, edges=e1)


def expand_labels(labels_q, W, H):
    """lines->canny on the host (main.cpp:558-568)."""
    hq, wq = labels_q.shape
    sx = np.arange(W) // 4
    sy = np.arange(H) // 4
    sx[sx >= wq] -= 1
    sy[sy >= hq] -= 1
    return labels_q[np.ix_(sy, sx)].astype(np.float32)
