"""TEST INFRASTRUCTURE -- plain-Python restatement of the scan-order dependent stages of the reference's weak-texture
detector, `texture()` main.cpp:365-596 (roberts 214-240, Connect 242-362).  Only tests/ may import this.  Parity
unpinned by the reference (no tests, and main.cpp cannot be built here: OpenCV C++, Windows calls); the OpenCV stages
(pyrDown / HoughLinesP / line) are the same cv2 calls on both sides.  Written with explicit loops in the reference's
statement order; use on small images only."""
import math

import numpy as np


def roberts_threshold(src, thr):
    """main.cpp:214-240 + 383."""
    h, w = src.shape
    out = np.zeros((h, w), np.uint8)
    for i in range(h):
        for j in range(w):
            if 0 < i < h - 1 and 0 < j < w - 1:
                t1 = (int(src[i, j]) - int(src[i + 1, j + 1])) ** 2
                t2 = (int(src[i + 1, j]) - int(src[i, j + 1])) ** 2
            else:
                t1 = t2 = 100 * 50
            mag = int(math.sqrt(t1 + t2)) & 0xFF          # (uchar) of a double: low byte in the compiled program
            out[i, j] = 255 if mag > thr else 0
    return out


def connect(dst):
    """main.cpp:242-362. Returns (labels, label_count list)."""
    h, w = dst.shape
    lab = np.zeros((h, w), np.int64)
    connection = [0]
    cnt = 1
    for y in range(h):
        for x in range(w):
            if dst[y, x] == 255:
                lab[y, x] = 0
                continue
            left = x > 0 and dst[y, x] == 0 and dst[y, x - 1] == 0
            up = y > 0 and dst[y, x] == 0 and dst[y - 1, x] == 0
            if left:
                lab[y, x] = lab[y, x - 1]
            if up:
                lab[y, x] = lab[y - 1, x]
            if not left and not up:
                lab[y, x] = cnt
                connection.append(cnt)
                cnt += 1
            elif left and up:
                ll, ul = int(lab[y, x - 1]), int(lab[y - 1, x])
                if ll > ul:
                    connection[ll] = ul
                    lab[y, x] = ul
                elif ll < ul:
                    connection[ul] = ll
                    lab[y, x] = ll
    for i in range(1, len(connection)):
        cur = connection[i]
        pre = connection[cur]
        while pre != cur:
            cur = pre
            pre = connection[pre]
        connection[i] = cur
    mapping = [0] * len(connection)
    labelnum = 1
    for i in range(1, len(connection)):
        if connection[i] == i:
            mapping[i] = labelnum
            labelnum += 1
    for i in range(1, len(connection)):
        connection[i] = mapping[connection[i]]
    counts = [0] * labelnum
    for y in range(h):
        for x in range(w):
            lab[y, x] = connection[int(lab[y, x])]
            counts[int(lab[y, x])] += 1
    return lab.astype(np.int32), np.array(counts, np.int32)


def boundary(lab, k):
    """main.cpp:392-421."""
    h, w = lab.shape
    out = np.zeros((h, w), np.uint8)
    for y in range(h):
        for x in range(w):
            if lab[y, x] != k:
                if (x > 0 and lab[y, x - 1] == k) or (x < w - 1 and lab[y, x + 1] == k) or \
                        (y > 0 and lab[y - 1, x] == k) or (y < h - 1 and lab[y + 1, x] == k):
                    out[y, x] = 255
    return out


def close_border(dst):
    """main.cpp:441-454."""
    d = dst.copy()
    h, w = d.shape
    for y in range(h):
        if d[y, 1] == 0:
            d[y, 0] = 0
        if d[y, w - 2] == 0:
            d[y, w - 1] = 0
    for x in range(w):
        if d[1, x] == 0:
            d[0, x] = 0
        if d[h - 2, x] == 0:
            d[h - 1, x] = 0
    return d


def regions(lab, counts, weaktextnum, sizerat):
    """main.cpp:478-536 + 570-593."""
    h, w = lab.shape
    n = len(counts)
    weak = [i for i in range(1, n) if counts[i] > weaktextnum]
    xmax = {i: 0 for i in weak}; xmin = {i: w - 1 for i in weak}; ymax = {i: 0 for i in weak}; ymin = {i: h - 1 for i in weak}
    lx, ly = [0] * n, [0] * n
    for y in range(h):
        for x in range(w):
            l = int(lab[y, x])
            lx[l] += x
            ly[l] += y
            if l in xmax:
                xmax[l] = max(xmax[l], x); ymax[l] = max(ymax[l], y)
                xmin[l] = min(xmin[l], x); ymin[l] = min(ymin[l], y)
    text = np.ones(n, np.float32)
    size = np.zeros(n, np.float32)
    cx, cy = np.zeros(n, np.int32), np.zeros(n, np.int32)

    def i32(v):
        v &= 0xFFFFFFFF
        return v - (1 << 32) if v & 0x80000000 else v

    def cdiv(a, b):  # C integer division truncates toward zero
        q = abs(a) // abs(b)
        return q if (a >= 0) == (b >= 0) else -q
    for t in range(1, n):
        cx[t] = cdiv(i32(i32(lx[t]) * 4), int(counts[t])) if counts[t] else 0
        cy[t] = cdiv(i32(i32(ly[t]) * 4), int(counts[t])) if counts[t] else 0
    for l in weak:
        xs, ys = xmax[l] - xmin[l], ymax[l] - ymin[l]
        if xs * ys < sizerat * counts[l] or counts[l] > 100000:
            text[l] = -1.0
            size[l] = max(xs, ys)
    return text, cx, cy, size


def expand(lab, W, H):
    """main.cpp:558-568."""
    hq, wq = lab.shape
    out = np.zeros((H, W), np.float32)
    for y in range(H):
        for x in range(W):
            sx, sy = x // 4, y // 4
            if sx >= wq:
                sx -= 1
            if sy >= hq:
                sy -= 1
            out[y, x] = lab[sy, sx]
    return out
