// weak_texture.cu -- host side of the weak-texture region detector, `texture()` in main.cpp:365-596 (SURVEY
// section 8 row f3).  The reference runs it once per reference view on the quarter-resolution grey image
// (<= 0.4 Mpx at C2) on the CPU; every stage below is a sequential raster scan whose result depends on the scan
// order (the label-equivalence table of Connect() is updated WITHOUT path compression or root lookup, so which
// components end up merged is a property of the visiting order) -- it stays host code, restated here so that the
// region labels the depth-completion kernels consume are the reference's.  The OpenCV calls in between
// (pyrDown, HoughLinesP, line) stay with the caller: the reference's host program and tsar-mvs_b200/texture.py
// both have OpenCV at hand.
//
// Stage order in the reference:  pyrDown x2 -> roberts -> threshold -> Connect -> per weak label: boundary image
// -> HoughLinesP -> line() into the edge map -> border closing -> Connect -> region statistics -> canny[] / text[]
// / cenxi / cenyi / size.
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include <vector>

#include "tsar_b200.h"

extern "C" {

// roberts() (main.cpp:214-240) followed by cv::threshold(dst, dst, thr, 255, THRESH_BINARY) (main.cpp:383).
// Interior pixels: sqrt((s(i,j)-s(i+1,j+1))^2 + (s(i+1,j)-s(i,j+1))^2); the outermost ring gets sqrt(2*5000) = 100.
// The reference stores `(uchar)sqrt(t1 + t2)`: magnitudes >= 256 keep only their low byte in the compiled program
// (double -> int -> 8 bit), so 256..260 fall back under a small threshold; reproduced with an explicit int cast.
int tsar_weak_edges(const unsigned char *gray, int w, int h, int thr, unsigned char *edges) {
    if (!gray || !edges || w < 1 || h < 1) return TSAR_ERR_ARG;
    for (int i = 0; i < h; i++)
        for (int j = 0; j < w; j++) {
            int t1, t2;
            if (i > 0 && i < h - 1 && j > 0 && j < w - 1) {
                const int a = (int)gray[(size_t)i * w + j] - (int)gray[(size_t)(i + 1) * w + j + 1];
                const int b = (int)gray[(size_t)(i + 1) * w + j] - (int)gray[(size_t)i * w + j + 1];
                t1 = a * a;
                t2 = b * b;
            } else {
                t1 = t2 = 100 * 50;
            }
            const unsigned char mag = (unsigned char)(int)sqrt((double)(t1 + t2));
            edges[(size_t)i * w + j] = mag > thr ? 255 : 0;
        }
    return TSAR_OK;
}

// Connect() (main.cpp:242-362): two-pass 4-connected labelling of the non-edge pixels (value != 255).
// labels: w*h int32 out (0 = edge pixel); label_count: up to cap entries out (entry 0 counts the edge pixels);
// *n_labels = number of labels including 0.  Returns TSAR_ERR_ARG when cap is too small (then *n_labels is the
// size needed).
int tsar_weak_connect(const unsigned char *edges, int w, int h, int *labels, int *label_count, int cap, int *n_labels) {
    if (!edges || !labels || !label_count || !n_labels || w < 1 || h < 1) return TSAR_ERR_ARG;
    std::vector<int> connection(1, 0);
    int cnt = 1;
    for (int y = 0; y < h; y++)
        for (int x = 0; x < w; x++) {
            const size_t p = (size_t)y * w + x;
            if (edges[p] == 255) { labels[p] = 0; continue; }
            // the reference's neighbour tables test `== 0` on both pixels (main.cpp:251, 267); its edge maps are binary
            const bool left = x > 0 && edges[p] == 0 && edges[p - 1] == 0;
            const bool up = y > 0 && edges[p] == 0 && edges[p - w] == 0;
            if (left) labels[p] = labels[p - 1];
            if (up) labels[p] = labels[p - w];
            if (!left && !up) {
                labels[p] = cnt;
                connection.push_back(cnt);
                cnt++;
            } else if (left && up) {
                const int ll = labels[p - 1], ul = labels[p - w];
                if (ll > ul) { connection[ll] = ul; labels[p] = ul; }        // direct overwrite, no root lookup
                else if (ll < ul) { connection[ul] = ll; labels[p] = ll; }
            }
        }
    for (size_t i = 1; i < connection.size(); i++) {
        int cur = connection[i], pre = connection[cur];
        while (pre != cur) { cur = pre; pre = connection[pre]; }
        connection[i] = cur;
    }
    std::vector<int> mapping(connection.size(), 0);
    int labelnum = 1;
    for (size_t i = 1; i < connection.size(); i++)
        if (connection[i] == (int)i) mapping[i] = labelnum++;
    for (size_t i = 1; i < connection.size(); i++) connection[i] = mapping[connection[i]];
    *n_labels = labelnum;
    if (labelnum > cap) return TSAR_ERR_ARG;
    for (int i = 0; i < labelnum; i++) label_count[i] = 0;
    for (size_t p = 0; p < (size_t)w * h; p++) {
        labels[p] = connection[labels[p]];
        label_count[labels[p]]++;
    }
    return TSAR_OK;
}

// The per-label boundary image that feeds HoughLinesP (main.cpp:392-421): 255 on pixels that do not carry `label`
// but have a 4-neighbour that does, else 0 (the reference builds it as BGR white/black and converts to grey).
int tsar_weak_boundary(const int *labels, int w, int h, int label, unsigned char *gray) {
    if (!labels || !gray || w < 1 || h < 1) return TSAR_ERR_ARG;
    for (int y = 0; y < h; y++)
        for (int x = 0; x < w; x++) {
            const size_t p = (size_t)y * w + x;
            unsigned char v = 0;
            if (labels[p] != label) {
                if (x > 0 && labels[p - 1] == label) v = 255;
                if (x < w - 1 && labels[p + 1] == label) v = 255;
                if (y > 0 && labels[p - w] == label) v = 255;
                if (y < h - 1 && labels[p + w] == label) v = 255;
            }
            gray[p] = v;
        }
    return TSAR_OK;
}

// Border closing before the second labelling (main.cpp:441-454): the outermost ring (always "edge" after roberts)
// is opened wherever its inner neighbour is free.  Rows first, then columns, in place, as written.
int tsar_weak_close_border(unsigned char *edges, int w, int h) {
    if (!edges || w < 2 || h < 2) return TSAR_ERR_ARG;
    for (int y = 0; y < h; y++) {
        if (edges[(size_t)y * w + 1] == 0) edges[(size_t)y * w] = 0;
        if (edges[(size_t)y * w + w - 2] == 0) edges[(size_t)y * w + w - 1] = 0;
    }
    for (int x = 0; x < w; x++) {
        if (edges[(size_t)1 * w + x] == 0) edges[x] = 0;
        if (edges[(size_t)(h - 2) * w + x] == 0) edges[(size_t)(h - 1) * w + x] = 0;
    }
    return TSAR_OK;
}

// Region statistics and the weak-texture decision (main.cpp:478-536, 570-593).  labels/label_count/n_labels from the
// second tsar_weak_connect.  Outputs, n_labels entries each (cannylines->text / cenxi / cenyi / size):
// text = 1, or -1 for "truly weak" regions: more than min_pixels pixels and (bounding-box area < size_ratio * pixels
// or more than 100 000 pixels); size = max bounding-box extent of those, 0 otherwise; cenxi/cenyi = centroid in
// full-resolution pixels (sum * 4 / count in `int`, wrapping like the compiled reference for very large regions).
// size_ratio: the reference declares `const int sizerat = 2.5` (main.cpp:64), i.e. 2.
int tsar_weak_regions(const int *labels, int w, int h, const int *label_count, int n_labels, int min_pixels, int size_ratio,
                      float *text, int *cenxi, int *cenyi, float *size) {
    if (!labels || !label_count || !text || !cenxi || !cenyi || !size || n_labels < 1) return TSAR_ERR_ARG;
    std::vector<uint32_t> sx(n_labels, 0), sy(n_labels, 0);
    std::vector<int> xmax(n_labels, 0), xmin(n_labels, w - 1), ymax(n_labels, 0), ymin(n_labels, h - 1);
    for (int y = 0; y < h; y++)
        for (int x = 0; x < w; x++) {
            const int l = labels[(size_t)y * w + x];
            if (l < 0 || l >= n_labels) return TSAR_ERR_ARG;
            sx[l] += (uint32_t)x;
            sy[l] += (uint32_t)y;
            if (x > xmax[l]) xmax[l] = x;
            if (x < xmin[l]) xmin[l] = x;
            if (y > ymax[l]) ymax[l] = y;
            if (y < ymin[l]) ymin[l] = y;
        }
    text[0] = 1.0f; cenxi[0] = 0; cenyi[0] = 0; size[0] = 0.0f;
    for (int l = 1; l < n_labels; l++) {
        const int n = label_count[l];
        cenxi[l] = n ? (int32_t)(sx[l] * 4u) / n : 0;
        cenyi[l] = n ? (int32_t)(sy[l] * 4u) / n : 0;
        text[l] = 1.0f;
        size[l] = 0.0f;
        if (n > min_pixels) {
            const int xs = xmax[l] - xmin[l], ys = ymax[l] - ymin[l];
            if (xs * ys < size_ratio * n || n > 100000) {
                text[l] = -1.0f;
                size[l] = (float)(xs > ys ? xs : ys);
            }
        }
    }
    return TSAR_OK;
}

}  // extern "C"
