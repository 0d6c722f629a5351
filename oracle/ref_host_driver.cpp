// oracle/ref_host_driver.cpp -- TEST INFRASTRUCTURE.  Compiles the reference's own HOST code for the two SURVEY section 8
// "next" rows that are CPU code in the reference:
//   f1  per-region RANSAC plane fit      main.cpp:1520-1730 (+ calcLinePara 147-164, disparityDepthConversion
//                                        cameraGeometryUtils.h:107-111)
//   f3  weak-texture detector texture()  main.cpp:365-596  (+ roberts / Connect 214-362, constants 59-64)
// main.cpp as a whole cannot be built in this image (OpenCV C++ headers absent, Windows calls), so oracle/build_ref.sh
// cuts exactly those line ranges out of the read-only checkout into a temp dir (ref_slice_*.inc, never stored in the
// repository) and this file wraps them: it supplies the few types the ranges name -- a cv::Mat with rows / cols / data /
// at<T>() / clone(), Vec3b, Vec4i, Point3f, Point2d, Scalar, Size, the GlobalState / LineState / InputFiles members the
// ranges touch -- and nothing of their logic.  The OpenCV LIBRARY calls inside texture() (pyrDown, threshold, cvtColor,
// HoughLinesP, line) are forwarded through callbacks to the caller's OpenCV (cv2 in the tests; the product side calls the
// same cv2 functions), imread returns the image handed in, imwrite is dropped.  rand() is redirected to an injected
// stream so that the RANSAC fit is reproducible (the reference never seeds it).
#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <ctime>
#include <iostream>
#include <memory>
#include <random>
#include <string>
#include <vector>

// ---- injected rand() stream ---------------------------------------------------------------------------------
static const uint32_t *g_rand = nullptr;
static size_t g_rand_n = 0, g_rand_pos = 0;
static int ref_host_rand() { return (g_rand && g_rand_pos < g_rand_n) ? (int)(g_rand[g_rand_pos++] & 0x7fffffffu) : 0; }
#define rand() ref_host_rand()

// ---- the OpenCV names the ranges use ----------------------------------------------------------------------------
namespace cv {
typedef unsigned char uchar;
struct Vec3b {
    uchar v[3];
    Vec3b() : v{0, 0, 0} {}
    Vec3b(int a, int b, int c) : v{(uchar)a, (uchar)b, (uchar)c} {}
    uchar &operator[](int i) { return v[i]; }
    bool operator==(const Vec3b &o) const { return v[0] == o.v[0] && v[1] == o.v[1] && v[2] == o.v[2]; }
};
struct Vec4i {
    int v[4];
    int &operator[](int i) { return v[i]; }
};
struct Point3f {
    float x, y, z;
    Point3f() : x(0), y(0), z(0) {}
    Point3f(float a, float b, float c) : x(a), y(b), z(c) {}
};
struct Point2d {
    double x, y;
    Point2d(double a, double b) : x(a), y(b) {}
};
struct Scalar {
    double v[4];
    Scalar(double a = 0, double b = 0, double c = 0, double d = 0) : v{a, b, c, d} {}
};
struct Size {
    int width, height;
    Size(int w, int h) : width(w), height(h) {}
};
enum { CV_8U = 0, CV_32S = 4, CV_8UC3 = 16, THRESH_BINARY = 0, COLOR_BGR2GRAY = 6 };
static const double CV_PI = 3.1415926535897932384626433832795;

struct Mat {
    int rows = 0, cols = 0, type = CV_8U;
    std::shared_ptr<std::vector<uchar>> buf;   // copies share the pixels, as cv::Mat copies do
    uchar *data = nullptr;
    Mat() {}
    Mat(int r, int c, int t) { create(r, c, t); }
    static size_t elt(int t) { return t == CV_32S ? 4 : (t == CV_8UC3 ? 3 : 1); }
    void create(int r, int c, int t) {
        rows = r; cols = c; type = t;
        buf = std::make_shared<std::vector<uchar>>((size_t)r * c * elt(t));
        data = buf->data();
    }
    template <class T> T &at(int y, int x) { return reinterpret_cast<T *>(data)[(size_t)y * cols + x]; }
    Mat clone() const {
        Mat m;
        m.create(rows, cols, type);
        if (data) memcpy(m.data, data, (size_t)rows * cols * elt(type));
        return m;
    }
};
}  // namespace cv

// ---- OpenCV library calls of texture(): forwarded to the caller's OpenCV ------------------------------------------
extern "C" {
typedef void (*pyrdown_cb_t)(const unsigned char *src, int rows, int cols, unsigned char *dst, int drows, int dcols);
typedef void (*threshold_cb_t)(unsigned char *img, int rows, int cols, double thresh, double maxval);
typedef void (*cvtcolor_cb_t)(const unsigned char *bgr, int rows, int cols, unsigned char *gray);
typedef int (*hough_cb_t)(const unsigned char *gray, int rows, int cols, double rho, double theta, int thr, double min_len, double max_gap,
                          int *lines4, int cap);
typedef void (*line_cb_t)(unsigned char *img, int rows, int cols, int x1, int y1, int x2, int y2, int value, int thickness);
}
static pyrdown_cb_t g_pyrdown;
static threshold_cb_t g_threshold;
static cvtcolor_cb_t g_cvtcolor;
static hough_cb_t g_hough;
static line_cb_t g_line;
static cv::Mat g_source;   // what imread "reads"

namespace cv {
static Mat imread(const std::string &, int) { return g_source; }
static bool imwrite(const std::string &, const Mat &) { return true; }
static void pyrDown(const Mat &src, Mat &dst, const Size &s) {
    dst.create(s.height, s.width, CV_8U);
    g_pyrdown(src.data, src.rows, src.cols, dst.data, dst.rows, dst.cols);
}
static void threshold(Mat &src, Mat &, double thresh, double maxval, int) { g_threshold(src.data, src.rows, src.cols, thresh, maxval); }
static void cvtColor(const Mat &src, Mat &dst, int) {
    dst.create(src.rows, src.cols, CV_8U);
    g_cvtcolor(src.data, src.rows, src.cols, dst.data);
}
static void HoughLinesP(const Mat &img, std::vector<Vec4i> &lines, double rho, double theta, int thr, double min_len, double max_gap) {
    std::vector<int> out(4 * 65536);
    const int n = g_hough(img.data, img.rows, img.cols, rho, theta, thr, min_len, max_gap, out.data(), 65536);
    lines.resize(n);
    for (int i = 0; i < n; i++)
        for (int k = 0; k < 4; k++) lines[i][k] = out[4 * i + k];
}
static void line(Mat &img, const Point2d &a, const Point2d &b, const Scalar &col, int thickness) {
    g_line(img.data, img.rows, img.cols, (int)lrint(a.x), (int)lrint(a.y), (int)lrint(b.x), (int)lrint(b.y), (int)col.v[0], thickness);
}
}  // namespace cv

using namespace cv;
using namespace std;

// ---- the state members the ranges touch (globalstate.h:25-54, linestate.h:10-221, cameraparameters.h, camera.h, main.h) ----
struct float3 { float x, y, z; };
struct float4 { float x, y, z, w; };
struct LineState {
    int n = 0;
    std::vector<float> canny_v, text_v, size_v, depth_v, scale_v;
    std::vector<int> cenxi_v, cenyi_v;
    std::vector<float4> norm4_v;
    float *canny = nullptr, *text = nullptr, *size = nullptr, *depth = nullptr, *scale = nullptr;
    int *cenxi = nullptr, *cenyi = nullptr;
    float4 *norm4 = nullptr;
    void resize(int k) {                       // linestate.h:64-109 (zero-filled arrays)
        canny_v.assign(k, 0.f); depth_v.assign(k, 0.f); scale_v.assign(k, 0.f);
        canny = canny_v.data(); depth = depth_v.data(); scale = scale_v.data();
    }
    void Cannyresize(int k) {                  // linestate.h:163-184
        text_v.assign(k, 0.f); size_v.assign(k, 0.f); cenxi_v.assign(k, 0); cenyi_v.assign(k, 0); norm4_v.assign(k, float4{0, 0, 0, 0});
        text = text_v.data(); size = size_v.data(); cenxi = cenxi_v.data(); cenyi = cenyi_v.data(); norm4 = norm4_v.data();
    }
};
struct Camera_cu { float f = 0, baseline = 1; float3 P_col34{0, 0, 0}; float M_inv[9] = {0}; };
struct CameraParameters_cu { float f = 0; Camera_cu cameras[1]; };
struct GlobalState {
    CameraParameters_cu *cameras = nullptr;
    LineState *lines = nullptr, *cannylines = nullptr;
    int col = 0, row = 0;
};
struct InputFiles {
    std::string images_folder, connect_folder, weak_folder;
    std::vector<std::string> img_filenames;
};
#define REFERENCE 0   // config.h:21

// ---- the reference's own code, cut out of the checkout by oracle/build_ref.sh -----------------------------------------
#include "ref_slice_constants.inc"    // main.cpp:59-64    Robthr .. sizerat
#include "ref_slice_ddc.inc"          // cameraGeometryUtils.h:107-111  disparityDepthConversion
#include "ref_slice_calcline.inc"     // main.cpp:147-164  calcLinePara
#include "ref_slice_connect.inc"      // main.cpp:214-362  roberts, Connect
#include "ref_slice_texture.inc"      // main.cpp:365-596  texture()

static void ref_ransac_block(GlobalState *gs, CameraParameters_cu &cameraParams) {
#include "ref_slice_ransac.inc"       // main.cpp:1520-1730  if (true) { ... per-region RANSAC ... }
}

// ---- C entry points for the tests -----------------------------------------------------------------------------------
extern "C" {

void ref_host_set_rand(const uint32_t *stream, size_t n) { g_rand = stream; g_rand_n = n; g_rand_pos = 0; }
size_t ref_host_rand_used(void) { return g_rand_pos; }

void ref_host_set_callbacks(pyrdown_cb_t a, threshold_cb_t b, cvtcolor_cb_t c, hough_cb_t d, line_cb_t e) {
    g_pyrdown = a; g_threshold = b; g_cvtcolor = c; g_hough = d; g_line = e;
}

// roberts() alone (main.cpp:214-240): uchar image in, Roberts magnitude out
void ref_host_roberts(const unsigned char *gray, int rows, int cols, unsigned char *out) {
    Mat src(rows, cols, CV_8U);
    memcpy(src.data, gray, (size_t)rows * cols);
    Mat dst = roberts(src);
    memcpy(out, dst.data, (size_t)rows * cols);
}

// Connect() alone (main.cpp:242-362): edge map (0 / 255) in; labels, per-label counts (cap entries), number of labels and
// the labels with more than weaktextnum pixels out.  Returns the number of weak labels.
int ref_host_connect(const unsigned char *edges, int rows, int cols, int *labels, int *label_count, int cap, int *n_labels, int *weak, int weak_cap) {
    Mat e(rows, cols, CV_8U);
    memcpy(e.data, edges, (size_t)rows * cols);
    Mat lab(rows, cols, CV_32S);
    memset(lab.data, 0, (size_t)rows * cols * 4);
    std::vector<int> cnt, wk;
    Connect(e, lab, cnt, wk);
    memcpy(labels, lab.data, (size_t)rows * cols * 4);
    *n_labels = (int)cnt.size();
    for (int i = 0; i < (int)cnt.size() && i < cap; i++) label_count[i] = cnt[i];
    for (int i = 0; i < (int)wk.size() && i < weak_cap; i++) weak[i] = wk[i];
    return (int)wk.size();
}

// texture() (main.cpp:365-596) on a full-resolution grey image.  Outputs: canny (rows*cols floats, lines->canny), the number
// of regions, and up to cap entries of cannylines->text / cenxi / cenyi / size.  Returns 0, or -1 when cap is too small.
int ref_host_texture(const unsigned char *gray, int rows, int cols, float *canny, int *n_regions, float *text, int *cenxi, int *cenyi,
                     float *size, int cap) {
    g_source = Mat(rows, cols, CV_8U);
    memcpy(g_source.data, gray, (size_t)rows * cols);
    InputFiles in;
    in.img_filenames.push_back("00000000.jpg");
    LineState lines, cannylines;
    GlobalState gs;
    gs.lines = &lines; gs.cannylines = &cannylines;
    texture(in, &gs);
    memcpy(canny, lines.canny, (size_t)rows * cols * sizeof(float));
    *n_regions = cannylines.n;
    if (cannylines.n > cap) return -1;
    for (int i = 0; i < cannylines.n; i++) { text[i] = cannylines.text[i]; cenxi[i] = cannylines.cenxi[i]; cenyi[i] = cannylines.cenyi[i]; size[i] = cannylines.size[i]; }
    return 0;
}

// The per-region RANSAC block of runGipuma (main.cpp:1520-1730).  Inputs: W x H maps lines->depth (a disparity), ->scale,
// ->canny; the region table (text, size, cenxi, cenyi; n_regions entries); camera 0's f, baseline, P_col34, M_inv and
// CameraParameters_cu::f; region_norm4 holds cannylines->norm4 on entry and receives the fitted planes.  rand() values come
// from ref_host_set_rand (46 000 per fitted region, in region order).
int ref_host_fit_regions(int W, int H, const float *depth, const float *scale, const float *canny, int n_regions, const float *text,
                         const float *size, const int *cenxi, const int *cenyi, float cam_f, float cam0_f, float baseline,
                         const float *P_col34, const float *M_inv, float *region_norm4) {
    LineState lines, cannylines;
    lines.resize(W * H);
    memcpy(lines.depth, depth, (size_t)W * H * 4);
    memcpy(lines.scale, scale, (size_t)W * H * 4);
    memcpy(lines.canny, canny, (size_t)W * H * 4);
    cannylines.n = n_regions;
    cannylines.Cannyresize(n_regions);
    for (int i = 0; i < n_regions; i++) {
        cannylines.text[i] = text[i]; cannylines.size[i] = size[i];
        cannylines.cenxi[i] = cenxi ? cenxi[i] : 0; cannylines.cenyi[i] = cenyi ? cenyi[i] : 0;
        cannylines.norm4[i] = float4{region_norm4[4 * i], region_norm4[4 * i + 1], region_norm4[4 * i + 2], region_norm4[4 * i + 3]};
    }
    CameraParameters_cu cp;
    cp.f = cam_f;
    cp.cameras[0].f = cam0_f; cp.cameras[0].baseline = baseline;
    cp.cameras[0].P_col34 = float3{P_col34[0], P_col34[1], P_col34[2]};
    for (int i = 0; i < 9; i++) cp.cameras[0].M_inv[i] = M_inv[i];
    GlobalState gs;
    gs.cameras = &cp; gs.lines = &lines; gs.cannylines = &cannylines; gs.col = W; gs.row = H;
    ref_ransac_block(&gs, cp);
    for (int i = 0; i < n_regions; i++) {
        region_norm4[4 * i] = cannylines.norm4[i].x; region_norm4[4 * i + 1] = cannylines.norm4[i].y;
        region_norm4[4 * i + 2] = cannylines.norm4[i].z; region_norm4[4 * i + 3] = cannylines.norm4[i].w;
    }
    return 0;
}

}  // extern "C"
