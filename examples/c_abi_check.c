/* Plain-C consumer of include/tsar_b200.h: proves the boundary is a C ABI (compiles with gcc -std=c11, no C++),
 * and shows the call order of a resident depthmap.  Without an sm_100 device tsar_create must fail with
 * TSAR_ERR_NODEVICE (there is no CPU path); with one, a tiny flat scene is processed end to end.
 *   gcc -std=c11 -Iinclude examples/c_abi_check.c -Ltsar-mvs_b200 -ltsar_b200 -Wl,-rpath,$PWD/tsar-mvs_b200 -lm -o /tmp/c_abi_check */
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "tsar_b200.h"

static void identity3(float *m) { memset(m, 0, 9 * sizeof(float)); m[0] = m[4] = m[8] = 1.0f; }

int main(void) {
    printf("library: %s\n", tsar_version());
    tsar_ctx *ctx = NULL;
    int rc = tsar_create(0, NULL, &ctx);
    if (rc == TSAR_ERR_NODEVICE) {
        printf("no sm_100 device: tsar_create -> TSAR_ERR_NODEVICE (expected on a CPU-only machine)\n");
        return 0;
    }
    if (rc != TSAR_OK) { fprintf(stderr, "tsar_create failed: %d\n", rc); return 1; }
    /* two views of a fronto-parallel textured plane at depth 2, second camera shifted by 0.1 along x */
    enum { W = 96, H = 64, NV = 2 };
    static float img[NV][W * H];
    const float fx = 120.0f, cx = (W - 1) * 0.5f, cy = (H - 1) * 0.5f, depth = 2.0f, base = 0.1f;
    for (int v = 0; v < NV; v++)
        for (int y = 0; y < H; y++)
            for (int x = 0; x < W; x++) {
                const float X = (x - cx) * depth / fx + (v ? base : 0.0f), Y = (y - cy) * depth / fx;
                img[v][y * W + x] = floorf(127.5f + 60.0f * sinf(9.0f * X) + 50.0f * sinf(13.0f * Y + 4.0f * X));
            }
    tsar_camera cams[NV];
    memset(cams, 0, sizeof(cams));
    for (int v = 0; v < NV; v++) {
        tsar_camera *c = &cams[v];
        identity3(c->K); identity3(c->K_inv); identity3(c->R); identity3(c->R_orig); identity3(c->R_orig_inv); identity3(c->M_inv);
        c->K[0] = c->K[4] = fx; c->K[2] = cx; c->K[5] = cy;
        c->K_inv[0] = c->K_inv[4] = 1.0f / fx; c->K_inv[2] = -cx / fx; c->K_inv[5] = -cy / fx;
        memcpy(c->M_inv, c->K_inv, sizeof(c->K_inv));              /* P = K [I | t]  ->  M^-1 = K^-1 */
        c->t4[0] = v ? -base : 0.0f;                                /* t = -R C */
        c->P_col34[0] = fx * c->t4[0];
        c->C4[0] = v ? base : 0.0f;
        c->fx = c->fy = c->f = fx; c->alpha = 1.0f; c->baseline = 1.0f; c->depthMin = 1.0f; c->depthMax = 4.0f;
    }
    const float *views[NV] = {img[0], img[1]};
    const int subset[1] = {1};
    tsar_params p;
    memset(&p, 0, sizeof(p));
    p.box_hsize = p.box_vsize = 11; p.iterations = 4; p.n_best = 1; p.cost_comb = 1;
    p.min_disparity = fx / 4.0f; p.max_disparity = fx / 1.0f;
    static float out[W * H * 4], conf[W * H];
    rc = tsar_depthmap_host(ctx, W, H, NV, views, cams, fx, subset, 1, &p, 1234u, out, conf);
    if (rc != TSAR_OK) { fprintf(stderr, "tsar_depthmap_host failed (%d): %s\n", rc, tsar_last_error(ctx)); return 1; }
    int good = 0, n = 0;
    for (int y = 8; y < H - 8; y++)
        for (int x = 8; x < W - 8; x++, n++) good += fabsf(out[(y * W + x) * 4 + 3] - depth) < 0.02f * depth;
    printf("depth within 2%% of the true plane on %d of %d interior pixels\n", good, n);
    tsar_destroy(ctx);
    return good > n * 8 / 10 ? 0 : 2;
}
