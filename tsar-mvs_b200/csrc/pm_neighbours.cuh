// pm_neighbours.cuh -- which neighbour's plane is tried at a pixel in each of the eight propagation steps of
// gipuma_checkerboard_spatialProp_cu (gipuma.cu:888-1042).  The pick depends only on the pre-launch costs (snapshot
// semantics), never on what the pixel has accepted so far, so all eight picks of a pixel can be made up
// front.  Shared by the checkerboard kernel and the candidate-statistics kernel.
#pragma once
#include <cuda_runtime.h>

namespace tsar {

// cO / cS: pre-launch costs of the opposite / the same colour (full-size arrays).  Returns the pixel index of the
// neighbour whose plane is the candidate of `step` (0..7, the reference's order) or -1 when the border guard of that
// direction fails; own = the neighbour has the pixel's own colour (its plane is read from the same-colour snapshot).
__device__ __forceinline__ int pick_neighbour(int step, int x, int y, int W, int H, int pidx, const float *__restrict__ cO,
                                              const float *__restrict__ cS, bool &own) {
    float cmin = 0.f;
    int best = -1;
    own = false;
    switch (step) {
        case 0:  // up_far: 11 samples, stride 2, opposite colour
            if (y > 2) {
                best = pidx - 3 * W; cmin = cO[best];
#pragma unroll
                for (int i = 1; i < 11; i++)
                    if (y > 2 + 2 * i) { const int pt = pidx - (3 + 2 * i) * W; const float v = cO[pt]; if (v < cmin) { cmin = v; best = pt; } }
            }
            break;
        case 1:  // down_far: the running minimum starts from c[up_far]; out of bounds for
                 // y < 3, where the reference reads the zero guard (Q5)
            if (y < H - 3) {
                cmin = (y >= 3) ? cO[pidx - 3 * W] : 0.0f;
                best = pidx + 3 * W;
#pragma unroll
                for (int i = 1; i < 11; i++)
                    if (y < H - 3 - 2 * i) { const int pt = pidx + (3 + 2 * i) * W; const float v = cO[pt]; if (v < cmin) { cmin = v; best = pt; } }
            }
            break;
        case 2:  // left_far
            if (x > 2) {
                best = pidx - 3; cmin = cO[best];
#pragma unroll
                for (int i = 1; i < 11; i++)
                    if (x > 2 + 2 * i) { const int pt = pidx - 3 - 2 * i; const float v = cO[pt]; if (v < cmin) { cmin = v; best = pt; } }
            }
            break;
        case 3:  // right_far: comparison inverted in the reference (tracks the maximum, Q6)
            if (x < W - 3) {
                best = pidx + 3; cmin = cO[best];
#pragma unroll
                for (int i = 1; i < 11; i++)
                    if (x < W - 3 - 2 * i) { const int pt = pidx + 3 + 2 * i; const float v = cO[pt]; if (cmin < v) { cmin = v; best = pt; } }
            }
            break;
        // near "V" areas: the direct neighbour (opposite colour) + same-colour extras, which are read
        // from the pre-launch snapshot (gipuma.cu:952-1042)
        case 4:  // up_near
            if (y > 0) {
                best = pidx - W; cmin = cO[best];
#pragma unroll
                for (int i = 0; i < 3; i++) {
                    if (y > 1 + i && x > i) { const int pt = pidx - (2 + i) * W - i; const float v = cS[pt]; if (v < cmin) { cmin = v; best = pt; own = true; } }
                    if (y > 1 + i && x < W - 1 - i) { const int pt = pidx - (2 + i) * W + i; const float v = cS[pt]; if (v < cmin) { cmin = v; best = pt; own = true; } }
                }
            }
            break;
        case 5:  // down_near
            if (y < H - 1) {
                best = pidx + W; cmin = cO[best];
#pragma unroll
                for (int i = 0; i < 3; i++) {
                    if (y < H - 2 - i && x > i) { const int pt = pidx + (2 + i) * W - i; const float v = cS[pt]; if (v < cmin) { cmin = v; best = pt; own = true; } }
                    if (y < H - 2 - i && x < W - 1 - i) { const int pt = pidx + (2 + i) * W + i; const float v = cS[pt]; if (v < cmin) { cmin = v; best = pt; own = true; } }
                }
            }
            break;
        case 6:  // left_near
            if (x > 0) {
                best = pidx - 1; cmin = cO[best];
#pragma unroll
                for (int i = 0; i < 3; i++) {
                    if (x > 1 + i && y > i) { const int pt = pidx - (2 + i) - i * W; const float v = cS[pt]; if (v < cmin) { cmin = v; best = pt; own = true; } }
                    if (x > 1 + i && y < H - 1 - i) { const int pt = pidx - (2 + i) + i * W; const float v = cS[pt]; if (v < cmin) { cmin = v; best = pt; own = true; } }
                }
            }
            break;
        default:  // 7: right_near
            if (x < W - 1) {
                best = pidx + 1; cmin = cO[best];
#pragma unroll
                for (int i = 0; i < 3; i++) {
                    if (x < W - 2 - i && y > i) { const int pt = pidx + (2 + i) - i * W; const float v = cS[pt]; if (v < cmin) { cmin = v; best = pt; own = true; } }
                    if (x < W - 2 - i && y < H - 1 - i) { const int pt = pidx + (2 + i) + i * W; const float v = cS[pt]; if (v < cmin) { cmin = v; best = pt; own = true; } }
                }
            }
            break;
    }
    return best;
}

__device__ __forceinline__ bool same_bits(const float4 &a, const float4 &b) {
    return __float_as_uint(a.x) == __float_as_uint(b.x) && __float_as_uint(a.y) == __float_as_uint(b.y) &&
           __float_as_uint(a.z) == __float_as_uint(b.z) && __float_as_uint(a.w) == __float_as_uint(b.w);
}

}  // namespace tsar
