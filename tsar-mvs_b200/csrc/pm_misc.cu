// XORWOW row-table and colour-merge launchers (window independent)
#include <curand_kernel.h>

#include "pm_launch.h"

namespace tsar {

// ---------------------------------------------------------------------------------------------
// XORWOW row tables.  The reference calls curand_init(seed, subsequence = y, offset = x) in every
// pixel of every init/refine launch (gipuma.cu:700, 1077): two matrix skip-aheads per pixel.  Offset
// x is just "x draws later in the stream of row y", so one thread per row initialises the row stream
// once (curand_init(seed, y, 0)) and steps it; the n-th draw of pixel (x,y) is table[y][x+n].
// One warp handles 32 rows and transposes 32x32 blocks through shared memory so the table is written
// with coalesced 128-byte rows.
// ---------------------------------------------------------------------------------------------
// A row is cut into segments of kRngSeg draws (blockIdx.y); a segment starts its stream at offset s*kRngSeg
// (curand_init's third argument), so 8x more threads share the serial stepping of a C2 row.
constexpr int kRngSeg = 416;  // multiple of 32

__global__ void __launch_bounds__(32) rng_rows_kernel(uint32_t *__restrict__ table, int pitch, int H, int len,
                                                      unsigned long long seed) {
    __shared__ uint32_t tile[32][33];
    const int lane = threadIdx.x;
    const int y = blockIdx.x * 32 + lane;
    const int seg0 = blockIdx.y * kRngSeg, seg1 = min(len, seg0 + kRngSeg);
    curandStateXORWOW_t st;
    curand_init(seed, (unsigned long long)min(y, H - 1), (unsigned long long)seg0, &st);
    for (int base = seg0; base < seg1; base += 32) {
#pragma unroll 4
        for (int k = 0; k < 32; k++) tile[lane][k] = curand(&st);
        __syncwarp();
        for (int r = 0; r < 32; r++) {
            const int yy = blockIdx.x * 32 + r;
            if (yy < H && base + lane < seg1) table[(size_t)yy * pitch + base + lane] = tile[r][lane];
        }
        __syncwarp();
    }
}

// copy one colour from `src` to `dst` (brings both colours back into one buffer before the per-pixel
// epilogue kernels / downloads)
__global__ void merge_colour_kernel(int W, int H, int colour, const float4 *__restrict__ psrc,
                                    const float *__restrict__ csrc, float4 *__restrict__ pdst,
                                    float *__restrict__ cdst) {
    const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y * blockDim.y + threadIdx.y;
    if (x >= W || y >= H || ((x + y) & 1) != colour) return;
    const size_t p = (size_t)y * W + x;
    pdst[p] = psrc[p];
    cdst[p] = csrc[p];
}

cudaError_t pm_launch_rng_table(uint32_t *table, int pitch, int H, int len, unsigned long long seed, cudaStream_t s) {
    rng_rows_kernel<<<dim3((H + 31) / 32, (len + kRngSeg - 1) / kRngSeg), 32, 0, s>>>(table, pitch, H, len, seed);
    return cudaGetLastError();
}
cudaError_t pm_launch_merge_colour(int W, int H, int colour, const float4 *psrc, const float *csrc, float4 *pdst,
                                   float *cdst, cudaStream_t s) {
    dim3 b(32, 8), g((W + 31) / 32, (H + 7) / 8);
    merge_colour_kernel<<<g, b, 0, s>>>(W, H, colour, psrc, csrc, pdst, cdst);
    return cudaGetLastError();
}
}  // namespace tsar
