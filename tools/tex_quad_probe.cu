// tex_quad_probe.cu -- what a warp-wide bilinear fetch costs on the B200 texture pipe as a function of how the four
// lanes of a quad are laid out in the texture (instrumentation for DESIGN.md section 4.4; not part of the product).
//
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o build/tex_quad_probe tools/tex_quad_probe.cu && build/tex_quad_probe
//
// Lane l of a warp samples (x0 + sx*l + jx(l), y0 + zy*(l & 1) + jy(l)); every fetch of a thread moves the origin so the
// working set streams through a 3100 x 2050 image like the PatchMatch kernel's source views.  Prints G samples/s and the
// implied pipe cycles per quad (= 148 SMs x clock / quads per second).
#include <cuda_runtime.h>

#include <cstdio>
#include <cstdlib>
#include <vector>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e_)); exit(1); } } while (0)

struct Pattern {
    float sx, zy;       // lane pitch in x, zig-zag in y between even and odd lanes
    float scale;        // multiplies both (a homography that magnifies / minifies)
    float shear;        // x += shear * (l & 1)   (slanted planes shift odd rows)
    int rows_per_warp;  // 1: 32 lanes on one line; 2: lanes 16..31 one "row pair" below (dy = 2)
};

template <bool ROWS2>
__global__ void __launch_bounds__(256) probe(cudaTextureObject_t tex, float *out, int iters, int W, int H, Pattern p) {
    const int lane = threadIdx.x & 31, warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int l = ROWS2 ? (lane & 15) : lane;
    float x = (float)((warp * 37) % (W - 200)) + 0.37f + p.scale * (p.sx * l + p.shear * (l & 1));
    float y = (float)((warp * 11) % (H - 64)) + 0.61f + p.scale * (p.zy * (l & 1) + (ROWS2 ? 2.0f * (lane >> 4) : 0.0f));
    float acc = 0.f;
    for (int i = 0; i < iters; i++) {
        // a 6 x 6 window at stride 2, like the kernel's sampling loop (36 fetches)
#pragma unroll 1
        for (int a = 0; a < 6; a++) {
#pragma unroll
            for (int b = 0; b < 6; b++) acc += tex2D<float>(tex, x + p.scale * 2.0f * a, y + p.scale * 2.0f * b);
        }
        x += 3.13f; y += 0.71f;
        if (x > W - 100) x -= W - 200;
        if (y > H - 32) y -= H - 64;
    }
    if (acc == 12345.678f) out[0] = acc;
}

static cudaTextureObject_t make_tex(int W, int H, bool u8, bool linear, bool pitch2d, std::vector<void *> &keep) {
    cudaResourceDesc rd = {};
    cudaTextureDesc td = {};
    td.addressMode[0] = td.addressMode[1] = cudaAddressModeClamp;
    td.filterMode = linear ? cudaFilterModeLinear : cudaFilterModePoint;
    td.readMode = u8 ? cudaReadModeNormalizedFloat : cudaReadModeElementType;
    td.normalizedCoords = 0;
    const size_t esz = u8 ? 1 : 4;
    std::vector<unsigned char> host(W * (size_t)H * esz);
    for (size_t i = 0; i < host.size(); i++) host[i] = (unsigned char)(i * 2654435761u >> 24);
    if (!u8) { float *f = (float *)host.data(); for (size_t i = 0; i < (size_t)W * H; i++) f[i] = (float)(i % 251); }
    cudaChannelFormatDesc cd = u8 ? cudaCreateChannelDesc<unsigned char>() : cudaCreateChannelDesc<float>();
    if (pitch2d) {
        void *d; size_t pitch;
        CK(cudaMallocPitch(&d, &pitch, W * esz, H));
        CK(cudaMemcpy2D(d, pitch, host.data(), W * esz, W * esz, H, cudaMemcpyHostToDevice));
        rd.resType = cudaResourceTypePitch2D;
        rd.res.pitch2D.devPtr = d; rd.res.pitch2D.desc = cd; rd.res.pitch2D.width = W; rd.res.pitch2D.height = H;
        rd.res.pitch2D.pitchInBytes = pitch;
        keep.push_back(d);
    } else {
        cudaArray_t a;
        CK(cudaMallocArray(&a, &cd, W, H));
        CK(cudaMemcpy2DToArray(a, 0, 0, host.data(), W * esz, W * esz, H, cudaMemcpyHostToDevice));
        rd.resType = cudaResourceTypeArray;
        rd.res.array.array = a;
    }
    cudaTextureObject_t t;
    CK(cudaCreateTextureObject(&t, &rd, &td, nullptr));
    return t;
}

int main() {
    const int W = 3100, H = 2050;
    std::vector<void *> keep;
    int sms = 148, khz = 1965000;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, 0);
    float *out;
    CK(cudaMalloc(&out, 256));
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    struct Fmt { const char *name; bool u8, linear, pitch; } fmts[] = {
        {"u8 linear array", true, true, false}, {"f32 linear array", false, true, false}, {"u8 point array", true, false, false},
        {"u8 linear pitch2D", true, true, true}};
    const Pattern pats[] = {
        // sx  zy  scale shear rows
        {1.f, 0.f, 1.0f, 0.f, 1},   // unit pitch on a line (the peak benchmark's pattern)
        {1.f, 1.f, 1.0f, 0.f, 1},   // the shipped kernel: columns c..c+3, alternating rows (checkerboard row pair)
        {1.f, 1.f, 1.0f, 0.f, 2},   // the same with two row pairs per warp (16 columns x 2)
        {2.f, 0.f, 1.0f, 0.f, 1},   // one image row per warp: same-colour pixels 2 apart
        {1.f, 1.f, 0.9f, 0.f, 1}, {1.f, 1.f, 1.1f, 0.f, 1}, {1.f, 1.f, 1.3f, 0.f, 1}, {1.f, 1.f, 0.7f, 0.f, 1}, {1.f, 1.f, 0.5f, 0.f, 1},
        {1.f, 1.f, 1.0f, 0.5f, 1}, {1.f, 1.f, 1.0f, -0.5f, 1},
        {0.f, 0.f, 1.0f, 0.f, 1},   // all lanes on one point (broadcast)
        {0.5f, 0.f, 1.0f, 0.f, 1}, {1.5f, 0.f, 1.0f, 0.f, 1}, {3.f, 0.f, 1.0f, 0.f, 1}, {4.f, 0.f, 1.0f, 0.f, 1}, {8.f, 0.f, 1.0f, 0.f, 1},
        {1.f, 2.f, 1.0f, 0.f, 1}, {1.f, 3.f, 1.0f, 0.f, 1}, {1.f, 0.5f, 1.0f, 0.f, 1},
    };
    const int blocks = sms * 8, threads = 256, iters = 64;
    for (auto &f : fmts) {
        cudaTextureObject_t tex = make_tex(W, H, f.u8, f.linear, f.pitch, keep);
        printf("== %s\n", f.name);
        for (auto &p : pats) {
            float best = 1e30f;
            for (int rep = 0; rep < 3; rep++) {
                CK(cudaEventRecord(e0));
                if (p.rows_per_warp == 2) probe<true><<<blocks, threads>>>(tex, out, iters, W, H, p);
                else probe<false><<<blocks, threads>>>(tex, out, iters, W, H, p);
                CK(cudaEventRecord(e1));
                CK(cudaEventSynchronize(e1));
                float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
                if (ms < best) best = ms;
            }
            const double samples = (double)blocks * threads * iters * 36;
            const double gs = samples / (best * 1e-3) / 1e9;
            const double cyc_per_quad = (double)sms * khz * 1e3 / (samples / 4.0 / (best * 1e-3));
            printf("sx %.1f zy %.1f scale %.1f shear %+.1f rows %d : %8.1f G samples/s  %.3f pipe cycles per quad\n", p.sx, p.zy, p.scale,
                   p.shear, p.rows_per_warp, gs, cyc_per_quad);
        }
        CK(cudaDestroyTextureObject(tex));
    }
    return 0;
}
