// Forward/backward consistency check of the epipolar variant — the `conf` and `bestD2` outputs the reference's gateway
// allocates (calc_cost_sgm.cpp:571-572) and would fill with the call it ships commented out (:589-590).
//
//   calc_disp_from_first   (calc_cost_sgm.cpp:429-486)  D2 := INVALID; every pixel of image 1 projects to p2 =
//                          (int)(Pd0 - 1 + d*u) (C truncation) and offers its D1 to the 2x2 pixels at p2; a target keeps
//                          the LARGEST offer ("== INVALID || D2 < D1") — order-independent, so the raster loop of the
//                          reference is a scatter with atomicMax here (offers are stored as D1+1, 0 = no offer yet)
//   forward_backward_check (:488-536)                   conf := 1; p2 = round(Pd0 - 1 + d*u); conf := 0 when p2 is outside the
//                          image, D2[p2] is INVALID or |int(D1) - int(D2[p2])| > thr
//   convert_vzInd_to_disp  (:414-426)                   also exported on its own here, because the check runs on the label
//                          map BEFORE that conversion (:589-593)
//
// D1 is the x256 fixed-point label map (bestD before the vz conversion).  fp64 with explicit round-to-nearest ops (the
// reference is built without FMA contraction); double->int follows x86 cvttsd2si (out of range / NaN -> INT_MIN).
#include "fsgm_internal.h"

namespace fsgm {

constexpr uint32_t FB_INVALID = 512u << 8;      // INVALID_DISPARITY, calc_cost_sgm.cpp:5

__device__ __forceinline__ int fb_x86_d2i(double v)
{
    if (!(v > -2147483649.0 && v < 2147483648.0)) return INT_MIN;
    return __double2int_rz(v);
}
__device__ __forceinline__ uint32_t fb_x86_d2u(double v)
{
    if (!(v > -9223372036854775809.0 && v < 9223372036854775808.0)) return 0u;
    return (uint32_t)(unsigned long long)__double2ll_rz(v);
}
__device__ __forceinline__ double fb_round(double v)          // C round(): half away from zero
{
    double t = trunc(v);
    if (fabs(__dsub_rn(v, t)) >= 0.5) t = __dadd_rn(t, copysign(1.0, v));
    return t;
}
// pixel displacement along the epipolar line for a x256 label (:448-456)
__device__ __forceinline__ double fb_disp(uint32_t D1, double O, double vMax, int n, int use_vzind)
{
    double d = __ddiv_rn((double)D1, 256.0);
    if (use_vzind) {
        const double r = __dmul_rn(__ddiv_rn(d, (double)n), vMax);
        d = __dmul_rn(O, __ddiv_rn(r, __dsub_rn(1.0, r)));
    }
    return d;
}

struct FbParams {
    const uint32_t* D1;
    const double *Pd0, *dirn, *O;
    uint32_t* D2;
    uint8_t* conf;
    int W, H, n, thr, use_vzind;
    double vMax;
};

__global__ void fb_scatter_kernel(const FbParams p)
{
    const size_t N = (size_t)p.W * p.H;
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= N) return;
    const size_t g = blockIdx.y * N + i;
    const uint32_t D1 = p.D1[g];
    const double d = fb_disp(D1, p.O[g], p.vMax, p.n, p.use_vzind);
    const double bx = __dsub_rn(p.Pd0[blockIdx.y * 2 * N + i], 1.0), by = __dsub_rn(p.Pd0[blockIdx.y * 2 * N + N + i], 1.0);
    const double ux = p.dirn[blockIdx.y * 2 * N + i], uy = p.dirn[blockIdx.y * 2 * N + N + i];
    const long long p2x = fb_x86_d2i(__dadd_rn(bx, __dmul_rn(d, ux)));
    const long long p2y = fb_x86_d2i(__dadd_rn(by, __dmul_rn(d, uy)));
    uint32_t* D2 = p.D2 + blockIdx.y * N;
#pragma unroll
    for (int dy = 0; dy <= 1; ++dy)
#pragma unroll
        for (int dx = 0; dx <= 1; ++dx) {
            const long long tx = p2x + dx, ty = p2y + dy;
            if (tx >= 0 && tx < p.W && ty >= 0 && ty < p.H) atomicMax(D2 + ty * p.W + tx, D1 + 1u);
        }
}

__global__ void fb_conf_kernel(const FbParams p)
{
    const size_t N = (size_t)p.W * p.H;
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= N) return;
    const size_t g = blockIdx.y * N + i;
    const uint32_t D1 = p.D1[g];
    const double d = fb_disp(D1, p.O[g], p.vMax, p.n, p.use_vzind);
    const double bx = __dsub_rn(p.Pd0[blockIdx.y * 2 * N + i], 1.0), by = __dsub_rn(p.Pd0[blockIdx.y * 2 * N + N + i], 1.0);
    const double ux = p.dirn[blockIdx.y * 2 * N + i], uy = p.dirn[blockIdx.y * 2 * N + N + i];
    const int p2x = fb_x86_d2i(fb_round(__dadd_rn(bx, __dmul_rn(d, ux))));
    const int p2y = fb_x86_d2i(fb_round(__dadd_rn(by, __dmul_rn(d, uy))));
    uint8_t ok = 1;
    if (p2x < 0 || p2x > p.W - 1 || p2y < 0 || p2y > p.H - 1) ok = 0;
    else {
        const uint32_t enc = p.D2[blockIdx.y * N + (size_t)p2y * p.W + p2x];      // offer + 1, 0 = INVALID
        if (enc == 0) ok = 0;
        else if (abs((int)D1 - (int)(enc - 1u)) > p.thr) ok = 0;
    }
    p.conf[g] = ok;
}

__global__ void fb_decode_kernel(uint32_t* __restrict__ D2, size_t total)
{
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= total) return;
    const uint32_t e = D2[i];
    D2[i] = e ? e - 1u : FB_INVALID;
}

__global__ void vz_to_disp_kernel(uint32_t* __restrict__ D, const double* __restrict__ O, size_t total, double vMax, int n)
{
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= total) return;
    const double d = __ddiv_rn((double)D[i], 256.0);
    const double r = __dmul_rn(__ddiv_rn(d, (double)n), vMax);
    D[i] = fb_x86_d2u(__dmul_rn(__dmul_rn(O[i], __ddiv_rn(r, __dsub_rn(1.0, r))), 256.0));
}

int launch_fb_check(fsgm_ctx* c, int n_pairs, const uint32_t* D1, int W, int H, const double* Pd0, const double* dirn, const double* O,
                    double vMax, int n, int thr, int use_vzind, uint8_t* conf, uint32_t* D2)
{
    StageScope ts(c, ST_MISC);
    const size_t N = (size_t)W * H;
    FSGM_CUDA(c, cudaMemsetAsync(D2, 0, n_pairs * N * sizeof(uint32_t), c->stream));
    FbParams p{D1, Pd0, dirn, O, D2, conf, W, H, n, thr, use_vzind, vMax};
    const dim3 grid((unsigned)((N + 255) / 256), n_pairs);
    fb_scatter_kernel<<<grid, 256, 0, c->stream>>>(p);
    FSGM_LAUNCHED(c);
    fb_conf_kernel<<<grid, 256, 0, c->stream>>>(p);
    FSGM_LAUNCHED(c);
    fb_decode_kernel<<<(unsigned)((n_pairs * N + 255) / 256), 256, 0, c->stream>>>(D2, n_pairs * N);
    FSGM_LAUNCHED(c);
    return FSGM_OK;
}

int launch_vz_to_disp(fsgm_ctx* c, uint32_t* D, const double* O, size_t total, double vMax, int n)
{
    StageScope ts(c, ST_MISC);
    vz_to_disp_kernel<<<(unsigned)((total + 255) / 256), 256, 0, c->stream>>>(D, O, total, vMax, n);
    FSGM_LAUNCHED(c);
    return FSGM_OK;
}

}  // namespace fsgm
