// Pyramid driver kernels — the on-device replacement of the MATLAB loop around calc_pyd_cost_sgm
// (reference pyramidal_sgm.m:24-75).
//
//   impyramid(.,'reduce')  (pyramidal_sgm.m:28-29)  MATLAB Image Processing Toolbox, not part of the reference tree.  Its published
//       algorithm: imresize by 1/2 with the separable Burt-Adelson kernel [1 4 6 4 1]/16 (a = 0.375), output size ceil(n/2),
//       output pixel o centred on input pixel 2o (0-based), symmetric border (the edge pixel is repeated: index -1 -> 0,
//       -2 -> 1, n -> n-1), no antialiasing; imresize resizes dimension 1 (MATLAB rows = image y) first and dimension 2 second,
//       and for uint8 input each pass saturates and rounds to uint8 (half away from zero).  The weights are multiples of 1/16, so
//       both passes are exact in integers: (sum + 8) >> 4.
//   label -> motion vector (pyramidal_sgm.m:57-64)  label = sx*Sy + sy (x-major, calc_pyd_cost_sgm.cpp:81), ind2sub over
//       [Sy, Sx]:  mv = (sx - rx, sy - ry) + preMv(1:rows,1:cols) + mvSub  (added in that order, fp64)
//   2*imresize(mv, 2, 'nearest') (pyramidal_sgm.m:72)  out(y,x) = 2*mv(y/2, x/2), size (2*rows) x (2*cols) — which is why the
//       next level's preMv has its own stride >= the image width.
#include "fsgm_internal.h"

namespace fsgm {

__device__ __forceinline__ int mirror_idx(int i, int n)          // imresize's aux = [1:n, n:-1:1] lookup, 0-based
{
    const int period = 2 * n;
    int m = i % period;
    if (m < 0) m += period;
    return m < n ? m : period - 1 - m;
}

// one thread per output pixel: five vertical 5-tap results (each rounded to u8), then the horizontal 5-tap over them
__global__ void pyr_reduce_kernel(const uint8_t* __restrict__ in, int W, int H, uint8_t* __restrict__ out, int Wo, int Ho)
{
    const int xo = blockIdx.x * blockDim.x + threadIdx.x, yo = blockIdx.y * blockDim.y + threadIdx.y;
    if (xo >= Wo || yo >= Ho) return;
    const uint8_t* img = in + (size_t)blockIdx.z * W * H;
    const int wgt[5] = {1, 4, 6, 4, 1};
    int rows[5];
#pragma unroll
    for (int k = 0; k < 5; ++k) rows[k] = mirror_idx(2 * yo - 2 + k, H) * W;
    unsigned acc = 0;
#pragma unroll
    for (int j = 0; j < 5; ++j) {
        const int x = mirror_idx(2 * xo - 2 + j, W);
        unsigned v = 0;
#pragma unroll
        for (int k = 0; k < 5; ++k) v += wgt[k] * img[rows[k] + x];
        acc += wgt[j] * ((v + 8u) >> 4);
    }
    out[(size_t)blockIdx.z * Wo * Ho + (size_t)yo * Wo + xo] = (uint8_t)((acc + 8u) >> 4);
}

// mvCur = label offset + preMv(1:H,1:W) + mvSub   (planes X, Y; preMv has its own size mvW x mvH)
__global__ void pyr_label_to_mv_kernel(const uint32_t* __restrict__ label, const double* __restrict__ preMv, int mvW, int mvH,
                                       const double* __restrict__ mvSub, int W, int H, int rx, int ry, double* __restrict__ mv)
{
    const size_t N = (size_t)W * H;
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= N) return;
    const int y = (int)(i / W), x = (int)(i - (size_t)y * W);
    const size_t pair = blockIdx.y, mvN = (size_t)mvW * mvH;
    const uint32_t l = label[pair * N + i];
    const int Sy = 2 * ry + 1;
    const int sx = (int)(l / Sy), sy = (int)(l % Sy);
    const double* pm = preMv + pair * 2 * mvN + (size_t)y * mvW + x;
    const double* sb = mvSub + pair * 2 * N + i;
    double* o = mv + pair * 2 * N + i;
    o[0] = __dadd_rn(__dadd_rn((double)(sx - rx), pm[0]), sb[0]);
    o[N] = __dadd_rn(__dadd_rn((double)(sy - ry), pm[mvN]), sb[N]);
}

// preMv(next level) = 2 * nearest-neighbour x2 upsample of mv: [2][2H][2W] from [2][H][W]
__global__ void pyr_upsample2_kernel(const double* __restrict__ mv, int W, int H, double* __restrict__ out)
{
    const int W2 = 2 * W, H2 = 2 * H;
    const size_t N2 = (size_t)W2 * H2, N = (size_t)W * H;
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= N2) return;
    const int y = (int)(i / W2), x = (int)(i - (size_t)y * W2);
    const size_t pp = blockIdx.y;          // pair * 2 + plane
    out[pp * N2 + i] = __dmul_rn(2.0, mv[pp * N + (size_t)(y >> 1) * W + (x >> 1)]);
}

int launch_pyr_reduce(fsgm_ctx* c, int n_images, const uint8_t* in, int W, int H, uint8_t* out)
{
    StageScope ts(c, ST_PYRAMID);
    const int Wo = (W + 1) / 2, Ho = (H + 1) / 2;
    const dim3 block(32, 8), grid((Wo + 31) / 32, (Ho + 7) / 8, n_images);
    pyr_reduce_kernel<<<grid, block, 0, c->stream>>>(in, W, H, out, Wo, Ho);
    FSGM_LAUNCHED(c);
    return FSGM_OK;
}

int launch_pyr_label_to_mv(fsgm_ctx* c, int n, const uint32_t* label, const double* preMv, int mvW, int mvH, const double* mvSub,
                           int W, int H, int rx, int ry, double* mv)
{
    StageScope ts(c, ST_PYRAMID);
    const size_t N = (size_t)W * H;
    pyr_label_to_mv_kernel<<<dim3((unsigned)((N + 255) / 256), n), 256, 0, c->stream>>>(label, preMv, mvW, mvH, mvSub, W, H, rx, ry, mv);
    FSGM_LAUNCHED(c);
    return FSGM_OK;
}

int launch_pyr_upsample2(fsgm_ctx* c, int n, const double* mv, int W, int H, double* out)
{
    StageScope ts(c, ST_PYRAMID);
    const size_t N2 = (size_t)W * H * 4;
    pyr_upsample2_kernel<<<dim3((unsigned)((N2 + 255) / 256), 2 * n), 256, 0, c->stream>>>(mv, W, H, out);
    FSGM_LAUNCHED(c);
    return FSGM_OK;
}

}  // namespace fsgm
