// 5x5 census transform — replaces census() (reference common.cpp:3-27).
//
// Bit layout (reference common.cpp:12-21): window scanned dy-outer / dx-inner, replicate border,
// bit = (neighbour >= centre); the code is shifted left after every tap including the last, so tap
// k (0..24) lands in bit 25-k, bit 0 is always 0 and the centre tap (k = 12) always sets bit 13.
//
// Roofline: HBM-bound, 1 byte read + 4 bytes written per pixel (5 B/px of the ~6.4 kB/px the whole
// hot path moves) — the image tile is staged in shared memory once so each byte is fetched from
// L2/HBM a single time.
#include "fsgm_internal.h"

namespace fsgm {

constexpr int CEN_TX = 64, CEN_TY = 8, CEN_HALO = 2;

__global__ void __launch_bounds__(CEN_TX * CEN_TY)
census5x5_kernel(const uint8_t* __restrict__ img, uint32_t* __restrict__ cen, int W, int H)
{
    __shared__ uint8_t tile[CEN_TY + 2 * CEN_HALO][CEN_TX + 2 * CEN_HALO + 4];
    const size_t N = (size_t)W * H;
    const uint8_t* src = img + blockIdx.z * N;
    const int x0 = blockIdx.x * CEN_TX, y0 = blockIdx.y * CEN_TY;
    const int tid = threadIdx.y * CEN_TX + threadIdx.x;
    constexpr int TW = CEN_TX + 2 * CEN_HALO, TH = CEN_TY + 2 * CEN_HALO;
    for (int i = tid; i < TW * TH; i += CEN_TX * CEN_TY) {
        int ty = i / TW, tx = i - ty * TW;
        int gy = min(max(y0 + ty - CEN_HALO, 0), H - 1);
        int gx = min(max(x0 + tx - CEN_HALO, 0), W - 1);       // replicate border
        tile[ty][tx] = src[(size_t)gy * W + gx];
    }
    __syncthreads();
    const int x = x0 + threadIdx.x, y = y0 + threadIdx.y;
    if (x >= W || y >= H) return;
    const uint8_t c = tile[threadIdx.y + CEN_HALO][threadIdx.x + CEN_HALO];
    uint32_t code = 0;
#pragma unroll
    for (int dy = 0; dy < 5; ++dy)
#pragma unroll
        for (int dx = 0; dx < 5; ++dx)
            code |= (uint32_t)(tile[threadIdx.y + dy][threadIdx.x + dx] >= c) << (25 - (dy * 5 + dx));
    cen[blockIdx.z * N + (size_t)y * W + x] = code;
}

int launch_census(fsgm_ctx* c, int n_images, const uint8_t* img, int W, int H, uint32_t* cen)
{
    StageScope ss(c, ST_CENSUS);
    dim3 block(CEN_TX, CEN_TY), grid((W + CEN_TX - 1) / CEN_TX, (H + CEN_TY - 1) / CEN_TY, n_images);
    census5x5_kernel<<<grid, block, 0, c->stream>>>(img, cen, W, H);
    FSGM_LAUNCHED(c);
    return FSGM_OK;
}

}  // namespace fsgm
