// C-ABI entry points of the pyramidal variant: gateway 2, calc_pyd_cost_sgm (reference calc_pyd_cost_sgm.cpp:439-510).
#include "fsgm_internal.h"
#include <algorithm>

using namespace fsgm;

namespace {

struct PydCfg { int rx, ry, agg, Sx, Sy, D, subpixel, P1, P2, diag, passes, adaptive; };

int pyd_check(fsgm_ctx* c, int n, int W, int H, int mvW, int mvH, int rx, int ry, int agg, int passes, PydCfg* o)
{
    if (!c) return FSGM_ERR_ARG;
    if (n < 1 || W < 1 || H < 1) return fail(c, FSGM_ERR_ARG, "n_pairs, width and height must be positive");
    if (rx < 0 || ry < 0 || agg < 0) return fail(c, FSGM_ERR_ARG, "window radii must be >= 0");
    // the reference indexes preMv with the image's (x, y) and the map's own stride (:388, :213): it must cover the image
    if (mvW < W || mvH < H) return fail(c, FSGM_ERR_ARG, "preMv must be at least as large as the image");
    if ((2 * rx + 1) * (2 * ry + 1) > 1024 || 2 * rx + 1 > 64 || 2 * ry + 1 > 64 || agg > 8)
        return fail(c, FSGM_ERR_DOMAIN, "search window above 1024 labels / 64 per side, or aggregation radius above 8");
    if (passes < 0 || passes > 16) return fail(c, FSGM_ERR_DOMAIN, "totalPass must be in 0..16");
    o->rx = rx; o->ry = ry; o->agg = agg; o->Sx = 2 * rx + 1; o->Sy = 2 * ry + 1; o->D = o->Sx * o->Sy; o->passes = passes;
    return FSGM_OK;
}

// enabled directions and their multiplicity: totalPass is a loop bound in the reference (:142) — pass 0 runs the four
// forward sweeps, every further pass repeats the four reversed sweeps and adds them to Sp again.
int pyd_dirs(const PydCfg& g, int* dirs, int* weights)
{
    int k = 0;
    for (int r = 0; r < 8; ++r) {
        int reps = r < 4 ? (g.passes >= 1) : (g.passes >= 2 ? g.passes - 1 : 0);
        if ((r & 3) >= 2 && !g.diag) reps = 0;
        if (reps) { dirs[k] = r; weights[k] = reps; ++k; }
    }
    return k;
}

int pyd_aggregate(fsgm_ctx* c, int n, const uint8_t* C, const uint8_t* I1, const double* preMv, int mvW, int mvH, int W, int H,
                  const PydCfg& g, uint16_t* Sp16, uint32_t* bestD, uint32_t* minC, double* mvSub)
{
    int dirs[8], weights[8];
    const int nd = pyd_dirs(g, dirs, weights);
    const size_t V = (size_t)W * H * g.D;
    uint8_t* L[8] = {};
    for (int k = 0; k < nd; ++k) FSGM_TRY(arena_get(c, n * V, &L[k]));
    if (nd) FSGM_TRY(launch_pyd_sweeps(c, n, C, I1, preMv, mvW, mvH, W, H, g.Sx, g.Sy, g.P1, g.P2, g.adaptive, dirs, nd, L));
    return launch_pyd_wta(c, n, L, weights, nd, W, H, g.Sx, g.Sy, g.subpixel, Sp16, bestD, minC, mvSub);
}

}  // namespace

extern "C" {

int fsgm_pyd_cost_dev(fsgm_ctx* c, int n, const uint32_t* d_cen1, const uint32_t* d_cen2, int W, int H,
                      const double* d_preMv, int mvW, int mvH, int agg, int rx, int ry, uint8_t* d_C)
{
    PydCfg g{};
    FSGM_TRY(pyd_check(c, n, W, H, mvW, mvH, rx, ry, agg, 2, &g));
    if (!d_cen1 || !d_cen2 || !d_preMv || !d_C) return fail(c, FSGM_ERR_ARG, "null pointer");
    FSGM_CUDA(c, cudaSetDevice(c->device));
    return launch_pyd_cost(c, n, d_cen1, d_cen2, W, H, d_preMv, mvW, mvH, agg, rx, ry, d_C);
}

int fsgm_pyd_sweep_dev(fsgm_ctx* c, int n, const uint8_t* d_C, const uint8_t* d_I1, const double* d_preMv, int mvW, int mvH,
                       int W, int H, int rx, int ry, int P1, int P2, int adaptive, int direction, uint8_t* d_L)
{
    PydCfg g{};
    FSGM_TRY(pyd_check(c, n, W, H, mvW, mvH, rx, ry, 0, 2, &g));
    if (!d_C || !d_I1 || !d_preMv || !d_L) return fail(c, FSGM_ERR_ARG, "null pointer");
    if (direction < 0 || direction > 7) return fail(c, FSGM_ERR_ARG, "direction must be 0..7");
    FSGM_CUDA(c, cudaSetDevice(c->device));
    uint8_t* L[1] = { d_L };
    return launch_pyd_sweeps(c, n, d_C, d_I1, d_preMv, mvW, mvH, W, H, g.Sx, g.Sy, P1, P2, adaptive, &direction, 1, L);
}

int fsgm_pyd_aggregate_dev(fsgm_ctx* c, int n, const uint8_t* d_C, const uint8_t* d_I1, const double* d_preMv, int mvW, int mvH,
                           int W, int H, int rx, int ry, int subPixelRefine, int P1, int P2, int enableDiagnalPath, int totalPass,
                           int adpativeP2, uint16_t* d_Sp, uint32_t* d_bestD, uint32_t* d_minC, double* d_mvSub)
{
    PydCfg g{};
    FSGM_TRY(pyd_check(c, n, W, H, mvW, mvH, rx, ry, 0, totalPass, &g));
    if (!d_C || !d_I1 || !d_preMv || !d_bestD || !d_minC || !d_mvSub) return fail(c, FSGM_ERR_ARG, "null pointer");
    g.subpixel = subPixelRefine; g.P1 = P1; g.P2 = P2; g.diag = enableDiagnalPath != 0; g.adaptive = adpativeP2 != 0;
    FSGM_CUDA(c, cudaSetDevice(c->device));
    FSGM_TRY(arena_reserve(c, 8 * align256((size_t)n * W * H * g.D)));
    ArenaScope scope(c);
    return pyd_aggregate(c, n, d_C, d_I1, d_preMv, mvW, mvH, W, H, g, d_Sp, d_bestD, d_minC, d_mvSub);
}

int fsgm_calc_pyd_cost_sgm_dev(fsgm_ctx* c, int n, const uint8_t* d_I1, const uint8_t* d_I2, int W, int H,
                               const double* d_preMv, int mvW, int mvH, int rx, int ry, int agg, int subPixelRefine,
                               int P1, int P2, int enableDiagnalPath, int totalPass, int adpativeP2,
                               uint32_t* d_bestD, uint32_t* d_minC, double* d_mvSub)
{
    PydCfg g{};
    FSGM_TRY(pyd_check(c, n, W, H, mvW, mvH, rx, ry, agg, totalPass, &g));
    if (!d_I1 || !d_I2 || !d_preMv || !d_bestD || !d_minC || !d_mvSub) return fail(c, FSGM_ERR_ARG, "null pointer");
    g.subpixel = subPixelRefine; g.P1 = P1; g.P2 = P2; g.diag = enableDiagnalPath != 0; g.adaptive = adpativeP2 != 0;
    FSGM_CUDA(c, cudaSetDevice(c->device));
    const size_t N = (size_t)W * H, V = N * g.D;
    FSGM_TRY(arena_reserve(c, 2 * align256(n * N * 4) + 9 * align256(n * V)));
    ArenaScope scope(c);
    uint32_t *cen1, *cen2; uint8_t* C;
    FSGM_TRY(arena_get(c, n * N, &cen1));
    FSGM_TRY(arena_get(c, n * N, &cen2));
    FSGM_TRY(arena_get(c, n * V, &C));
    FSGM_TRY(launch_census(c, n, d_I1, W, H, cen1));
    FSGM_TRY(launch_census(c, n, d_I2, W, H, cen2));
    FSGM_TRY(launch_pyd_cost(c, n, cen1, cen2, W, H, d_preMv, mvW, mvH, agg, rx, ry, C));
    return pyd_aggregate(c, n, C, d_I1, d_preMv, mvW, mvH, W, H, g, nullptr, d_bestD, d_minC, d_mvSub);
}

int fsgm_calc_pyd_cost_sgm(fsgm_ctx* c, const uint8_t* I1, const uint8_t* I2, int W, int H,
                           const double* preMv, int mvW, int mvH, int rx, int ry, int agg, int subPixelRefine,
                           int P1, int P2, int enableDiagnalPath, int totalPass, int adpativeP2,
                           uint32_t* bestD, uint32_t* minC, double* mvSub)
{
    PydCfg g{};
    FSGM_TRY(pyd_check(c, 1, W, H, mvW, mvH, rx, ry, agg, totalPass, &g));
    if (!I1 || !I2 || !preMv || !bestD || !minC || !mvSub) return fail(c, FSGM_ERR_ARG, "null pointer");
    FSGM_CUDA(c, cudaSetDevice(c->device));
    const size_t N = (size_t)W * H, mvN = (size_t)mvW * mvH;
    const size_t bytes = 2 * align256(N) + align256(2 * mvN * 8) + 2 * align256(N * 4) + align256(2 * N * 8);
    FSGM_TRY(pipe_reserve(c, bytes));
    char* base = c->pipe.buf[0];
    uint8_t* dI1 = (uint8_t*)base;  base += align256(N);
    uint8_t* dI2 = (uint8_t*)base;  base += align256(N);
    double* dMv = (double*)base;    base += align256(2 * mvN * 8);
    uint32_t* dB = (uint32_t*)base; base += align256(N * 4);
    uint32_t* dM = (uint32_t*)base; base += align256(N * 4);
    double* dS = (double*)base;
    cudaStream_t s = c->stream;
    FSGM_CUDA(c, cudaMemcpyAsync(dI1, I1, N, cudaMemcpyHostToDevice, s));
    FSGM_CUDA(c, cudaMemcpyAsync(dI2, I2, N, cudaMemcpyHostToDevice, s));
    FSGM_CUDA(c, cudaMemcpyAsync(dMv, preMv, 2 * mvN * 8, cudaMemcpyHostToDevice, s));
    int rc = fsgm_calc_pyd_cost_sgm_dev(c, 1, dI1, dI2, W, H, dMv, mvW, mvH, rx, ry, agg, subPixelRefine, P1, P2,
                                        enableDiagnalPath, totalPass, adpativeP2, dB, dM, dS);
    if (rc != FSGM_OK) { cudaStreamSynchronize(s); return rc; }
    FSGM_CUDA(c, cudaMemcpyAsync(bestD, dB, N * 4, cudaMemcpyDeviceToHost, s));
    FSGM_CUDA(c, cudaMemcpyAsync(minC, dM, N * 4, cudaMemcpyDeviceToHost, s));
    FSGM_CUDA(c, cudaMemcpyAsync(mvSub, dS, 2 * N * 8, cudaMemcpyDeviceToHost, s));
    FSGM_CUDA(c, cudaStreamSynchronize(s));
    return FSGM_OK;
}

}  // extern "C"
