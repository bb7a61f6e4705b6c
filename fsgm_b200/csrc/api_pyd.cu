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

size_t pyd_scratch_bytes(int n, int W, int H, int D)
{
    const size_t N = (size_t)W * H;
    // generic path: C + 8 L volumes of D bytes per pixel; cluster path: 4 padded volumes (<= 176 B per pixel) + records + flags
    return 2 * align256(n * N * 4) + std::max(std::max(9 * align256(n * N * D), 4 * align256(n * N * 176) + align256(n * N * 16) + align256(n * N)),
                                              9 * align256(n * N * 176) + 9 * align256(n * N * 4) + align256((size_t)n * 8) + 256);
}

// The lane = path kernels (pydl.cu): cost volume and per-direction volumes as [y][label column][x][16-byte frame], shift
// descriptors per pixel and direction, one sweep launch for all directions, winner-take-all with lane = pixel.
int pyd_pipeline_lane(fsgm_ctx* c, int n, const uint32_t* cen1, const uint32_t* cen2, const uint8_t* d_I1, int W, int H,
                      const double* d_preMv, int mvW, int mvH, const PydCfg& g, const int* dirs, int nd,
                      uint32_t* d_bestD, uint32_t* d_minC, double* d_mvSub)
{
    const size_t N = (size_t)W * H;
    const int PITCH = 16 * g.Sx;
    uint8_t* C; uint8_t* L[8] = {}; uint32_t* desc[8] = {};
    FSGM_TRY(arena_get(c, n * N * PITCH, &C));
    for (int k = 0; k < nd; ++k) { FSGM_TRY(arena_get(c, n * N * PITCH, &L[k])); FSGM_TRY(arena_get(c, n * N, &desc[k])); }
    if (c->pyd_direct_cost) {
        FSGM_TRY(launch_pyd_cost(c, n, cen1, cen2, W, H, d_preMv, mvW, mvH, g.agg, g.rx, g.ry, C, PITCH, 1));
    } else {
        uint32_t *list, *count;
        FSGM_TRY(arena_get(c, n * N, &list));
        FSGM_TRY(arena_get(c, (size_t)2 * n, &count));
        FSGM_TRY(launch_pyd_cost_sep(c, n, cen1, cen2, W, H, d_preMv, mvW, mvH, g.agg, g.rx, g.ry, C, list, count));
    }
    FSGM_TRY(launch_pydl_desc(c, n, d_preMv, mvW, mvH, W, H, g.Sx, g.Sy, dirs, nd, desc, c->pyd_cluster == -2));
    FSGM_TRY(launch_pydl_sweeps(c, n, C, d_I1, d_preMv, mvW, mvH, W, H, g.Sx, g.Sy, g.P1, g.P2, g.adaptive, dirs, nd, desc, L));
    return launch_pydl_wta(c, n, L, nd, W, H, g.Sx, g.Sy, g.subpixel, d_bestD, d_minC, d_mvSub);
}

// The row-synchronous cluster path (pydv.cu): padded-grid cost volume, the two horizontal directions through the scanline kernel
// (padded in and out), down pass -> byte volume, up pass + WTA -> records, finalize.
int pyd_pipeline_cluster(fsgm_ctx* c, int n, int cs, const uint32_t* cen1, const uint32_t* cen2, const uint8_t* d_I1, int W, int H,
                         const double* d_preMv, int mvW, int mvH, const PydCfg& g, uint32_t* d_bestD, uint32_t* d_minC, double* d_mvSub)
{
    const size_t N = (size_t)W * H;
    const int PITCH = 16 * g.Sx;
    uint8_t *C, *H0, *H1, *S1, *flags; uint4* rec;
    FSGM_TRY(arena_get(c, n * N * PITCH, &C));
    FSGM_TRY(arena_get(c, n * N * PITCH, &H0));
    FSGM_TRY(arena_get(c, n * N * PITCH, &H1));
    FSGM_TRY(arena_get(c, n * N * PITCH, &S1));
    FSGM_TRY(arena_get(c, n * N, &rec));
    FSGM_TRY(arena_get(c, n * N, &flags));
    FSGM_TRY(launch_pyd_cost(c, n, cen1, cen2, W, H, d_preMv, mvW, mvH, g.agg, g.rx, g.ry, C, PITCH));
    FSGM_TRY(launch_pyd_shift_flags(c, n, d_preMv, mvW, mvH, W, H, flags));
    const int hd[2] = {0, 4};
    uint8_t* Lh[2] = {H0, H1};
    FSGM_TRY(launch_pyd_sweeps(c, n, C, d_I1, d_preMv, mvW, mvH, W, H, g.Sx, g.Sy, g.P1, g.P2, 0, hd, 2, Lh, PITCH));
    FSGM_TRY(launch_pydv(c, n, cs, false, C, nullptr, nullptr, nullptr, S1, nullptr, flags, d_preMv, mvW, mvH, W, H, g.Sx, g.Sy, g.P1, g.P2));
    FSGM_TRY(launch_pydv(c, n, cs, true, C, H0, H1, S1, nullptr, rec, flags, d_preMv, mvW, mvH, W, H, g.Sx, g.Sy, g.P1, g.P2));
    return launch_pydv_finalize(c, n, rec, W, H, g.Sx, g.Sy, g.subpixel, d_bestD, d_minC, d_mvSub);
}

// census -> cost -> sweeps -> WTA for one level; the caller has reserved pyd_scratch_bytes() and holds the arena scope
int pyd_pipeline(fsgm_ctx* c, int n, const uint8_t* d_I1, const uint8_t* d_I2, int W, int H, const double* d_preMv, int mvW, int mvH,
                 const PydCfg& g, uint32_t* d_bestD, uint32_t* d_minC, double* d_mvSub)
{
    const size_t N = (size_t)W * H, V = N * g.D;
    ArenaScope scope(c);
    uint32_t *cen1, *cen2; uint8_t* C;
    FSGM_TRY(arena_get(c, n * N, &cen1));
    FSGM_TRY(arena_get(c, n * N, &cen2));
    FSGM_TRY(launch_census(c, n, d_I1, W, H, cen1));
    FSGM_TRY(launch_census(c, n, d_I2, W, H, cen2));
    if ((c->pyd_cluster == 0 || c->pyd_cluster == -2) && (g.agg == 1 || g.agg == 2)) {
        int dirs[8], weights[8];
        const int nd = pyd_dirs(g, dirs, weights);
        if (pydl_applicable(g.Sx, g.Sy, g.P1, g.P2, nd, weights))
            return pyd_pipeline_lane(c, n, cen1, cen2, d_I1, W, H, d_preMv, mvW, mvH, g, dirs, nd, d_bestD, d_minC, d_mvSub);
    }
    if (c->pyd_cluster > 0 && (g.agg == 1 || g.agg == 2) && pydv_applicable(g.Sx, g.Sy, g.P1, g.P2, g.diag, g.passes, g.adaptive)) {
        const int cs = pydv_pick_cluster(c, n, W, g.Sx, c->pyd_cluster);
        if (cs) return pyd_pipeline_cluster(c, n, cs, cen1, cen2, d_I1, W, H, d_preMv, mvW, mvH, g, d_bestD, d_minC, d_mvSub);
    }
    FSGM_TRY(arena_get(c, n * V, &C));
    FSGM_TRY(launch_pyd_cost(c, n, cen1, cen2, W, H, d_preMv, mvW, mvH, g.agg, g.rx, g.ry, C));
    return pyd_aggregate(c, n, C, d_I1, d_preMv, mvW, mvH, W, H, g, nullptr, d_bestD, d_minC, d_mvSub);
}

constexpr int PYR_MAX_LEVELS = 16;

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
    FSGM_TRY(arena_reserve(c, pyd_scratch_bytes(n, W, H, g.D)));
    return pyd_pipeline(c, n, d_I1, d_I2, W, H, d_preMv, mvW, mvH, g, d_bestD, d_minC, d_mvSub);
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
    FSGM_TRY(fsgm_synchronize(c));               // an earlier enqueue-only call may still own staging slot 0
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

// ---------------------------------------------------------------------------------------------------------------------
// N1: the pyramid driver (pyramidal_sgm.m:1-77) with every level resident on the device
// ---------------------------------------------------------------------------------------------------------------------
void fsgm_pyd_opts_default(fsgm_pyd_opts* o)
{
    if (!o) return;
    o->numPyd = 5; o->P1 = 6; o->P2 = 32; o->aggHalfWinSize = 2; o->verSearchHalfWinSize = 5; o->horSearchHalfWinSize = 5;
    o->enableDiagonal = 1; o->totalPass = 2; o->adaptiveP2 = 0;
}

int fsgm_pyramid_dims(int W, int H, int numPyd, int* widths, int* heights)
{
    if (W < 1 || H < 1 || numPyd < 1 || numPyd > PYR_MAX_LEVELS || !widths || !heights) return FSGM_ERR_ARG;
    for (int l = 0; l < numPyd; ++l) {
        widths[l] = W; heights[l] = H;
        W = (W + 1) / 2; H = (H + 1) / 2;
    }
    return FSGM_OK;
}

int fsgm_impyramid_reduce_dev(fsgm_ctx* c, int n_images, const uint8_t* d_img, int W, int H, uint8_t* d_out)
{
    if (!c) return FSGM_ERR_ARG;
    if (n_images < 1 || W < 1 || H < 1 || !d_img || !d_out) return fail(c, FSGM_ERR_ARG, "bad argument");
    FSGM_CUDA(c, cudaSetDevice(c->device));
    return launch_pyr_reduce(c, n_images, d_img, W, H, d_out);
}

int fsgm_pyramidal_sgm_dev(fsgm_ctx* c, int n, const uint8_t* d_I0, const uint8_t* d_I1, int W, int H, const fsgm_pyd_opts* opts,
                           double* d_mv, uint32_t* d_minC, double* d_mvPyd)
{
    if (!c) return FSGM_ERR_ARG;
    fsgm_pyd_opts o;
    if (opts) o = *opts; else fsgm_pyd_opts_default(&o);
    if (o.numPyd < 1 || o.numPyd > PYR_MAX_LEVELS) return fail(c, FSGM_ERR_ARG, "numPyd must be in 1..16");
    if (!d_I0 || !d_I1 || !d_mv || !d_minC) return fail(c, FSGM_ERR_ARG, "null pointer");
    int Ws[PYR_MAX_LEVELS], Hs[PYR_MAX_LEVELS];
    if (fsgm_pyramid_dims(W, H, o.numPyd, Ws, Hs) != FSGM_OK) return fail(c, FSGM_ERR_ARG, "bad size");
    PydCfg g{};
    FSGM_TRY(pyd_check(c, n, W, H, W, H, o.horSearchHalfWinSize, o.verSearchHalfWinSize, o.aggHalfWinSize, o.totalPass, &g));
    g.P1 = o.P1; g.P2 = o.P2; g.diag = o.enableDiagonal != 0; g.adaptive = o.adaptiveP2 != 0;
    FSGM_CUDA(c, cudaSetDevice(c->device));
    const int L = o.numPyd;
    const size_t N0 = (size_t)W * H;
    // driver-owned buffers: both image pyramids above level 0, labels + mvSub + mvCur at the largest level, and the two
    // prior-flow maps (current and next) at up to (W+1) x (H+1)
    size_t own = 0;
    for (int l = 1; l < L; ++l) own += 2 * align256((size_t)n * Ws[l] * Hs[l]);
    own += align256(n * N0 * 4) + 2 * align256(n * 2 * N0 * 8) + 2 * align256((size_t)n * 2 * (W + 1) * (H + 1) * 8);
    FSGM_TRY(arena_reserve(c, own + pyd_scratch_bytes(n, W, H, g.D)));
    ArenaScope scope(c);
    const uint8_t *I0l[PYR_MAX_LEVELS], *I1l[PYR_MAX_LEVELS];
    I0l[0] = d_I0; I1l[0] = d_I1;
    for (int l = 1; l < L; ++l) {                                        // pyramidal_sgm.m:27-30
        uint8_t *a, *b;
        FSGM_TRY(arena_get(c, (size_t)n * Ws[l] * Hs[l], &a));
        FSGM_TRY(arena_get(c, (size_t)n * Ws[l] * Hs[l], &b));
        FSGM_TRY(launch_pyr_reduce(c, n, I0l[l - 1], Ws[l - 1], Hs[l - 1], a));
        FSGM_TRY(launch_pyr_reduce(c, n, I1l[l - 1], Ws[l - 1], Hs[l - 1], b));
        I0l[l] = a; I1l[l] = b;
    }
    uint32_t* label; double *mvSub, *mvCur, *pre[2];
    FSGM_TRY(arena_get(c, n * N0, &label));
    FSGM_TRY(arena_get(c, n * 2 * N0, &mvSub));
    FSGM_TRY(arena_get(c, n * 2 * N0, &mvCur));
    FSGM_TRY(arena_get(c, (size_t)n * 2 * (W + 1) * (H + 1), &pre[0]));
    FSGM_TRY(arena_get(c, (size_t)n * 2 * (W + 1) * (H + 1), &pre[1]));
    int mvW = Ws[L - 1], mvH = Hs[L - 1];
    FSGM_CUDA(c, cudaMemsetAsync(pre[0], 0, (size_t)n * 2 * mvW * mvH * 8, c->stream));   // :33 zero prior at the coarsest level
    size_t pyd_off = 0;
    for (int l = 0; l < L; ++l) pyd_off += (size_t)n * 2 * Ws[l] * Hs[l];                   // d_mvPyd is finest-first
    for (int l = L - 1, k = 0; l >= 0; --l, k ^= 1) {                                        // :36-75
        const int Wl = Ws[l], Hl = Hs[l];
        g.subpixel = l == 0;                                                                 // :48
        FSGM_TRY(pyd_pipeline(c, n, I0l[l], I1l[l], Wl, Hl, pre[k], mvW, mvH, g, label, d_minC, mvSub));
        double* out = l == 0 ? d_mv : mvCur;
        FSGM_TRY(launch_pyr_label_to_mv(c, n, label, pre[k], mvW, mvH, mvSub, Wl, Hl, g.rx, g.ry, out));   // :57-64
        pyd_off -= (size_t)n * 2 * Wl * Hl;
        if (d_mvPyd)
            FSGM_CUDA(c, cudaMemcpyAsync(d_mvPyd + pyd_off, out, (size_t)n * 2 * Wl * Hl * 8, cudaMemcpyDeviceToDevice, c->stream));
        if (l > 0) {                                                                         // :72
            FSGM_TRY(launch_pyr_upsample2(c, n, out, Wl, Hl, pre[k ^ 1]));
            mvW = 2 * Wl; mvH = 2 * Hl;
        }
    }
    return FSGM_OK;
}

int fsgm_pyramidal_sgm(fsgm_ctx* c, const uint8_t* I0, const uint8_t* I1, int W, int H, const fsgm_pyd_opts* opts,
                       double* mv, uint32_t* minC, double* mvPyd)
{
    if (!c) return FSGM_ERR_ARG;
    if (!I0 || !I1 || !mv || !minC || W < 1 || H < 1) return fail(c, FSGM_ERR_ARG, "bad argument");
    fsgm_pyd_opts o;
    if (opts) o = *opts; else fsgm_pyd_opts_default(&o);
    int Ws[PYR_MAX_LEVELS], Hs[PYR_MAX_LEVELS];
    if (fsgm_pyramid_dims(W, H, o.numPyd, Ws, Hs) != FSGM_OK) return fail(c, FSGM_ERR_ARG, "numPyd must be in 1..16");
    FSGM_CUDA(c, cudaSetDevice(c->device));
    FSGM_TRY(fsgm_synchronize(c));
    const size_t N = (size_t)W * H;
    size_t all = 0;
    for (int l = 0; l < o.numPyd; ++l) all += (size_t)2 * Ws[l] * Hs[l];
    FSGM_TRY(pipe_reserve(c, 2 * align256(N) + align256(2 * N * 8) + align256(N * 4) + align256(all * 8)));
    char* base = c->pipe.buf[0];
    uint8_t* dI0 = (uint8_t*)base;  base += align256(N);
    uint8_t* dI1 = (uint8_t*)base;  base += align256(N);
    double* dMv = (double*)base;    base += align256(2 * N * 8);
    uint32_t* dM = (uint32_t*)base; base += align256(N * 4);
    double* dAll = (double*)base;
    cudaStream_t s = c->stream;
    FSGM_CUDA(c, cudaMemcpyAsync(dI0, I0, N, cudaMemcpyHostToDevice, s));
    FSGM_CUDA(c, cudaMemcpyAsync(dI1, I1, N, cudaMemcpyHostToDevice, s));
    int rc = fsgm_pyramidal_sgm_dev(c, 1, dI0, dI1, W, H, &o, dMv, dM, mvPyd ? dAll : nullptr);
    if (rc != FSGM_OK) { cudaStreamSynchronize(s); return rc; }
    FSGM_CUDA(c, cudaMemcpyAsync(mv, dMv, 2 * N * 8, cudaMemcpyDeviceToHost, s));
    FSGM_CUDA(c, cudaMemcpyAsync(minC, dM, N * 4, cudaMemcpyDeviceToHost, s));
    if (mvPyd) FSGM_CUDA(c, cudaMemcpyAsync(mvPyd, dAll, all * 8, cudaMemcpyDeviceToHost, s));
    FSGM_CUDA(c, cudaStreamSynchronize(s));
    return FSGM_OK;
}

}  // extern "C"
