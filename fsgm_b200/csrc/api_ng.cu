// C-ABI entry points of the neighbour-guided variants: gateway 3 calc_cost_sgm_ng (reference calc_cost_sgm_ng.cpp:484-527)
// and gateway 4 calc_pyd_cost_sgm_ng (calc_pyd_cost_sgm_ng.cpp:448-523).
#include "fsgm_internal.h"
#include <algorithm>
#include <vector>

using namespace fsgm;

namespace fsgm {
size_t ng_scratch_bytes(int n, int W);
int launch_ng(fsgm_ctx* c, int n, const uint8_t* I1, const uint32_t* cen1, const uint32_t* cen2, int W, int H, int P1, int P2,
              const uint32_t* host_rng_states, const int* d_rand_stream, uint32_t* Sp32, int* Cent, uint32_t* minC, double* flow);
int launch_pydng(fsgm_ctx* c, int n, const uint32_t* cen1, const uint32_t* cen2, int W, int H,
                 const double* preMv, int mvW, int mvH, int r, int agg, int subpixel, int P1, int P2,
                 uint8_t* cost, int* XY, int16_t* const* L, uint32_t* Sp32, uint32_t* minC, double* flow);
}

// glibc's rand()/srand() (TYPE_3 additive feedback generator, degree 31, separation 3): the reference draws its
// random hints from libc rand() (calc_cost_sgm_ng.cpp:148-149), so the stream is part of the result.  This is the
// published algorithm of glibc's random_r.c / srandom_r: LCG seeding (x * 16807 mod 2^31-1), 310 discarded outputs,
// then r[i] = r[i-31] + r[i-3], output r[i] >> 1.
static void glibc_srand_state(unsigned seed, uint32_t st[31])
{
    int32_t word = seed ? (int32_t)seed : 1;
    st[0] = (uint32_t)word;
    for (int i = 1; i < 31; ++i) {
        long hi = word / 127773, lo = word % 127773;
        word = (int32_t)(16807 * lo - 2836 * hi);
        if (word < 0) word += 2147483647;
        st[i] = (uint32_t)word;
    }
    int f = 3, r = 0;
    for (int i = 0; i < 310; ++i) {
        st[f] += st[r];
        f = (f + 1 == 31) ? 0 : f + 1; r = (r + 1 == 31) ? 0 : r + 1;
    }
}

extern "C" {

void fsgm_ng_opts_default(fsgm_ng_opts* o) { if (o) { o->seed = 1; o->rand_stream = nullptr; } }

int fsgm_glibc_rand_fill(unsigned seed, size_t count, int32_t* out)
{
    if (!out) return FSGM_ERR_ARG;
    uint32_t st[31];
    glibc_srand_state(seed, st);
    int f = 3, r = 0;
    for (size_t i = 0; i < count; ++i) {
        st[f] += st[r];
        out[i] = (int32_t)((st[f] >> 1) & 0x7FFFFFFFu);
        f = (f + 1 == 31) ? 0 : f + 1; r = (r + 1 == 31) ? 0 : r + 1;
    }
    return FSGM_OK;
}

int fsgm_calc_cost_sgm_ng_dev(fsgm_ctx* c, int n, const uint8_t* d_I1, const uint8_t* d_I2, int W, int H, int P1, int P2,
                              const unsigned* seeds, const int32_t* d_rand_stream,
                              uint32_t* d_minC, double* d_flow, uint32_t* d_Sp, int32_t* d_Centries)
{
    if (!c) return FSGM_ERR_ARG;
    if (n < 1 || W < 1 || H < 1) return fail(c, FSGM_ERR_ARG, "n_pairs, width and height must be positive");
    if (!d_I1 || !d_I2 || !d_minC || !d_flow) return fail(c, FSGM_ERR_ARG, "null pointer");
    FSGM_CUDA(c, cudaSetDevice(c->device));
    const size_t N = (size_t)W * H;
    FSGM_TRY(arena_reserve(c, 2 * align256(n * N * 4) + ng_scratch_bytes(n, W)));
    ArenaScope scope(c);
    uint32_t *cen1, *cen2;
    FSGM_TRY(arena_get(c, n * N, &cen1));
    FSGM_TRY(arena_get(c, n * N, &cen2));
    FSGM_TRY(launch_census(c, n, d_I1, W, H, cen1));
    FSGM_TRY(launch_census(c, n, d_I2, W, H, cen2));
    std::vector<uint32_t> states((size_t)n * 31);
    for (int i = 0; i < n; ++i) glibc_srand_state(seeds ? seeds[i] : 1u, &states[(size_t)i * 31]);
    return launch_ng(c, n, d_I1, cen1, cen2, W, H, P1, P2, states.data(), d_rand_stream, d_Sp, d_Centries, d_minC, d_flow);
}

int fsgm_calc_cost_sgm_ng(fsgm_ctx* c, const uint8_t* I1, const uint8_t* I2, int W, int H,
                          const double* preMv, double halfSearchWinSize, double aggSize, int subPixelRefine,
                          int P1, int P2, const fsgm_ng_opts* opts, uint32_t* minC, double* flow)
{
    // the reference reads and then ignores these four operands (calc_cost_sgm_ng.cpp:497-503)
    (void)preMv; (void)halfSearchWinSize; (void)aggSize; (void)subPixelRefine;
    if (!c) return FSGM_ERR_ARG;
    if (W < 1 || H < 1) return fail(c, FSGM_ERR_ARG, "width and height must be positive");
    if (!I1 || !I2 || !minC || !flow) return fail(c, FSGM_ERR_ARG, "null pointer");
    FSGM_CUDA(c, cudaSetDevice(c->device));
    fsgm_ng_opts o;
    if (opts) o = *opts; else fsgm_ng_opts_default(&o);
    const size_t N = (size_t)W * H;
    const size_t bytes = 2 * align256(N) + align256(N * 4) + align256(2 * N * 8) + (o.rand_stream ? align256(N * 8 * 4) : 0);
    FSGM_TRY(pipe_reserve(c, bytes));
    FSGM_TRY(fsgm_synchronize(c));               // an earlier enqueue-only call may still own staging slot 0
    char* base = c->pipe.buf[0];
    uint8_t* dI1 = (uint8_t*)base;  base += align256(N);
    uint8_t* dI2 = (uint8_t*)base;  base += align256(N);
    uint32_t* dM = (uint32_t*)base; base += align256(N * 4);
    double* dF = (double*)base;     base += align256(2 * N * 8);
    int32_t* dR = o.rand_stream ? (int32_t*)base : nullptr;
    cudaStream_t s = c->stream;
    FSGM_CUDA(c, cudaMemcpyAsync(dI1, I1, N, cudaMemcpyHostToDevice, s));
    FSGM_CUDA(c, cudaMemcpyAsync(dI2, I2, N, cudaMemcpyHostToDevice, s));
    if (dR) FSGM_CUDA(c, cudaMemcpyAsync(dR, o.rand_stream, N * 8 * 4, cudaMemcpyHostToDevice, s));
    int rc = fsgm_calc_cost_sgm_ng_dev(c, 1, dI1, dI2, W, H, P1, P2, &o.seed, dR, dM, dF, nullptr, nullptr);
    if (rc != FSGM_OK) { cudaStreamSynchronize(s); return rc; }
    FSGM_CUDA(c, cudaMemcpyAsync(minC, dM, N * 4, cudaMemcpyDeviceToHost, s));
    FSGM_CUDA(c, cudaMemcpyAsync(flow, dF, 2 * N * 8, cudaMemcpyDeviceToHost, s));
    FSGM_CUDA(c, cudaStreamSynchronize(s));
    return FSGM_OK;
}

int fsgm_calc_pyd_cost_sgm_ng_dev(fsgm_ctx* c, int n, const uint8_t* d_I1, const uint8_t* d_I2, int W, int H,
                                  const double* d_preMv, int mvW, int mvH, int halfSearchWinSize, int aggSize,
                                  int subPixelRefine, int P1, int P2, uint32_t* d_minC, double* d_flow,
                                  uint32_t* d_Sp, uint8_t* d_cost, int32_t* d_XY)
{
    if (!c) return FSGM_ERR_ARG;
    if (n < 1 || W < 1 || H < 1 || mvW < 1 || mvH < 1) return fail(c, FSGM_ERR_ARG, "sizes must be positive");
    if (!d_I1 || !d_I2 || !d_preMv || !d_minC || !d_flow) return fail(c, FSGM_ERR_ARG, "null pointer");
    const int r = halfSearchWinSize, agg = aggSize / 2;          // calc_pyd_cost_sgm_ng.cpp:488-490
    if (r < 0 || r > 5 || agg < 0 || agg > 4) return fail(c, FSGM_ERR_DOMAIN, "halfSearchWinSize must be 0..5 and aggSize 0..9");
    FSGM_CUDA(c, cudaSetDevice(c->device));
    const size_t N = (size_t)W * H;
    const int S = 2 * r + 1, D = 9 * S * S;
    FSGM_TRY(arena_reserve(c, 2 * align256(n * N * 4) + align256(n * N * D) + align256(n * N * 18 * S * 4) +
                              4 * align256(n * N * D * 2)));
    ArenaScope scope(c);
    uint32_t *cen1, *cen2; uint8_t* cost = d_cost; int* XY = d_XY; int16_t* L[4];
    FSGM_TRY(arena_get(c, n * N, &cen1));
    FSGM_TRY(arena_get(c, n * N, &cen2));
    if (!cost) FSGM_TRY(arena_get(c, n * N * D, &cost));
    if (!XY) FSGM_TRY(arena_get(c, n * N * 18 * S, &XY));
    for (int k = 0; k < 4; ++k) FSGM_TRY(arena_get(c, n * N * D, &L[k]));
    FSGM_TRY(launch_census(c, n, d_I1, W, H, cen1));
    FSGM_TRY(launch_census(c, n, d_I2, W, H, cen2));
    return launch_pydng(c, n, cen1, cen2, W, H, d_preMv, mvW, mvH, r, agg, subPixelRefine, P1, P2, cost, XY, L, d_Sp, d_minC, d_flow);
}

int fsgm_calc_pyd_cost_sgm_ng(fsgm_ctx* c, const uint8_t* I1, const uint8_t* I2, int W, int H,
                              const double* preMv, int mvW, int mvH, int halfSearchWinSize, int aggSize, int subPixelRefine,
                              int P1, int P2, uint32_t* minC, double* flow)
{
    if (!c) return FSGM_ERR_ARG;
    if (W < 1 || H < 1 || mvW < 1 || mvH < 1) return fail(c, FSGM_ERR_ARG, "sizes must be positive");
    if (!I1 || !I2 || !preMv || !minC || !flow) return fail(c, FSGM_ERR_ARG, "null pointer");
    FSGM_CUDA(c, cudaSetDevice(c->device));
    const size_t N = (size_t)W * H, mvN = (size_t)mvW * mvH;
    FSGM_TRY(pipe_reserve(c, 2 * align256(N) + align256(2 * mvN * 8) + align256(N * 4) + align256(2 * N * 8)));
    FSGM_TRY(fsgm_synchronize(c));               // an earlier enqueue-only call may still own staging slot 0
    char* base = c->pipe.buf[0];
    uint8_t* dI1 = (uint8_t*)base;  base += align256(N);
    uint8_t* dI2 = (uint8_t*)base;  base += align256(N);
    double* dMv = (double*)base;    base += align256(2 * mvN * 8);
    uint32_t* dM = (uint32_t*)base; base += align256(N * 4);
    double* dF = (double*)base;
    cudaStream_t s = c->stream;
    FSGM_CUDA(c, cudaMemcpyAsync(dI1, I1, N, cudaMemcpyHostToDevice, s));
    FSGM_CUDA(c, cudaMemcpyAsync(dI2, I2, N, cudaMemcpyHostToDevice, s));
    FSGM_CUDA(c, cudaMemcpyAsync(dMv, preMv, 2 * mvN * 8, cudaMemcpyHostToDevice, s));
    int rc = fsgm_calc_pyd_cost_sgm_ng_dev(c, 1, dI1, dI2, W, H, dMv, mvW, mvH, halfSearchWinSize, aggSize, subPixelRefine,
                                           P1, P2, dM, dF, nullptr, nullptr, nullptr);
    if (rc != FSGM_OK) { cudaStreamSynchronize(s); return rc; }
    FSGM_CUDA(c, cudaMemcpyAsync(minC, dM, N * 4, cudaMemcpyDeviceToHost, s));
    FSGM_CUDA(c, cudaMemcpyAsync(flow, dF, 2 * N * 8, cudaMemcpyDeviceToHost, s));
    FSGM_CUDA(c, cudaStreamSynchronize(s));
    return FSGM_OK;
}

}  // extern "C"
