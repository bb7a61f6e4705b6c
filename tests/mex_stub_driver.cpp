/* Harness around the replacement MEX gateways in integration/mex/ — TEST INFRASTRUCTURE ONLY.
 *
 * tests/test_mex_stubs.py compiles one stub (the file a maintainer drops in place of the reference's calc_*sgm*.cpp) together with
 * this driver against oracle/mex_shim/mex.h, exactly as oracle/Makefile compiles the reference's own gateway, and calls the stub's
 * mexFunction with shim mxArrays built from numpy buffers — the same operands, in the same order, as oracle/ref_driver.cpp hands
 * to the reference's mexFunction.  The outputs are then compared with oracle/_ref's.
 */
#include "mex.h"
#include <stdint.h>
#include <string>

void mexFunction(int nlhs, mxArray* plhs[], int nrhs, const mxArray* prhs[]);

namespace {
struct In {
    mxArray a;
    In(const void* p, mwSize m, mwSize n, mxClassID c) { a.data = const_cast<void*>(p); a.m = m; a.n = n; a.cls = c; a.owned = false; }
};
struct Scalar {
    double v; mxArray a;
    explicit Scalar(double x) : v(x) { a.data = &v; a.m = 1; a.n = 1; a.cls = mxDOUBLE_CLASS; a.owned = false; }
};
std::string g_err;
int call(int nlhs, mxArray** plhs, int nrhs, const mxArray** prhs)
{
    try { mexFunction(nlhs, plhs, nrhs, prhs); }
    catch (const std::exception& e) { g_err = e.what(); return -1; }
    return 0;
}
void take(void* dst, mxArray* a, size_t bytes) { if (dst && a) std::memcpy(dst, a->data, bytes); }
}

extern "C" {

const char* stub_last_error(void) { return g_err.c_str(); }
void stub_shutdown(void) { shim_run_atexit(); ref_shim_release(); }

#if defined(STUB_EPI)
int stub_epi(const uint8_t* I1, const uint8_t* I2, int W, int H, int D, double vMax, const double* Pd0, const double* dir,
             const double* O, int P1, int P2, uint32_t* bestD, uint32_t* minC, uint8_t* conf, uint32_t* bestD2)
{
    const size_t N = (size_t)W * H;
    In i1(I1, W, H, mxUINT8_CLASS), i2(I2, W, H, mxUINT8_CLASS);
    In pd(Pd0, W, (mwSize)H * 2, mxDOUBLE_CLASS), dr(dir, W, (mwSize)H * 2, mxDOUBLE_CLASS), of(O, W, H, mxDOUBLE_CLASS);
    Scalar sD(D), sV(vMax), sP1(P1), sP2(P2);
    const mxArray* prhs[9] = { &i1.a, &i2.a, &sD.a, &sV.a, &pd.a, &dr.a, &of.a, &sP1.a, &sP2.a };
    mxArray* plhs[4] = { 0, 0, 0, 0 };
    const int rc = call(4, plhs, 9, prhs);
    if (rc == 0) { take(bestD, plhs[0], N * 4); take(minC, plhs[1], N * 4); take(conf, plhs[2], N); take(bestD2, plhs[3], N * 4); }
    for (int i = 0; i < 4; ++i) shim_destroy(plhs[i]);
    return rc;
}
#elif defined(STUB_PYD)
int stub_pyd(const uint8_t* I1, const uint8_t* I2, int W, int H, const double* preMv, int mvW, int mvH, int rx, int ry, int agg,
             int subpix, int P1, int P2, int diag, int passes, int adaptive, uint32_t* bestD, uint32_t* minC, double* mvSub)
{
    const size_t N = (size_t)W * H;
    In i1(I1, W, H, mxUINT8_CLASS), i2(I2, W, H, mxUINT8_CLASS), mv(preMv, mvW, (mwSize)mvH * 2, mxDOUBLE_CLASS);
    Scalar a3(rx), a4(ry), a5(agg), a6(subpix), a7(P1), a8(P2), a9(diag), a10(passes), a11(adaptive);
    const mxArray* prhs[12] = { &i1.a, &i2.a, &mv.a, &a3.a, &a4.a, &a5.a, &a6.a, &a7.a, &a8.a, &a9.a, &a10.a, &a11.a };
    mxArray* plhs[3] = { 0, 0, 0 };
    const int rc = call(3, plhs, 12, prhs);
    if (rc == 0) { take(bestD, plhs[0], N * 4); take(minC, plhs[1], N * 4); take(mvSub, plhs[2], N * 16); }
    for (int i = 0; i < 3; ++i) shim_destroy(plhs[i]);
    return rc;
}
#elif defined(STUB_NG) || defined(STUB_PYDNG)
/* both neighbour-guided gateways take the same eight operands (calc_cost_sgm_ng.cpp:494-506, calc_pyd_cost_sgm_ng.cpp:458-470) */
int stub_ng(const uint8_t* I1, const uint8_t* I2, int W, int H, const double* preMv, int mvW, int mvH, double halfWin, double aggSize,
            int subpix, int P1, int P2, int seed, uint32_t* minC, double* flow)
{
    const size_t N = (size_t)W * H;
    In i1(I1, W, H, mxUINT8_CLASS), i2(I2, W, H, mxUINT8_CLASS), mv(preMv, mvW, (mwSize)mvH * 2, mxDOUBLE_CLASS);
    Scalar a3(halfWin), a4(aggSize), a5(subpix), a6(P1), a7(P2);
    const mxArray* prhs[8] = { &i1.a, &i2.a, &mv.a, &a3.a, &a4.a, &a5.a, &a6.a, &a7.a };
    mxArray* plhs[2] = { 0, 0 };
    if (seed >= 0) srand((unsigned)seed);          /* seed < 0: keep the process's stream going, as a second call in MATLAB would */
    const int rc = call(2, plhs, 8, prhs);
    if (rc == 0) { take(minC, plhs[0], N * 4); take(flow, plhs[1], N * 16); }
    for (int i = 0; i < 2; ++i) shim_destroy(plhs[i]);
    ref_shim_release();
    return rc;
}
#else
#error "define STUB_EPI, STUB_PYD, STUB_NG or STUB_PYDNG"
#endif

}  /* extern "C" */
