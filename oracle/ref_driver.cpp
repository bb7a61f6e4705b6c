/* Parity-harness driver around the UNMODIFIED reference MEX sources — TEST INFRASTRUCTURE ONLY.
 *
 * This file is never compiled on its own: oracle/Makefile streams one reference
 * translation unit (read in place from /root/reference, never copied into the repo)
 * followed by this driver into g++, with -DREF_VARIANT_{EPI,PYD,NG,PYDNG}.  The driver
 * builds shim mxArrays, calls the reference's own mexFunction and then recovers the
 * intermediate buffers (census, raw cost, C, Sp) from the shim's allocation log
 * (see mex_shim/mex.h) so every stage can be compared, not just the gateway outputs.
 *
 * Nothing under fsgm_b200/ links or loads this.  Only tests/, __graft_entry__.smoke()
 * and bench.py's cpu_baseline / --impl reference leg may.
 */
#include <stdint.h>
#include <chrono>

namespace {
struct In {
    mxArray a;
    In(const void* p, mwSize m, mwSize n, mxClassID c) { a.data = const_cast<void*>(p); a.m = m; a.n = n; a.cls = c; a.owned = false; }
};
struct Scalar {
    double v; mxArray a;
    explicit Scalar(double x) : v(x) { a.data = &v; a.m = 1; a.n = 1; a.cls = mxDOUBLE_CLASS; a.owned = false; }
};
inline void grab(void* dst, size_t idx, size_t bytes) {
    if (!dst) return;
    std::vector<ShimAlloc>& v = shim_allocs();
    if (idx < v.size() && v[idx].bytes == bytes) std::memcpy(dst, v[idx].p, bytes);
    else std::memset(dst, 0xEE, bytes);   /* loud: layout assumption broken */
}
inline double now_s() {
    return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count();
}
}

extern "C" {

#if defined(REF_VARIANT_EPI)
/* calc_cost_sgm.cpp:539 gateway.  Allocation order: C(:579) cen1,cen2(:325-326) Ctmp(:341) L1..L4,Sp(:91-95). */
double ref_epi(const uint8_t* I1, const uint8_t* I2, int W, int H, int D, double vMax,
               const double* Pd0, const double* dir, const double* O, int P1, int P2,
               uint32_t* bestD, uint32_t* minC,
               uint32_t* cen1, uint32_t* cen2, uint8_t* Craw, uint8_t* C, uint32_t* Sp)
{
    size_t N = (size_t)W * H;
    In i1(I1, W, H, mxUINT8_CLASS), i2(I2, W, H, mxUINT8_CLASS);
    In pd(Pd0, W, (mwSize)H * 2, mxDOUBLE_CLASS), dr(dir, W, (mwSize)H * 2, mxDOUBLE_CLASS), of(O, W, H, mxDOUBLE_CLASS);
    Scalar sD(D), sV(vMax), sP1(P1), sP2(P2);
    const mxArray* prhs[9] = { &i1.a, &i2.a, &sD.a, &sV.a, &pd.a, &dr.a, &of.a, &sP1.a, &sP2.a };
    mxArray* plhs[4] = { 0, 0, 0, 0 };
    double t0 = now_s();
    mexFunction(4, plhs, 9, prhs);
    double dt = now_s() - t0;
    if (bestD) std::memcpy(bestD, plhs[0]->data, N * 4);
    if (minC)  std::memcpy(minC,  plhs[1]->data, N * 4);
    grab(C, 0, N * D); grab(cen1, 1, N * 4); grab(cen2, 2, N * 4); grab(Craw, 3, N * D); grab(Sp, 8, N * D * 4);
    for (int i = 0; i < 4; ++i) shim_destroy(plhs[i]);
    ref_shim_release();
    return dt;
}
/* stage entry points (calc_cost_sgm.cpp:319, :86, :414) for stage-isolated timing */
double ref_epi_stage_times(const uint8_t* I1, const uint8_t* I2, int W, int H, int D, double vMax,
                           const double* Pd0, const double* dir, const double* O, int P1, int P2, double* t_cost, double* t_sgm)
{
    size_t N = (size_t)W * H;
    std::vector<unsigned> bd(N), mc(N);
    CostType* C = (CostType*)mxMalloc(N * D);
    double t0 = now_s();
    calc_cost(C, (PixelType*)I1, (PixelType*)I2, W, H, D, vMax, (double*)Pd0, (double*)dir, (double*)O);
    double t1 = now_s();
    sgm(bd.data(), mc.data(), (PixelType*)I1, C, W, H, D, P1, P2, true);
    double t2 = now_s();
    *t_cost = t1 - t0; *t_sgm = t2 - t1;
    ref_shim_release();
    return t2 - t0;
}
void ref_census(const uint8_t* I, uint32_t* cen, int W, int H) { census((PixelType*)I, cen, W, H, 2); }
/* forward_backward_check (calc_cost_sgm.cpp:488) and convert_vzInd_to_disp (:414) called directly */
void ref_fb_check(const uint32_t* D1, int W, int H, const double* Pd0, const double* dir, const double* O,
                  double vMax, int n, int thr, uint8_t* conf, uint32_t* D2)
{
    forward_backward_check(conf, D2, (unsigned*)D1, W, H, (double*)Pd0, (double*)dir, (double*)O, vMax, n, thr);
}
void ref_vz_to_disp(uint32_t* D, int W, int H, const double* O, double vMax, int n)
{
    convert_vzInd_to_disp(D, W, H, (double*)O, vMax, n);
}
/* all four gateway outputs (conf and bestD2 are zeros unless the build re-enables the call at :589-590, see Makefile) */
double ref_epi_all(const uint8_t* I1, const uint8_t* I2, int W, int H, int D, double vMax,
                   const double* Pd0, const double* dir, const double* O, int P1, int P2,
                   uint32_t* bestD, uint32_t* minC, uint8_t* conf, uint32_t* bestD2)
{
    size_t N = (size_t)W * H;
    In i1(I1, W, H, mxUINT8_CLASS), i2(I2, W, H, mxUINT8_CLASS);
    In pd(Pd0, W, (mwSize)H * 2, mxDOUBLE_CLASS), dr(dir, W, (mwSize)H * 2, mxDOUBLE_CLASS), of(O, W, H, mxDOUBLE_CLASS);
    Scalar sD(D), sV(vMax), sP1(P1), sP2(P2);
    const mxArray* prhs[9] = { &i1.a, &i2.a, &sD.a, &sV.a, &pd.a, &dr.a, &of.a, &sP1.a, &sP2.a };
    mxArray* plhs[4] = { 0, 0, 0, 0 };
    double t0 = now_s();
    mexFunction(4, plhs, 9, prhs);
    double dt = now_s() - t0;
    std::memcpy(bestD, plhs[0]->data, N * 4);
    std::memcpy(minC, plhs[1]->data, N * 4);
    std::memcpy(conf, plhs[2]->data, N);
    std::memcpy(bestD2, plhs[3]->data, N * 4);
    for (int i = 0; i < 4; ++i) shim_destroy(plhs[i]);
    ref_shim_release();
    return dt;
}
#endif

#if defined(REF_VARIANT_PYD)
/* calc_pyd_cost_sgm.cpp:439 gateway.  Allocation order: cen1,cen2(:482-483) C(:496) L1..L4,Sp(:121-125). */
double ref_pyd(const uint8_t* I1, const uint8_t* I2, int W, int H,
               const double* preMv, int mvW, int mvH, int rx, int ry, int agg, int subpix,
               int P1, int P2, int diag, int passes, int adaptive,
               uint32_t* bestD, uint32_t* minC, double* mvSub,
               uint32_t* cen1, uint32_t* cen2, uint8_t* C, uint32_t* Sp)
{
    size_t N = (size_t)W * H; int D = (2 * rx + 1) * (2 * ry + 1);
    In i1(I1, W, H, mxUINT8_CLASS), i2(I2, W, H, mxUINT8_CLASS), mv(preMv, mvW, (mwSize)mvH * 2, mxDOUBLE_CLASS);
    Scalar a3(rx), a4(ry), a5(agg), a6(subpix), a7(P1), a8(P2), a9(diag), a10(passes), a11(adaptive);
    const mxArray* prhs[12] = { &i1.a, &i2.a, &mv.a, &a3.a, &a4.a, &a5.a, &a6.a, &a7.a, &a8.a, &a9.a, &a10.a, &a11.a };
    mxArray* plhs[3] = { 0, 0, 0 };
    double t0 = now_s();
    mexFunction(3, plhs, 12, prhs);
    double dt = now_s() - t0;
    if (bestD) std::memcpy(bestD, plhs[0]->data, N * 4);
    if (minC)  std::memcpy(minC,  plhs[1]->data, N * 4);
    if (mvSub) std::memcpy(mvSub, plhs[2]->data, N * 16);
    grab(cen1, 0, N * 4); grab(cen2, 1, N * 4); grab(C, 2, N * D); grab(Sp, 7, N * D * 4);
    for (int i = 0; i < 3; ++i) shim_destroy(plhs[i]);
    ref_shim_release();
    return dt;
}
#endif

#if defined(REF_VARIANT_NG)
/* calc_cost_sgm_ng.cpp:484 gateway.  Allocation order in sgm2d: L1..L4(:197-200) C(:201) Sp(:202) cen1,cen2(:230-231).
 * libc rand() state is global (calc_cost_sgm_ng.cpp:148-149): seed it before every call. */
double ref_ng(const uint8_t* I1, const uint8_t* I2, int W, int H, int P1, int P2, unsigned seed,
              uint32_t* minC, double* flow, int32_t* Centries, uint32_t* Sp)
{
    const int D = DIRECTION_NUM * (::N + ::M) * MV_PER_HINT;   /* the reference's globals N=2, M=1 */
    size_t N = (size_t)W * H;
    std::vector<double> zeros(N, 0.0);
    In i1(I1, W, H, mxUINT8_CLASS), i2(I2, W, H, mxUINT8_CLASS), mv(zeros.data(), W, H, mxDOUBLE_CLASS);
    Scalar a3(1), a4(2), a5(0), a6(P1), a7(P2);
    const mxArray* prhs[8] = { &i1.a, &i2.a, &mv.a, &a3.a, &a4.a, &a5.a, &a6.a, &a7.a };
    mxArray* plhs[2] = { 0, 0 };
    srand(seed);
    double t0 = now_s();
    mexFunction(2, plhs, 8, prhs);
    double dt = now_s() - t0;
    if (minC) std::memcpy(minC, plhs[0]->data, N * 4);
    if (flow) std::memcpy(flow, plhs[1]->data, N * 16);
    grab(Centries, 4, N * D * 12); grab(Sp, 5, N * D * 4);
    for (int i = 0; i < 2; ++i) shim_destroy(plhs[i]);
    ref_shim_release();
    return dt;
}
#endif

#if defined(REF_VARIANT_PYDNG)
/* calc_pyd_cost_sgm_ng.cpp:448 gateway.  Allocation order: cen1,cen2(:482-483) C2(:504) L1..L4,Sp(:106-110). */
double ref_pydng(const uint8_t* I1, const uint8_t* I2, int W, int H,
                 const double* preMv, int mvW, int mvH, int r, int aggSize, int subpix, int P1, int P2,
                 uint32_t* minC, double* flow, int32_t* Centries, uint32_t* Sp)
{
    size_t N = (size_t)W * H; int D = 9 * (2 * r + 1) * (2 * r + 1);
    In i1(I1, W, H, mxUINT8_CLASS), i2(I2, W, H, mxUINT8_CLASS), mv(preMv, mvW, (mwSize)mvH * 2, mxDOUBLE_CLASS);
    Scalar a3(r), a4(aggSize), a5(subpix), a6(P1), a7(P2);
    const mxArray* prhs[8] = { &i1.a, &i2.a, &mv.a, &a3.a, &a4.a, &a5.a, &a6.a, &a7.a };
    mxArray* plhs[2] = { 0, 0 };
    double t0 = now_s();
    mexFunction(2, plhs, 8, prhs);
    double dt = now_s() - t0;
    if (minC) std::memcpy(minC, plhs[0]->data, N * 4);
    if (flow) std::memcpy(flow, plhs[1]->data, N * 16);
    grab(Centries, 2, N * D * 12); grab(Sp, 7, N * D * 4);
    for (int i = 0; i < 2; ++i) shim_destroy(plhs[i]);
    ref_shim_release();
    return dt;
}
#endif

}  /* extern "C" */
