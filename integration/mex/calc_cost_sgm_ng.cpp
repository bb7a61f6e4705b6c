// MEX gateway stub: drop-in replacement for the reference's calc_cost_sgm_ng.cpp (gateway at calc_cost_sgm_ng.cpp:484-527).
//   [minC, flow] = calc_cost_sgm_ng(I1, I2, preMv, halfSearchWinSize, aggSize, subPixelRefine, P1, P2)
// Build:  mex calc_cost_sgm_ng.cpp -I<repo>/include -L<repo>/fsgm_b200 -lfsgm
// The reference draws its hints from libc rand() (:148-149), 8 values per pixel, from whatever state the process is in.  The
// stub reproduces exactly that: it draws the 8*W*H values from the process's rand() here and hands them over as the explicit
// stream, so a sequence of calls sees one continuing stream like the reference's (and MSVC's generator on Windows).
#include <stdlib.h>
#include "mex.h"
#include "fsgm.h"

static fsgm_ctx* g_ctx = 0;
static void release_ctx(void) { fsgm_destroy(g_ctx); g_ctx = 0; }

void mexFunction(int nlhs, mxArray* plhs[], int nrhs, const mxArray* prhs[])
{
    (void)nlhs; (void)nrhs;
    if (!g_ctx) {
        if (fsgm_create(0, &g_ctx) != FSGM_OK) mexErrMsgTxt("fsgm: no usable sm_100 device (there is no CPU fallback)");
        mexAtExit(release_ctx);
    }
    const mwSize W = mxGetM(prhs[0]), H = mxGetN(prhs[0]);                         // :498-499
    const mwSize d2[2] = { W, H }, d3[3] = { W, H, 2 };
    plhs[0] = mxCreateNumericArray(2, d2, mxUINT32_CLASS, mxREAL);                 // minC (:512)
    plhs[1] = mxCreateNumericArray(3, d3, mxDOUBLE_CLASS, mxREAL);                 // flow (:513)
    const size_t count = (size_t)W * H * 8;
    int32_t* stream = (int32_t*)mxMalloc(count * sizeof(int32_t));
    for (size_t i = 0; i < count; ++i) stream[i] = rand();
    fsgm_ng_opts o;
    fsgm_ng_opts_default(&o);
    o.rand_stream = stream;
    const int rc = fsgm_calc_cost_sgm_ng(g_ctx,
        (const uint8_t*)mxGetData(prhs[0]), (const uint8_t*)mxGetData(prhs[1]), (int)W, (int)H,
        mxGetPr(prhs[2]), mxGetScalar(prhs[3]), mxGetScalar(prhs[4]), (int)mxGetScalar(prhs[5]),   // read and ignored (:497-503)
        (int)mxGetScalar(prhs[6]), (int)mxGetScalar(prhs[7]), &o,                                  // P1, P2
        (uint32_t*)mxGetData(plhs[0]), mxGetPr(plhs[1]));
    mxFree(stream);
    if (rc != FSGM_OK) mexErrMsgTxt(fsgm_last_error(g_ctx));
}
