// MEX gateway stub: drop-in replacement for the reference's calc_cost_sgm.cpp (gateway at calc_cost_sgm.cpp:539-598).
//   [bestD, minC, conf, bestD2] = calc_cost_sgm(I1, I2, dMax, vMax, pixelPosD0, normlizeDirection, offsetFromPosD0, P1, P2)
// Build:  mex calc_cost_sgm.cpp -I<repo>/include -L<repo>/fsgm_b200 -lfsgm
// The reference fixes the number of paths at compile time (enableDiagnalPath, calc_cost_sgm.cpp:104); so does the stub:
// -DFSGM_MEX_PATHS=8 switches the diagonals on, the default is the reference as shipped (4).
#include "mex.h"
#include "fsgm.h"

#ifndef FSGM_MEX_PATHS
#define FSGM_MEX_PATHS 4
#endif

static fsgm_ctx* g_ctx = 0;                      // one context per MATLAB process
static void release_ctx(void) { fsgm_destroy(g_ctx); g_ctx = 0; }

void mexFunction(int nlhs, mxArray* plhs[], int nrhs, const mxArray* prhs[])
{
    (void)nlhs; (void)nrhs;                      // the reference validates nothing either (:539-558)
    if (!g_ctx) {
        if (fsgm_create(0, &g_ctx) != FSGM_OK) mexErrMsgTxt("fsgm: no usable sm_100 device (there is no CPU fallback)");
        mexAtExit(release_ctx);
    }
    const mwSize W = mxGetM(prhs[0]), H = mxGetN(prhs[0]);                         // :562-563
    const mwSize dims[2] = { W, H };
    plhs[0] = mxCreateNumericArray(2, dims, mxUINT32_CLASS, mxREAL);               // bestD   (:569)
    plhs[1] = mxCreateNumericArray(2, dims, mxUINT32_CLASS, mxREAL);               // minC    (:570)
    plhs[2] = mxCreateNumericArray(2, dims, mxUINT8_CLASS, mxREAL);                // conf    (:571, stays zero)
    plhs[3] = mxCreateNumericArray(2, dims, mxUINT32_CLASS, mxREAL);               // bestD2  (:572, stays zero)
    fsgm_epi_opts o;
    fsgm_epi_opts_default(&o);
    o.paths = FSGM_MEX_PATHS;
    const int rc = fsgm_calc_cost_sgm(g_ctx,
        (const uint8_t*)mxGetData(prhs[0]), (const uint8_t*)mxGetData(prhs[1]), (int)W, (int)H,
        (int)mxGetScalar(prhs[2]), mxGetScalar(prhs[3]),                           // dMax, vMax (:551-552)
        mxGetPr(prhs[4]), mxGetPr(prhs[5]), mxGetPr(prhs[6]),                      // pixelPosD0, normlizeDirection, offsetFromPosD0
        (int)mxGetScalar(prhs[7]), (int)mxGetScalar(prhs[8]), &o,                  // P1, P2 (:557-558)
        (uint32_t*)mxGetData(plhs[0]), (uint32_t*)mxGetData(plhs[1]),
        (uint8_t*)mxGetData(plhs[2]), (uint32_t*)mxGetData(plhs[3]));
    if (rc != FSGM_OK) mexErrMsgTxt(fsgm_last_error(g_ctx));
}
