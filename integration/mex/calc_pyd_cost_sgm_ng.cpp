// MEX gateway stub: drop-in replacement for the reference's calc_pyd_cost_sgm_ng.cpp (gateway at calc_pyd_cost_sgm_ng.cpp:448-523).
//   [minC, flow] = calc_pyd_cost_sgm_ng(I1, I2, preMv, halfSearchWinSize, aggSize, subPixelRefine, P1, P2)
// Build:  mex calc_pyd_cost_sgm_ng.cpp -I<repo>/include -L<repo>/fsgm_b200 -lfsgm
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
    const mwSize W = mxGetM(prhs[0]), H = mxGetN(prhs[0]);                         // :462-463
    const int mvW = (int)mxGetM(prhs[2]), mvH = (int)mxGetN(prhs[2]) / 2;          // :501-502
    const mwSize d2[2] = { W, H }, d3[3] = { W, H, 2 };
    plhs[0] = mxCreateNumericArray(2, d2, mxUINT32_CLASS, mxREAL);                 // minC (:476)
    plhs[1] = mxCreateNumericArray(3, d3, mxDOUBLE_CLASS, mxREAL);                 // flow (:477)
    const int rc = fsgm_calc_pyd_cost_sgm_ng(g_ctx,
        (const uint8_t*)mxGetData(prhs[0]), (const uint8_t*)mxGetData(prhs[1]), (int)W, (int)H,
        mxGetPr(prhs[2]), mvW, mvH,
        (int)mxGetScalar(prhs[3]), (int)mxGetScalar(prhs[4]), (int)mxGetScalar(prhs[5]),   // halfSearchWinSize, aggSize, subPixelRefine
        (int)mxGetScalar(prhs[6]), (int)mxGetScalar(prhs[7]),                              // P1, P2
        (uint32_t*)mxGetData(plhs[0]), mxGetPr(plhs[1]));
    if (rc != FSGM_OK) mexErrMsgTxt(fsgm_last_error(g_ctx));
}
