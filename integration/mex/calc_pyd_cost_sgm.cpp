// MEX gateway stub: drop-in replacement for the reference's calc_pyd_cost_sgm.cpp (gateway at calc_pyd_cost_sgm.cpp:439-510).
//   [bestD, minC, mvSub] = calc_pyd_cost_sgm(I1, I2, preMv, halfSearchWinSizeX, halfSearchWinSizeY, aggHalfWinSize,
//                                            subPixelRefine, P1, P2, enableDiagnalPath, totalPass, adpativeP2)
// Build:  mex calc_pyd_cost_sgm.cpp -I<repo>/include -L<repo>/fsgm_b200 -lfsgm
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
    const mwSize W = mxGetM(prhs[0]), H = mxGetN(prhs[0]);                         // :454-455
    const int mvW = (int)mxGetM(prhs[2]), mvH = (int)mxGetN(prhs[2]) / 2;          // :493-494
    const mwSize d2[2] = { W, H }, d3[3] = { W, H, 2 };
    plhs[0] = mxCreateNumericArray(2, d2, mxUINT32_CLASS, mxREAL);                 // bestD (raw label, :474)
    plhs[1] = mxCreateNumericArray(2, d2, mxUINT32_CLASS, mxREAL);                 // minC  (:475)
    plhs[2] = mxCreateNumericArray(3, d3, mxDOUBLE_CLASS, mxREAL);                 // mvSub (:476)
    const int rc = fsgm_calc_pyd_cost_sgm(g_ctx,
        (const uint8_t*)mxGetData(prhs[0]), (const uint8_t*)mxGetData(prhs[1]), (int)W, (int)H,
        mxGetPr(prhs[2]), mvW, mvH,
        (int)mxGetScalar(prhs[3]), (int)mxGetScalar(prhs[4]), (int)mxGetScalar(prhs[5]),   // halfSearchWinSizeX/Y, aggHalfWinSize
        (int)mxGetScalar(prhs[6]), (int)mxGetScalar(prhs[7]), (int)mxGetScalar(prhs[8]),   // subPixelRefine, P1, P2
        mxGetScalar(prhs[9]) != 0, (int)mxGetScalar(prhs[10]), mxGetScalar(prhs[11]) != 0, // enableDiagnalPath, totalPass, adpativeP2
        (uint32_t*)mxGetData(plhs[0]), (uint32_t*)mxGetData(plhs[1]), mxGetPr(plhs[2]));
    if (rc != FSGM_OK) mexErrMsgTxt(fsgm_last_error(g_ctx));
}
