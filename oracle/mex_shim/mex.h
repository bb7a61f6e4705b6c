/* Minimal stand-in for MATLAB's mex.h — TEST INFRASTRUCTURE ONLY.
 *
 * It exists so that the reference's four MEX translation units
 * (/root/reference/calc_*sgm*.cpp) compile unmodified with g++ into
 * oracle/_ref (libref_<variant>.so).  Only the handful of symbols those files touch are
 * provided.  Two deliberate twists for the parity harness:
 *   - mxMalloc logs every allocation and mxFree only *marks* it, so the driver
 *     can read intermediate buffers (census, raw cost, C, Sp) after
 *     mexFunction returns; ref_shim_release() really frees them.
 *   - mxMalloc pads each block by 64 zeroed bytes: the reference reads 4 bytes
 *     past Sp for the last pixel when argmin == D-1 (calc_cost_sgm.cpp:293-296);
 *     padding makes that read deterministic (0) instead of UB-by-heap-layout.
 */
#ifndef FSGM_ORACLE_MEX_SHIM_H
#define FSGM_ORACLE_MEX_SHIM_H
#include <cstddef>
#include <cstdlib>
#include <cstring>
#include <cstdio>
#include <cmath>
#include <vector>
#include <stdexcept>

typedef size_t mwSize;
typedef enum { mxUINT8_CLASS = 9, mxUINT32_CLASS = 13, mxDOUBLE_CLASS = 6 } mxClassID;
typedef enum { mxREAL = 0 } mxComplexity;

struct mxArray {
    void*     data;
    mwSize    m;        /* first dimension                      */
    mwSize    n;        /* product of the remaining dimensions  */
    mxClassID cls;
    bool      owned;
};

struct ShimAlloc { void* p; size_t bytes; bool freed; };
inline std::vector<ShimAlloc>& shim_allocs() { static std::vector<ShimAlloc> v; return v; }

inline void* mxMalloc(size_t bytes) {
    void* p = std::calloc(bytes + 64, 1);
    shim_allocs().push_back({p, bytes, false});
    return p;
}
inline void mxFree(void* p) {
    for (auto& a : shim_allocs()) if (a.p == p) a.freed = true;
}
inline void ref_shim_release() {
    for (auto& a : shim_allocs()) std::free(a.p);
    shim_allocs().clear();
}

inline size_t shim_elem_size(mxClassID c) {
    return c == mxUINT8_CLASS ? 1 : c == mxUINT32_CLASS ? 4 : 8;
}
inline mxArray* mxCreateNumericArray(int ndim, const mwSize* dims, mxClassID cls, mxComplexity) {
    mxArray* a = new mxArray;
    a->m = dims[0];
    a->n = 1;
    for (int i = 1; i < ndim; ++i) a->n *= dims[i];
    a->cls = cls;
    a->data = std::calloc(a->m * a->n * shim_elem_size(cls) + 64, 1);
    a->owned = true;
    return a;
}
inline void   shim_destroy(mxArray* a) { if (a) { if (a->owned) std::free(a->data); delete a; } }
inline void*  mxGetData(const mxArray* a) { return a->data; }
inline double* mxGetPr(const mxArray* a) { return (double*)a->data; }
inline double mxGetScalar(const mxArray* a) { return *(const double*)a->data; }
inline mwSize mxGetM(const mxArray* a) { return a->m; }
inline mwSize mxGetN(const mxArray* a) { return a->n; }

/* Used by the replacement gateways under integration/mex/ (the reference itself signals nothing):
 *   mexErrMsgTxt leaves the MEX function (MATLAB longjmps; here a C++ exception the harness catches);
 *   mexAtExit registers a clean-up MATLAB runs when the MEX file is cleared: the harness runs it from shim_run_atexit(). */
inline void mexErrMsgTxt(const char* msg) { throw std::runtime_error(msg ? msg : "mexErrMsgTxt"); }
inline std::vector<void (*)(void)>& shim_atexit() { static std::vector<void (*)(void)> v; return v; }
inline int mexAtExit(void (*fn)(void)) { shim_atexit().push_back(fn); return 0; }
inline void shim_run_atexit() { for (auto f : shim_atexit()) f(); shim_atexit().clear(); }

#define mxAssert(cond, msg) ((void)0)          /* release-MEX behaviour: compiled out */
inline int mexPrintf(const char*, ...) { return 0; }

#endif
