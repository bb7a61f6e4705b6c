// Dense epipolar prologue / epilogue around calc_cost_sgm — the per-pixel part of the MATLAB driver:
//
//   rotation_motion.m:7-35        Rflow(p) = H*p/(H*p)_3 - p + c*l(1:2),  l = F*p / max(|l(1:2)|, 1e-6 -> 1),  c = -l'*(H*p/(H*p)_3),
//                                 p = (x, y, 1) 0-based
//   epipolar_geometry.m:104-119   PrefD0 = P + Rflow (P 1-based), Direct = PrefD0 - epipole (negated when `direction`),
//                                 Offset = |Direct|, NormlizeDirection = Direct / |Direct|
//   epipolar_sgm_of.m:46-51       flow = (bestD/256) .* NormlizeDirection + Rflow
//
// From F, H, the epipole and the expansion flag (four small host inputs; SURF matching / LMedS / SVD stay on the host) the
// three fp64 maps the gateway wants (40 B per pixel) are produced in HBM instead of crossing PCIe.
// MATLAB evaluates F*P0 and H*P0 through BLAS, whose summation order and FMA use are unspecified, so bit parity with MATLAB
// is unpinned; the order is fixed here as (m1*x + m2*y) + m3 with separately rounded operations, and the numpy restatement
// in oracle/geometry_oracle.py uses the same order, so the two agree bit for bit.
#include "fsgm_internal.h"

namespace fsgm {

struct GeoPair { double F[9], H[9], ex, ey; int direction, pad; };      // row-major 3x3

__device__ __forceinline__ double mat_row(const double* m, double x, double y)
{
    return __dadd_rn(__dadd_rn(__dmul_rn(m[0], x), __dmul_rn(m[1], y)), m[2]);
}

__global__ void geo_prologue_kernel(const GeoPair* __restrict__ gp, int W, int H, double* __restrict__ Pd0, double* __restrict__ dirn,
                                    double* __restrict__ O, double* __restrict__ Rflow)
{
    const size_t N = (size_t)W * H;
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= N) return;
    const GeoPair& g = gp[blockIdx.y];
    const int yi = (int)(i / W), xi = (int)(i - (size_t)yi * W);
    const double x = (double)xi, y = (double)yi;
    // epipolar line of p in image 2, normalised (rotation_motion.m:16, :49-54)
    double l0 = mat_row(g.F, x, y), l1 = mat_row(g.F + 3, x, y), l2 = mat_row(g.F + 6, x, y);
    double nf = sqrt(__dadd_rn(__dmul_rn(l0, l0), __dmul_rn(l1, l1)));
    if (nf < 1e-6) nf = 1.0;
    l0 = __ddiv_rn(l0, nf); l1 = __ddiv_rn(l1, nf); l2 = __ddiv_rn(l2, nf);
    // rotated point (:22-23)
    const double q0 = mat_row(g.H, x, y), q1 = mat_row(g.H + 3, x, y), q2 = mat_row(g.H + 6, x, y);
    const double px = __ddiv_rn(q0, q2), py = __ddiv_rn(q1, q2), pz = __ddiv_rn(q2, q2);
    // back onto the epipolar line (:24-31)
    const double coeff = -__dadd_rn(__dadd_rn(__dmul_rn(l0, px), __dmul_rn(l1, py)), __dmul_rn(l2, pz));
    const double rx = __dadd_rn(__dsub_rn(px, x), __dmul_rn(coeff, l0));
    const double ry = __dadd_rn(__dsub_rn(py, y), __dmul_rn(coeff, l1));
    // epipolar_geometry.m:107-118
    const double p0x = __dadd_rn(__dadd_rn(x, 1.0), rx), p0y = __dadd_rn(__dadd_rn(y, 1.0), ry);
    double dx = __dsub_rn(p0x, g.ex), dy = __dsub_rn(p0y, g.ey);
    if (g.direction) { dx = -dx; dy = -dy; }
    const double len = sqrt(__dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy)));
    const size_t b2 = blockIdx.y * 2 * N + i;
    Pd0[b2] = p0x; Pd0[b2 + N] = p0y;
    dirn[b2] = __ddiv_rn(dx, len); dirn[b2 + N] = __ddiv_rn(dy, len);
    O[blockIdx.y * N + i] = len;
    if (Rflow) { Rflow[b2] = rx; Rflow[b2 + N] = ry; }
}

__global__ void geo_epilogue_kernel(const uint32_t* __restrict__ bestD, const double* __restrict__ dirn, const double* __restrict__ Rflow,
                                    size_t N, double* __restrict__ flow)
{
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= N) return;
    const size_t b2 = blockIdx.y * 2 * N + i;
    const double d = __ddiv_rn((double)bestD[blockIdx.y * N + i], 256.0);
    flow[b2] = __dadd_rn(__dmul_rn(d, dirn[b2]), Rflow[b2]);
    flow[b2 + N] = __dadd_rn(__dmul_rn(d, dirn[b2 + N]), Rflow[b2 + N]);
}

// the same flow rounded to float and interleaved (u, v) per pixel: the layout of the reference's return type, CV_32FC2
// (proj/include/epi_sgm.h:6-12).  8 instead of 16 bytes per pixel on the way back to the host.
__global__ void geo_epilogue_f32_kernel(const uint32_t* __restrict__ bestD, const double* __restrict__ dirn, const double* __restrict__ Rflow,
                                        size_t N, float2* __restrict__ flow)
{
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= N) return;
    const size_t b2 = blockIdx.y * 2 * N + i;
    const double d = __ddiv_rn((double)bestD[blockIdx.y * N + i], 256.0);
    const double u = __dadd_rn(__dmul_rn(d, dirn[b2]), Rflow[b2]);
    const double v = __dadd_rn(__dmul_rn(d, dirn[b2 + N]), Rflow[b2 + N]);
    flow[blockIdx.y * N + i] = make_float2(__double2float_rn(u), __double2float_rn(v));
}

int launch_geo_prologue(fsgm_ctx* c, int n, const double* F, const double* Hm, const double* epi, const int* direction, int W, int H,
                        double* Pd0, double* dirn, double* O, double* Rflow)
{
    StageScope ts(c, ST_GEOMETRY);
    if ((size_t)n > c->geo_cap) {
        if (c->geo_params) { FSGM_CUDA(c, cudaStreamSynchronize(c->stream)); cudaFree(c->geo_params); c->geo_params = nullptr; c->geo_cap = 0; }
        const size_t cap = std::max<size_t>(64, (size_t)n);
        if (cudaMalloc(&c->geo_params, cap * sizeof(GeoPair)) != cudaSuccess) { cudaGetLastError(); return fail(c, FSGM_ERR_NOMEM, "cudaMalloc(geometry parameters)"); }
        c->geo_cap = cap;
    }
    // Pinned staging ring (4 slots of geo_cap pairs, an event per slot): the copy is truly asynchronous, so the call stays
    // enqueue-only, and a slot is rewritten only after the copy that read it has executed.
    if (c->geo_host_cap < c->geo_cap) {
        FSGM_CUDA(c, cudaStreamSynchronize(c->stream));
        if (c->geo_host) cudaFreeHost(c->geo_host);
        c->geo_host = nullptr; c->geo_host_cap = 0;
        if (cudaMallocHost(&c->geo_host, 4 * c->geo_cap * sizeof(GeoPair)) != cudaSuccess) { cudaGetLastError(); return fail(c, FSGM_ERR_NOMEM, "cudaMallocHost(geometry parameters)"); }
        c->geo_host_cap = c->geo_cap;
        for (int k = 0; k < 4; ++k)
            if (!c->geo_ev[k]) FSGM_CUDA(c, cudaEventCreateWithFlags(&c->geo_ev[k], cudaEventDisableTiming));
        c->geo_turn = 0;
    }
    const int slot = c->geo_turn++ & 3;
    if (c->geo_turn > 4) FSGM_CUDA(c, cudaEventSynchronize(c->geo_ev[slot]));
    GeoPair* h = static_cast<GeoPair*>(c->geo_host) + (size_t)slot * c->geo_host_cap;
    for (int i = 0; i < n; ++i) {
        for (int k = 0; k < 9; ++k) { h[i].F[k] = F[i * 9 + k]; h[i].H[k] = Hm[i * 9 + k]; }
        h[i].ex = epi[i * 2]; h[i].ey = epi[i * 2 + 1]; h[i].direction = direction ? direction[i] != 0 : 0; h[i].pad = 0;
    }
    FSGM_CUDA(c, cudaMemcpyAsync(c->geo_params, h, n * sizeof(GeoPair), cudaMemcpyHostToDevice, c->stream));
    FSGM_CUDA(c, cudaEventRecord(c->geo_ev[slot], c->stream));
    const size_t N = (size_t)W * H;
    geo_prologue_kernel<<<dim3((unsigned)((N + 255) / 256), n), 256, 0, c->stream>>>(static_cast<const GeoPair*>(c->geo_params), W, H, Pd0, dirn, O, Rflow);
    FSGM_LAUNCHED(c);
    return FSGM_OK;
}

int launch_geo_epilogue(fsgm_ctx* c, int n, const uint32_t* bestD, const double* dirn, const double* Rflow, int W, int H, double* flow)
{
    StageScope ts(c, ST_GEOMETRY);
    const size_t N = (size_t)W * H;
    geo_epilogue_kernel<<<dim3((unsigned)((N + 255) / 256), n), 256, 0, c->stream>>>(bestD, dirn, Rflow, N, flow);
    FSGM_LAUNCHED(c);
    return FSGM_OK;
}

int launch_geo_epilogue_f32(fsgm_ctx* c, int n, const uint32_t* bestD, const double* dirn, const double* Rflow, int W, int H, float* flow)
{
    StageScope ts(c, ST_GEOMETRY);
    const size_t N = (size_t)W * H;
    geo_epilogue_f32_kernel<<<dim3((unsigned)((N + 255) / 256), n), 256, 0, c->stream>>>(bestD, dirn, Rflow, N, reinterpret_cast<float2*>(flow));
    FSGM_LAUNCHED(c);
    return FSGM_OK;
}

}  // namespace fsgm
