// Pyramidal 2-D-window variant — replaces calc_cost(), sgm_step() and sgm2d() of the reference's
// calc_pyd_cost_sgm.cpp (:374-437, :34-89, :114-372).
//
// Labels are a (2rx+1) x (2ry+1) window of integer offsets around a per-pixel prior flow, ordered
// offx-outer / offy-inner: d = (offx+rx)*Sy + (offy+ry) (:392-393, :81).
//
//  cost   C[p][d] = (u8)(1.0*sum_{agg window} h / winPixels + 0.5), h = popc(cen1[p+a] ^ cen2[q]) or the constant 5
//         when p+a or q is outside the image; q = ((int)(1.0*(offx+x1)+mvx+0.5), (int)(1.0*(offy+y1)+mvy+0.5)) with the
//         prior taken at the CENTRE pixel and C truncation toward zero (:388-389, :415-416).
//  sweep  the 1-D recurrence with a 2-D label neighbourhood: the predecessor label of (sx,sy) is
//         ((int)(sx+ddx+0.5), (int)(sy+ddy+0.5)), (ddx,ddy) = prior(cur) - prior(prev on the path) read with the
//         mv map's own stride (:213-254); same-label term only if inside the window; P1 term = min over the 5x5
//         label neighbourhood, centre excluded, inside-window only (:56-76).  All in unsigned char (mod 256).
//  WTA    first minimum; per-axis parabola if the argmin is not on the window edge (:298-360).
//
// Mapping: sweeps: one warp per scanline exactly as in aggregate.cu (wrapped columns for the six non-horizontal
// directions); the previous L lives in a per-warp padded label grid in shared memory, the 5x5 label neighbourhood is a
// separable 5+5 minimum (10 byte taps per label).  Cost volume: a warp owns 32 consecutive pixels (lane = pixel).
// This path is integer-issue-bound, not HBM-bound.
#include "fsgm_internal.h"

namespace fsgm {

constexpr int PYD_WARPS = 8;
constexpr int PYD_MAXD = 1024;        // (2rx+1)(2ry+1) <= 1024
constexpr int PYD_MAXS = 64;          // window side <= 64

__device__ __forceinline__ int x86_d2i(double v)
{
    if (!(v > -2147483649.0 && v < 2147483648.0)) return INT_MIN;
    return __double2int_rz(v);
}

// ------------------------------------------------------------------------------------------------
// cost volume: one warp per pixel, lanes strided over labels
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
pyd_cost_kernel(const uint32_t* __restrict__ cen1, const uint32_t* __restrict__ cen2, int W, int H,
                const double* __restrict__ preMv, int mvW, int mvH, int agg, int rx, int ry, uint8_t* __restrict__ C)
{
    __shared__ int fx_s[8][PYD_MAXS + 32], fy_s[8][PYD_MAXS + 32];
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    const size_t N = (size_t)W * H;
    const size_t p = (size_t)blockIdx.x * 8 + wib;
    if (p >= N) return;
    const int pair = blockIdx.y;
    const int x = (int)(p % W), y = (int)(p / W);
    const int Sy = 2 * ry + 1, D = (2 * rx + 1) * Sy;
    const double* mvp = preMv + (size_t)pair * 2 * mvW * mvH;
    const double mvx = mvp[(size_t)mvW * y + x], mvy = mvp[(size_t)mvW * mvH + (size_t)mvW * y + x];
    // x2 depends only on s = offx + ax (+x): tabulate it for s in [-(rx+agg), rx+agg]; same for y
    int* fx = fx_s[wib]; int* fy = fy_s[wib];
    for (int s = lane; s < 2 * (rx + agg) + 1; s += 32) {
        int v = x86_d2i(__dadd_rn(__dadd_rn((double)(s - rx - agg + x), mvx), 0.5));
        fx[s] = (v < 0 || v > W - 1) ? -1 : v;
    }
    for (int s = lane; s < 2 * (ry + agg) + 1; s += 32) {
        int v = x86_d2i(__dadd_rn(__dadd_rn((double)(s - ry - agg + y), mvy), 0.5));
        fy[s] = (v < 0 || v > H - 1) ? -1 : v;
    }
    __syncwarp();
    const uint32_t* c1 = cen1 + pair * N;
    const uint32_t* c2 = cen2 + pair * N;
    const int wp = (2 * agg + 1) * (2 * agg + 1);
    uint8_t* out = C + (pair * N + p) * D;
    for (int d = lane; d < D; d += 32) {
        const int ox = d / Sy, oy = d - ox * Sy;           // offx + rx, offy + ry
        uint32_t s = 0;
        for (int ay = -agg; ay <= agg; ++ay) {
            const int y1 = y + ay;
            for (int ax = -agg; ax <= agg; ++ax) {
                const int x1 = x + ax;
                int h = 5;
                if (y1 >= 0 && y1 < H && x1 >= 0 && x1 < W) {
                    const int x2 = fx[ox + ax + agg], y2 = fy[oy + ay + agg];
                    if (x2 >= 0 && y2 >= 0) h = __popc(__ldg(c1 + (size_t)W * y1 + x1) ^ __ldg(c2 + (size_t)W * y2 + x2));
                }
                s += h;
            }
        }
        // (u8)(1.0*s/wp + 0.5): the exact quotient (2s+wp)/(2wp) is at least 1/(2wp) away from the next integer
        // (wp is odd), far more than the fp64 rounding error, so integer floor division is identical
        out[d] = (uint8_t)((2 * s + wp) / (2 * wp));
    }
}

// ------------------------------------------------------------------------------------------------
// sweep
// ------------------------------------------------------------------------------------------------
struct PydSweepParams {
    const uint8_t* C; const uint8_t* I1; const double* preMv;
    uint8_t* L[8]; int dir[8]; int line_start[9]; int n_dirs;
    int W, H, Sx, Sy, mvW, mvH, P1, P2, adaptive;
    int pitch;            // 0: compact volumes (D bytes per pixel); 16*Sx: C and L in the padded grid layout (vector mode only)
};

// FAST: parameters inside the no-wrap domain (P1,P2 >= 0, cmax+P1+P2 <= 255, 2*cmax+P2 <= 255).  There the P1 term may
// include the centre of the 5x5 label neighbourhood (min(Lc, min(Lc, rest)+P1) == min(Lc, rest+P1) for P1 >= 0), which
// makes it a separable 5+5 minimum: R[tx][sy] = min over the y-window of column tx, then a min over the x-window of R.
// 10 shared-memory taps per label instead of 24.  Outside the domain the direct form reproduces every u8 truncation.
// Shared-memory layout: the previous row lives in a PADDED label grid, column pitch P = Sy + 2 with two leading pad bytes per
// column (the two trailing pads of a column are the leading pads of the next) and the y-window minima R in the same grid with
// two more pad columns on either side; pads hold 255, which never beats the far term (M + P2 <= 255 - P1 in the FAST domain).
// (Occupancy is not the limit: capping the kernel at 80 / 64 registers for 3 / 4 blocks per SM gives 1.43 / 1.79 ms against 1.41.)
// When the prior does not change between the two pixels of a step (ddx = ddy = 0: always at the coarsest level, and wherever
// the 2x-upsampled prior is locally constant) the predecessor of label (sx, sy) is (sx, sy) itself and both 5-tap minima
// are five unconditional byte loads at fixed offsets — no coordinate tables, no bounds tests.
template <int NJ, bool FAST>
__global__ void __launch_bounds__(PYD_WARPS * 32)
pyd_sweep_kernel(const PydSweepParams prm)
{
    constexpr int LS_SZ = NJ * 32 + 2 * PYD_MAXS + 32;             // 8 lead + Sx * P + 8 (P = Sy + 2, or 16 in vector mode: Sx <= 16)
    constexpr int RS_SZ = NJ * 32 + 6 * PYD_MAXS + 16;             // (Sx + 4) * (Sy + 2) + 2
    __shared__ __align__(16) uint8_t Ls[PYD_WARPS][2][LS_SZ];
    __shared__ uint8_t Rs[PYD_WARPS][FAST ? RS_SZ : 4];
    __shared__ __align__(16) uint16_t Rv[PYD_WARPS][FAST ? 20 * 16 : 8];   // vector mode: y-window minima, [Sx + 4][16] u16
    __shared__ __align__(16) uint8_t Cp[PYD_WARPS][FAST ? 16 * 16 + 16 : 16];  // vector mode: the cost row in the padded grid
    __shared__ int xt[PYD_WARPS][PYD_MAXS], yt[PYD_WARPS][PYD_MAXS];
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    const int gw = blockIdx.x * PYD_WARPS + wib;
    if (gw >= prm.line_start[prm.n_dirs]) return;
    int k = 0;
    while (gw >= prm.line_start[k + 1]) ++k;
    const int line = gw - prm.line_start[k], r = prm.dir[k];
    const int dx = dir_dx(r), dy = dir_dy(r);
    const int W = prm.W, H = prm.H, Sx = prm.Sx, Sy = prm.Sy, D = Sx * Sy, mvW = prm.mvW;
    // Vector mode (Sy <= 12, Sx <= 16, e.g. the reference's 11 x 11 and BASELINE's 9 x 9 windows): column pitch 16, so a lane
    // owns 8 aligned slots (half a column: two leading pads, Sy labels, trailing pads) and the zero-shift step runs on u16x2
    // registers: y-window from the lane's own 8 bytes + 4 on either side, x-window through a u16 grid in shared memory.
    const bool vec = FAST && Sy <= 12 && Sx <= 16;
    const int P = vec ? 16 : Sy + 2;                               // column pitch of the padded grids
    const size_t N = (size_t)W * H, mvN = (size_t)mvW * prm.mvH;
    const uint8_t* __restrict__ Cb = prm.C + blockIdx.y * N * D;
    uint8_t* __restrict__ Lb = prm.L[k] + blockIdx.y * N * D;
    const uint8_t* __restrict__ Ib = prm.I1 + blockIdx.y * N;
    const double* __restrict__ mvx = prm.preMv + blockIdx.y * 2 * mvN;
    const double* __restrict__ mvy = mvx + mvN;

    int x, y, len;
    if (dy == 0) { y = line; x = dx > 0 ? 0 : W - 1; len = W; }
    else         { x = line; y = dy > 0 ? 0 : H - 1; len = H; }

    // per-lane label geometry (d = lane + 32 j) and the label's place in the padded grid
    int lsx[NJ], lsy[NJ], pidx[NJ];
#pragma unroll
    for (int j = 0; j < NJ; ++j) { int d = lane + 32 * j; lsx[j] = d / Sy; lsy[j] = d - lsx[j] * Sy; pidx[j] = lsx[j] * P + 2 + lsy[j]; }
    for (int i = lane; i < 2 * LS_SZ; i += 32) (&Ls[wib][0][0])[i] = 255;
    if (FAST) {
        for (int i = lane; i < RS_SZ; i += 32) Rs[wib][i] = 255;
        for (int i = lane; i < 20 * 16; i += 32) Rv[wib][i] = 255;
        for (int i = lane; i < 16 * 16 + 16; i += 32) Cp[wib][i] = 255;
    }
    __syncwarp();
    const int vc = lane >> 1, vh = lane & 1;                       // vector mode: column and half of this lane
    const bool vact = vec && vc < Sx;

    uint32_t M = 0;
    int cur = 0, px = 0, py = 0;
    uint8_t cnext[NJ];
    // padded volumes (prm.pitch): a lane's 8 grid slots travel as one 64-bit word in both directions
    const int PT = prm.pitch;
    const uint8_t* __restrict__ Cq = prm.C + blockIdx.y * N * (size_t)PT + vc * 16 + 8 * vh;
    uint8_t* __restrict__ Lq = prm.L[k] + blockIdx.y * N * (size_t)PT + vc * 16 + 8 * vh;
    uint2 cqn = make_uint2(0xFFFFFFFFu, 0xFFFFFFFFu);
    // 32-bit pixel indices, one 64-bit row pointer per step (the per-label offsets lane + 32 j are immediates)
    if (PT) {
        if (vact) cqn = __ldg(reinterpret_cast<const uint2*>(Cq + (size_t)((uint32_t)y * (uint32_t)W + (uint32_t)x) * PT));
    } else {
        const uint8_t* crow = Cb + (size_t)((uint32_t)y * (uint32_t)W + (uint32_t)x) * D + lane;
#pragma unroll
        for (int j = 0; j < NJ; ++j) cnext[j] = lane + 32 * j < D ? __ldg(crow + 32 * j) : 0;
    }
    // the prior at the current pixel travels with the cost row: loaded one step ahead, handed on to the next step as "previous"
    double mxc = mvx[(uint32_t)y * (uint32_t)mvW + (uint32_t)x], myc = mvy[(uint32_t)y * (uint32_t)mvW + (uint32_t)x];
    double mxp = 0.0, myp = 0.0, mxn = 0.0, myn = 0.0;

    for (int t = 0; t < len; ++t) {
        uint8_t c[NJ];
        if (PT) {
            // the padded cost row goes straight into the padded grid; the byte-wise paths read their labels back from there
            if (vact) *reinterpret_cast<uint2*>(&Cp[wib][vc * 16 + 8 * vh]) = cqn;
            __syncwarp();
#pragma unroll
            for (int j = 0; j < NJ; ++j) c[j] = lane + 32 * j < D ? Cp[wib][pidx[j]] : 0;
        } else {
#pragma unroll
            for (int j = 0; j < NJ; ++j) c[j] = cnext[j];
        }
        // next position + prefetch of its cost row and prior
        int nx = x, ny = y;
        if (dy == 0) nx += dx; else { ny += dy; nx += dx; nx = nx < 0 ? W - 1 : (nx >= W ? 0 : nx); }
        if (t + 1 < len) {
            if (PT) {
                if (vact) cqn = __ldg(reinterpret_cast<const uint2*>(Cq + (size_t)((uint32_t)ny * (uint32_t)W + (uint32_t)nx) * PT));
            } else {
                const uint8_t* crow = Cb + (size_t)((uint32_t)ny * (uint32_t)W + (uint32_t)nx) * D + lane;
#pragma unroll
                for (int j = 0; j < NJ; ++j) cnext[j] = lane + 32 * j < D ? __ldg(crow + 32 * j) : 0;
            }
            const uint32_t mi = (uint32_t)ny * (uint32_t)mvW + (uint32_t)nx;
            mxn = mvx[mi]; myn = mvy[mi];
        }
        const uint32_t pix = (uint32_t)y * (uint32_t)W + (uint32_t)x;
        const bool start = (t == 0) || (dy != 0 && dx != 0 && x == (dx > 0 ? 0 : W - 1));
        uint8_t* Lnew = Ls[wib][cur] + 8;                    // grid base: 8 pad bytes in front, 8-byte aligned
        const uint8_t* Lpre = Ls[wib][cur ^ 1] + 8;
        uint32_t m = 255;
        if (start) {
#pragma unroll
            for (int j = 0; j < NJ; ++j) if (lane + 32 * j < D) Lnew[pidx[j]] = c[j];
            M = 0;
            __syncwarp();
        } else {
            int P2 = prm.P2;
            if (prm.adaptive && abs((int)Ib[pix] - (int)Ib[(uint32_t)py * (uint32_t)W + (uint32_t)px]) > 50) P2 = P2 / 8;
            const double ddx = __dsub_rn(mxc, mxp), ddy = __dsub_rn(myc, myp);
            const uint32_t far_ = (M + (uint32_t)P2) & 0xFFu;
            if (FAST && vec && ddx == 0.0 && ddy == 0.0) {
                // predecessor label = the label itself ((int)(s + 0.0 + 0.5) == s); u16x2 registers, 8 slots per lane
                if (!PT) {
#pragma unroll
                    for (int j = 0; j < NJ; ++j) if (lane + 32 * j < D) Cp[wib][pidx[j]] = c[j];
                }
                uint32_t E[6];
                if (vact) {
                    const uint8_t* g = Lpre + vc * 16 + 8 * vh;
                    const uint2 w = *reinterpret_cast<const uint2*>(g);
                    const uint32_t wl = *reinterpret_cast<const uint32_t*>(g - 4), wr = *reinterpret_cast<const uint32_t*>(g + 8);
                    E[0] = __byte_perm(wl, 0, 0x4342);               // slots (-2,-1) of this lane's 8, zero-extended
                    E[1] = __byte_perm(w.x, 0, 0x4140); E[2] = __byte_perm(w.x, 0, 0x4342);
                    E[3] = __byte_perm(w.y, 0, 0x4140); E[4] = __byte_perm(w.y, 0, 0x4342);
                    E[5] = __byte_perm(wr, 0, 0x4140);               // slots (8,9)
                    uint32_t O[5], rr[4];
#pragma unroll
                    for (int i = 0; i < 5; ++i) O[i] = __byte_perm(E[i], E[i + 1], 0x5432);     // odd-aligned pairs (2i-1, 2i)
                    // slots (2i, 2i+1): windows [2i-2, 2i+2] and [2i-1, 2i+3] = min(E_i, O_i, E_i+1, O_i+1, E_i+2) half by half
#pragma unroll
                    for (int i = 0; i < 4; ++i)
                        rr[i] = __vminu2(__vminu2(__vminu2(E[i], O[i]), __vminu2(E[i + 1], O[i + 1])), E[i + 2]);
                    *reinterpret_cast<uint4*>(&Rv[wib][(vc + 2) * 16 + 8 * vh]) = make_uint4(rr[0], rr[1], rr[2], rr[3]);
                }
                __syncwarp();
                uint32_t mm2 = 0xFFFFFFFFu;
                if (vact) {
                    const uint16_t* rq = &Rv[wib][vc * 16 + 8 * vh];                    // column c-2 in the padded R grid
                    const uint4 a0 = *reinterpret_cast<const uint4*>(rq), a1 = *reinterpret_cast<const uint4*>(rq + 16),
                                a2 = *reinterpret_cast<const uint4*>(rq + 32), a3 = *reinterpret_cast<const uint4*>(rq + 48),
                                a4 = *reinterpret_cast<const uint4*>(rq + 64);
                    const uint32_t m5[4] = {
                        __vminu2(__vminu2(__vminu2(a0.x, a1.x), __vminu2(a2.x, a3.x)), a4.x),
                        __vminu2(__vminu2(__vminu2(a0.y, a1.y), __vminu2(a2.y, a3.y)), a4.y),
                        __vminu2(__vminu2(__vminu2(a0.z, a1.z), __vminu2(a2.z, a3.z)), a4.z),
                        __vminu2(__vminu2(__vminu2(a0.w, a1.w), __vminu2(a2.w, a3.w)), a4.w)};
                    const uint2 cw = *reinterpret_cast<const uint2*>(&Cp[wib][vc * 16 + 8 * vh]);
                    const uint32_t cc[4] = {__byte_perm(cw.x, 0, 0x4140), __byte_perm(cw.x, 0, 0x4342),
                                            __byte_perm(cw.y, 0, 0x4140), __byte_perm(cw.y, 0, 0x4342)};
                    const uint32_t far2 = far_ * 0x10001u, P1P1 = (uint32_t)prm.P1 * 0x10001u, MM = M * 0x10001u;
                    uint32_t l[4];
#pragma unroll
                    for (int i = 0; i < 4; ++i) {
                        const uint32_t best = __vminu2(__vminu2(far2, m5[i] + P1P1), E[i + 1]);   // every candidate >= M: no borrow below
                        l[i] = __vminu2(cc[i] + best - MM, 0x00FF00FFu);                          // pad slots (cost 255) saturate back to 255
                        mm2 = __vminu2(mm2, l[i]);
                    }
                    *reinterpret_cast<uint2*>(Lnew + vc * 16 + 8 * vh) =
                        make_uint2(__byte_perm(l[0], l[1], 0x6420), __byte_perm(l[2], l[3], 0x6420));
                }
                m = min(mm2 & 0xFFFFu, mm2 >> 16);
            } else if (FAST && ddx == 0.0 && ddy == 0.0) {
                // predecessor label = the label itself: fixed offsets in the padded grids
                uint8_t* R = Rs[wib] + 2 * P;                        // skip the two pad columns
#pragma unroll
                for (int j = 0; j < NJ; ++j) {
                    if (lane + 32 * j < D) {
                        const uint8_t* q = Lpre + pidx[j];
                        const uint32_t r5 = min(min(min((uint32_t)q[-2], (uint32_t)q[-1]), min((uint32_t)q[0], (uint32_t)q[1])), (uint32_t)q[2]);
                        R[pidx[j]] = (uint8_t)r5;
                    }
                }
                __syncwarp();
#pragma unroll
                for (int j = 0; j < NJ; ++j) {
                    if (lane + 32 * j < D) {
                        const uint8_t* q = R + pidx[j];
                        const uint32_t m5 = min(min(min((uint32_t)q[-2 * P], (uint32_t)q[-P]), min((uint32_t)q[0], (uint32_t)q[P])), (uint32_t)q[2 * P]);
                        const uint32_t best = min(min(far_, m5 + (uint32_t)prm.P1), (uint32_t)Lpre[pidx[j]]);
                        const uint32_t l = c[j] + best - M;
                        Lnew[pidx[j]] = (uint8_t)l;
                        m = min(m, l);
                    }
                }
            } else {
            for (int s = lane; s < Sx; s += 32) {
                int v = x86_d2i(__dadd_rn(__dadd_rn((double)s, ddx), 0.5));
                xt[wib][s] = min(max(v, -8), Sx + 8);          // anything further out behaves the same: no neighbour inside
            }
            for (int s = lane; s < Sy; s += 32) {
                int v = x86_d2i(__dadd_rn(__dadd_rn((double)s, ddy), 0.5));
                yt[wib][s] = min(max(v, -8), Sy + 8);
            }
            __syncwarp();
            if (FAST) {
                uint8_t* R = Rs[wib] + 2 * P;
#pragma unroll
                for (int j = 0; j < NJ; ++j) {                       // pass 1: y-window minimum for cell (tx, sy')
                    const int d = lane + 32 * j;
                    if (d < D) {
                        const int yp = yt[wib][lsy[j]];
                        const uint8_t* col = Lpre + lsx[j] * P + 2;
                        uint32_t r = 255;
#pragma unroll
                        for (int kk = -2; kk <= 2; ++kk) {
                            const int ty = yp + kk;
                            if ((unsigned)ty < (unsigned)Sy) r = min(r, (uint32_t)col[ty]);
                        }
                        R[pidx[j]] = (uint8_t)r;
                    }
                }
                __syncwarp();
#pragma unroll
                for (int j = 0; j < NJ; ++j) {                       // pass 2: x-window minimum of R, then the step
                    const int d = lane + 32 * j;
                    if (d < D) {
                        const int xp = xt[wib][lsx[j]], yp = yt[wib][lsy[j]];
                        uint32_t m5 = 255;
#pragma unroll
                        for (int mm = -2; mm <= 2; ++mm) {
                            const int tx = xp + mm;
                            if ((unsigned)tx < (unsigned)Sx) m5 = min(m5, (uint32_t)R[tx * P + 2 + lsy[j]]);
                        }
                        uint32_t best = min(far_, m5 + (uint32_t)prm.P1);
                        if ((unsigned)xp < (unsigned)Sx && (unsigned)yp < (unsigned)Sy) best = min(best, (uint32_t)Lpre[xp * P + 2 + yp]);
                        const uint32_t l = c[j] + best - M;
                        Lnew[pidx[j]] = (uint8_t)l;
                        m = min(m, l);
                    }
                }
            } else {
#pragma unroll
            for (int j = 0; j < NJ; ++j) {
                const int d = lane + 32 * j;
                if (d < D) {
                    const int xp = xt[wib][lsx[j]], yp = yt[wib][lsy[j]];
                    uint32_t best = far_;
                    for (int mm = -2; mm <= 2; ++mm) {
                        const int tx = xp + mm;
                        if (tx < 0 || tx >= Sx) continue;
                        for (int kk = -2; kk <= 2; ++kk) {
                            const int ty = yp + kk;
                            if (ty < 0 || ty >= Sy) continue;
                            const uint32_t v = Lpre[tx * P + 2 + ty];
                            const uint32_t cand = (mm == 0 && kk == 0) ? v : ((v + (uint32_t)prm.P1) & 0xFFu);
                            best = min(best, cand);
                        }
                    }
                    const uint32_t l = (c[j] + best - M) & 0xFFu;
                    Lnew[pidx[j]] = (uint8_t)l;
                    m = min(m, l);
                }
            }
            }
            }
            M = __reduce_min_sync(0xffffffffu, m);
            __syncwarp();
        }
        // store this pixel's L row (coalesced bytes; padded volumes: the lane's 8 grid slots as one word)
        if (PT) {
            if (vact) *reinterpret_cast<uint2*>(Lq + (size_t)pix * PT) = *reinterpret_cast<const uint2*>(Lnew + vc * 16 + 8 * vh);
        } else {
            uint8_t* lrow = Lb + (size_t)pix * D + lane;
#pragma unroll
            for (int j = 0; j < NJ; ++j) if (lane + 32 * j < D) lrow[32 * j] = Lnew[pidx[j]];
        }
        cur ^= 1;
        px = x; py = y; x = nx; y = ny;
        mxp = mxc; myp = myc; mxc = mxn; myc = myn;
    }
}

// ------------------------------------------------------------------------------------------------
// cost volume, lane = pixel: a warp owns 32 consecutive pixels of a row.  For one label and one window tap the 32 lanes read
// neighbouring census words (the prior is piecewise constant over a warp in practice), so a gather touches 4-5 sectors
// instead of the ~20 of the one-warp-per-pixel kernel above, whose lanes spread over the search window.  Per label column ox
// the T x (2ry+T) reference census words the column needs are gathered ONCE, row by row, into a ring of T rows held in REGISTERS
// (the label loop is unrolled by T, so every ring index is a compile-time constant); the T*T taps of a label are XORs against the
// pixel's own T*T census words + POPC.  (A carry-save adder tree that needs 7 POPC instead of 25 was measured: not faster, 775 against 780 pairs/s — POPC is not this kernel's limit.)  The constant-5 rule (sample or window pixel
// outside the image, calc_pyd_cost_sgm.cpp:405-421) is a per-lane validity word per row / column, all-ones in the interior,
// so the inner loop stays branch-free.  Results go through a shared-memory tile and leave as one contiguous run of bytes.
// ------------------------------------------------------------------------------------------------
constexpr int PYC_WARPS = 4;

template <int AGG>
__global__ void __launch_bounds__(PYC_WARPS * 32)
pyd_cost_px_kernel(const uint32_t* __restrict__ cen1, const uint32_t* __restrict__ cen2, int W, int H,
                   const double* __restrict__ preMv, int mvW, int mvH, int rx, int ry, int pitch, int soa, uint8_t* __restrict__ C,
                   const uint32_t* __restrict__ list, const uint32_t* __restrict__ list_count)
{
    // list != nullptr (soa only): the warp's lanes take 32 entries of the pair's pixel list instead of 32 consecutive pixels
    // pitch == 0: compact label rows (D bytes per pixel, the reference's layout).  pitch == 16*Sx: the PADDED GRID of the
    // row-synchronous aggregation kernels (pydv.cu): column ox at byte 16*ox, two leading pad bytes, Sy labels, trailing pads,
    // every pad byte 255.  soa: the same 16-byte frames as [y][label column][x] (pydl.cu: lane = path aggregation).
    constexpr int T = 2 * AGG + 1, WPX = T * T;
    extern __shared__ __align__(16) unsigned char pyc_smem[];
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    const int Sx = 2 * rx + 1, Sy = 2 * ry + 1, D = Sx * Sy, SX2 = Sx + T - 1, SY2 = Sy + T - 1;
    // per warp: fx[SX2][32], fy[SY2][32] (int), tile[32][D] (u8, padded to 16 bytes)
    const int PT = pitch ? pitch : D;                     // bytes per pixel in the output tile
    const size_t per_warp = (size_t)(SX2 + SY2) * 32 * 4 + (((size_t)32 * PT + 15) & ~(size_t)15);
    unsigned char* base = pyc_smem + wib * per_warp;
    int* fx = reinterpret_cast<int*>(base);
    int* fy = fx + SX2 * 32;
    uint8_t* tile = reinterpret_cast<uint8_t*>(fy + SY2 * 32);

    const int xblocks = (W + 31) / 32;
    const int job = blockIdx.x * PYC_WARPS + wib;
    if (job >= xblocks * H + (list ? 1 : 0)) return;     // warps are independent (only __syncwarp below); the two-ended list rounds up twice
    const size_t N = (size_t)W * H;
    const int pair = blockIdx.y;
    int y = job / xblocks, x0 = (job - y * xblocks) * 32, x = min(x0 + lane, W - 1);      // lanes past the row end repeat its last pixel
    int nlive = min(32, W - x0);
    if (list) {
        // two-ended list (pyd_uniform_kernel): cf entries from the front, cb entries from the back; a warp never mixes them
        const int cf = (int)list_count[2 * pair], cb = (int)list_count[2 * pair + 1];
        const int jf = (cf + 31) / 32;
        if (job < jf) {
            nlive = min(32, cf - job * 32);
            const uint32_t pi = list[(size_t)pair * N + job * 32 + min(lane, nlive - 1)];
            y = (int)(pi / (uint32_t)W); x = (int)(pi - (uint32_t)y * (uint32_t)W);
        } else {
            const int jb = job - jf;
            if (jb * 32 >= cb) return;
            nlive = min(32, cb - jb * 32);
            const uint32_t pi = list[(size_t)pair * N + (N - 1) - (size_t)(jb * 32 + min(lane, nlive - 1))];
            y = (int)(pi / (uint32_t)W); x = (int)(pi - (uint32_t)y * (uint32_t)W);
        }
    }
    const double* mvp = preMv + (size_t)pair * 2 * mvW * mvH;
    const double mvx = mvp[(size_t)mvW * y + x], mvy = mvp[(size_t)mvW * mvH + (size_t)mvW * y + x];
    const uint32_t* c1 = cen1 + pair * N;
    const uint32_t* c2 = cen2 + pair * N;

    if (pitch) {                                          // pad bytes of the grid
        for (int i = lane; i < 32 * PT / 4; i += 32) reinterpret_cast<uint32_t*>(tile)[i] = 0xFFFFFFFFu;
    }
    // sample coordinates: x2 depends only on s = offx + ax (tabulated per lane); same for y.  The tables hold CLAMPED coordinates
    // (y premultiplied by W: a gather's word index is one add) and validity travels as one bit per entry in two registers.
    uint32_t xvalid = 0, yvalid = 0;
    for (int sI = 0; sI < SX2; ++sI) {
        const int v = x86_d2i(__dadd_rn(__dadd_rn((double)(sI - rx - AGG + x), mvx), 0.5));
        const bool ok = !(v < 0 || v > W - 1);
        fx[sI * 32 + lane] = ok ? v : 0;
        xvalid |= (ok ? 1u : 0u) << sI;
    }
    for (int sI = 0; sI < SY2; ++sI) {
        const int v = x86_d2i(__dadd_rn(__dadd_rn((double)(sI - ry - AGG + y), mvy), 0.5));
        const bool ok = !(v < 0 || v > H - 1);
        fy[sI * 32 + lane] = ok ? v * W : 0;
        yvalid |= (ok ? 1u : 0u) << sI;
    }
    // window taps of the current image; a tap outside the image contributes the constant
    uint32_t t1[T][T];
    uint32_t tapok = 0;                                   // bit ay*T+ax
#pragma unroll
    for (int ay = 0; ay < T; ++ay)
#pragma unroll
        for (int ax = 0; ax < T; ++ax) {
            const int x1 = x + ax - AGG, y1 = y + ay - AGG;
            const bool in = x1 >= 0 && x1 < W && y1 >= 0 && y1 < H;
            t1[ay][ax] = in ? __ldg(c1 + (size_t)W * y1 + x1) : 0u;
            tapok |= (in ? 1u : 0u) << (ay * T + ax);
        }
    uint32_t tybits = 0, txbits = 0;                      // window rows / columns inside the image (tapok = their outer product)
#pragma unroll
    for (int a = 0; a < T; ++a) {
        tybits |= (y + a - AGG >= 0 && y + a - AGG < H ? 1u : 0u) << a;
        txbits |= (x + a - AGG >= 0 && x + a - AGG < W ? 1u : 0u) << a;
    }
    __syncwarp();
    const bool rows_ok = tapok == (1u << WPX) - 1u && yvalid == (1u << SY2) - 1u;   // every window tap inside the image and every sample row valid
    asm volatile("" : "+l"(c2));                          // one opaque 64-bit base: a gather's address is IMAD.WIDE.U32(index, 4, base)

    for (int ox = 0; ox < Sx; ++ox) {
        // column validity (bit ax) and the T x SY2 reference words of this label column
        uint32_t fxv[T];
#pragma unroll
        for (int ax = 0; ax < T; ++ax) fxv[ax] = (uint32_t)fx[(ox + ax) * 32 + lane];
        const uint32_t colok = (xvalid >> ox) & ((1u << T) - 1u);
        uint32_t cmw[T];                                  // checked path: all-ones where window column ax and sample column ox + ax are inside
#pragma unroll
        for (int ax = 0; ax < T; ++ax) cmw[ax] = ((colok & txbits) >> ax) & 1u ? 0xFFFFFFFFu : 0u;
        const int ncol = __popc(colok & txbits);
        // Sample rows live in a ring of T rows in registers.  The label loop is unrolled by T, so ring slots are compile-time
        // constants (row sy sits in slot sy % T), and the row a label adds to the window is fetched one label ahead.
        uint32_t nxt[T];
        auto fetch_row = [&](int sy) {
            const uint32_t rb = (uint32_t)fy[sy * 32 + lane];
#pragma unroll
            for (int ax = 0; ax < T; ++ax) nxt[ax] = __ldg(c2 + (rb + fxv[ax]));
        };
        uint32_t vr[T][T];                                // the ring: T sample rows x T columns in registers (all indices static)
        auto commit_row = [&](int slot) {
#pragma unroll
            for (int ax = 0; ax < T; ++ax) vr[slot][ax] = nxt[ax];
        };
#pragma unroll
        for (int sy = 0; sy < T - 1; ++sy) { fetch_row(sy); commit_row(sy); }
        fetch_row(T - 1);
        // every tap of every label of the column is valid for every lane -> no checks (the interior of the image)
        const bool clean = __all_sync(0xffffffffu, rows_ok && colok == (1u << T) - 1u);
        for (int oy0 = 0; oy0 < Sy; oy0 += T) {
#pragma unroll
            for (int u = 0; u < T; ++u) {
                const int oy = oy0 + u;
                if (oy < Sy) {
                    commit_row((u + T - 1) % T);                  // sample row oy + T - 1
                    if (oy + 1 < Sy) fetch_row(oy + T);
                    uint32_t sum = 0;
                    if (clean) {
#pragma unroll
                        for (int ay = 0; ay < T; ++ay)
#pragma unroll
                            for (int ax = 0; ax < T; ++ax) sum += __popc(t1[ay][ax] ^ vr[(u + ay) % T][ax]);
                    } else {
                        // validity is separable — a tap counts iff its window row, its sample row, its window column and its sample
                        // column are inside: invalid columns are masked out of the XOR (one LOP3), invalid rows drop their row sum, and
                        // every invalid tap adds the constant once
                        const uint32_t rowbits = (yvalid >> oy) & tybits;
#pragma unroll
                        for (int ay = 0; ay < T; ++ay) {
                            uint32_t rs = 0;
#pragma unroll
                            for (int ax = 0; ax < T; ++ax) rs += __popc((t1[ay][ax] ^ vr[(u + ay) % T][ax]) & cmw[ax]);
                            sum += ((rowbits >> ay) & 1u) ? rs : 0u;
                        }
                        sum += 5u * (uint32_t)(WPX - __popc(rowbits) * ncol);
                    }
                    // (u8)(1.0*s/wp + 0.5) == (2s + wp) / (2wp) in integers (see pyd_cost_kernel)
                    tile[lane * PT + (pitch ? ox * 16 + 2 + oy : ox * Sy + oy)] = (uint8_t)((2 * sum + WPX) / (2 * WPX));
                }
            }
        }
        __syncwarp();
    }
    // the warp's pixels are consecutive and so are their label rows: one contiguous run of nlive * D bytes
    uint8_t* out = C + (pair * N + (size_t)y * W + x0) * PT;
    const int nbytes = nlive * PT;
    if (soa) {                                            // one 512-byte run per label column
        uint8_t* o2 = C + (size_t)pair * N * PT + (((size_t)y * Sx) * W + x) * 16;
        if (lane < nlive)
            for (int ox = 0; ox < Sx; ++ox)
                *reinterpret_cast<uint4*>(o2 + (size_t)ox * W * 16) = *reinterpret_cast<const uint4*>(tile + lane * PT + ox * 16);
    } else if (pitch) {                                   // 16-byte aligned on both sides
        for (int i = lane; i < nbytes / 16; i += 32) reinterpret_cast<uint4*>(out)[i] = reinterpret_cast<const uint4*>(tile)[i];
    } else {
        for (int i = lane; i < nbytes; i += 32) out[i] = tile[i];
    }
}


// ------------------------------------------------------------------------------------------------
// cost volume where the prior is locally constant: a separable box filter over SHARED raw costs.
// The reference samples every pixel q of p's aggregation window with the prior of the CENTRE p (:388-389); when all pixels of
// the window carry bitwise the same prior, that is the pixel's own prior and h(q, o) = popc(cen1[q] ^ cen2[q + o + mv(q)])
// (or the constant 5) does not depend on p any more: cost(p, o) is a T x T box sum of one raw volume, T*T times fewer
// Hamming distances.  This kernel computes that for EVERY pixel; pyd_uniform_kernel lists the pixels whose window is not
// uniform and pyd_cost_px_kernel (list mode) recomputes exactly those.
// A thread owns one image column of a 128-column strip (T - 1 of them halo) and marches down a segment of rows: per row the raw
// costs of its pixel (three words per label column: rows 0..11 as bytes), a horizontal T-sum through shared memory (bytes:
// <= 5 * 25), a vertical running sum in u16x2 registers with the row that leaves the window read back from a ring of T rows.
// Output in the [y][label column][x][16-byte frame] layout of pydl.cu.
// ------------------------------------------------------------------------------------------------
constexpr int PCS_COLS = 128;            // image columns per block (T - 1 of them halo)
constexpr int PCS_ROWS = 32;             // rows per segment (the T - 1 halo rows are recomputed per segment)

// G threads share an image column: thread (column, g) owns label columns [g*OXG, (g+1)*OXG) — the per-column shared-memory
// footprint (ring of T rows) caps the columns in flight per SM, the split multiplies the warps that work on them.
template <int AGG, int SX, int G>
__global__ void __launch_bounds__(PCS_COLS * G, 2)
pyd_cost_sep_kernel(const uint32_t* __restrict__ cen1, const uint32_t* __restrict__ cen2, int W, int H,
                    const double* __restrict__ preMv, int mvW, int mvH, int ry, uint8_t* __restrict__ C)
{
    constexpr int T = 2 * AGG + 1, WPX = T * T, NW = 3 * SX, RX = (SX - 1) / 2, OXG = (SX + G - 1) / G;
    static_assert(WPX == 25 || WPX == 9, "normalisation constants");
    extern __shared__ uint32_t pcs_smem[];
    uint32_t* rawb = pcs_smem;                                           // [NW][128]
    uint32_t* ring = pcs_smem + NW * PCS_COLS;                           // [T][NW][128]
    const int tid = threadIdx.x % PCS_COLS, g = threadIdx.x / PCS_COLS;
    const int ox0 = g * OXG;
    const int Sy = 2 * ry + 1;
    const int TW = PCS_COLS - 2 * AGG;
    const int xq = blockIdx.x * TW - AGG + tid;
    const int y0 = blockIdx.y * PCS_ROWS, y1 = min(y0 + PCS_ROWS, H);
    const int pair = blockIdx.z;
    const size_t N = (size_t)W * H;
    const uint32_t* c1 = cen1 + pair * N;
    const uint32_t* c2 = cen2 + pair * N;
    asm volatile("" : "+l"(c2));                          // one opaque 64-bit base: a gather's address is IMAD.WIDE.U32(index, 4, base)
    const double* mvxp = preMv + (size_t)pair * 2 * mvW * mvH;
    const double* mvyp = mvxp + (size_t)mvW * mvH;
    uint8_t* Cb = C + (size_t)pair * N * (SX * 16);
    const bool col_in = xq >= 0 && xq < W;
    const bool interior = tid >= AGG && tid < PCS_COLS - AGG && col_in;

    for (int i = threadIdx.x; i < T * NW * PCS_COLS; i += PCS_COLS * G) ring[i] = 0;
    uint32_t vs[OXG][6];
#pragma unroll
    for (int o = 0; o < OXG; ++o)
#pragma unroll
        for (int i = 0; i < 6; ++i) vs[o][i] = 0;
    uint32_t padw[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        uint32_t m = 0;
#pragma unroll
        for (int b = 0; b < 4; ++b) { const int row = 4 * j + b - 2; if (row < 0 || row >= Sy) m |= 0xFFu << (8 * b); }
        padw[j] = m;
    }
    __syncthreads();

    int slot = 0;
    for (int yq = y0 - AGG; yq < y1 + AGG; ++yq) {
        // ---- raw costs of pixel (xq, yq), this thread's label columns ----------------------------------------------------------
        const bool inimg = col_in && yq >= 0 && yq < H;
        int fx[OXG], fy[12];
        uint32_t w1 = 0;
        bool allok = true;
        if (inimg) {
            const double mvx = mvxp[(size_t)mvW * yq + xq], mvy = mvyp[(size_t)mvW * yq + xq];
            w1 = __ldg(c1 + (size_t)W * yq + xq);
#pragma unroll
            for (int o = 0; o < OXG; ++o) {
                const int v = x86_d2i(__dadd_rn(__dadd_rn((double)(ox0 + o - RX + xq), mvx), 0.5));
                fx[o] = (v < 0 || v > W - 1) ? -1 : v;
                if (ox0 + o < SX) allok &= fx[o] >= 0;
            }
#pragma unroll
            for (int oy = 0; oy < 12; ++oy) {
                const int v = x86_d2i(__dadd_rn(__dadd_rn((double)(oy - ry + yq), mvy), 0.5));
                fy[oy] = (oy >= Sy || v < 0 || v > H - 1) ? -1 : v * W;           // premultiplied: a gather's word index is one add
                if (oy < Sy) allok &= fy[oy] >= 0;
            }
        }
        const bool fast = __all_sync(0xffffffffu, allok);          // every thread of the block gets here: no divergence above
        if (inimg) {
            if (fast) {
#pragma unroll
                for (int o = 0; o < OXG; ++o) {
                    if (ox0 + o < SX) {
#pragma unroll
                        for (int j = 0; j < 3; ++j) {
                            uint32_t w = 0;
#pragma unroll
                            for (int b = 0; b < 4; ++b) {
                                const int oy = 4 * j + b;
                                if (oy < Sy) w += (uint32_t)__popc(w1 ^ __ldg(c2 + (uint32_t)(fy[oy] + fx[o]))) << (8 * b);
                            }
                            rawb[((ox0 + o) * 3 + j) * PCS_COLS + tid] = w;
                        }
                    }
                }
            } else {
#pragma unroll
                for (int o = 0; o < OXG; ++o) {
                    if (ox0 + o < SX) {
#pragma unroll
                        for (int j = 0; j < 3; ++j) {
                            uint32_t w = 0;
#pragma unroll
                            for (int b = 0; b < 4; ++b) {
                                const int oy = 4 * j + b;
                                const bool ok = fx[o] >= 0 && fy[oy] >= 0;
                                const uint32_t h = ok ? (uint32_t)__popc(w1 ^ __ldg(c2 + (uint32_t)(max(fy[oy], 0) + max(fx[o], 0)))) : 5u;
                                w += h << (8 * b);
                            }
                            rawb[((ox0 + o) * 3 + j) * PCS_COLS + tid] = w;
                        }
                    }
                }
            }
        } else {
            // a window pixel outside the image contributes the constant (:405-408)
#pragma unroll
            for (int o = 0; o < OXG; ++o)
                if (ox0 + o < SX)
#pragma unroll
                    for (int j = 0; j < 3; ++j) rawb[((ox0 + o) * 3 + j) * PCS_COLS + tid] = 0x05050505u;
        }
        __syncthreads();
        // ---- horizontal T-sum (bytes), vertical running sum (u16x2), ring of the last T rows ------------------------------------
        if (interior) {
#pragma unroll
            for (int o = 0; o < OXG; ++o) {
                if (ox0 + o < SX) {
#pragma unroll
                    for (int j = 0; j < 3; ++j) {
                        const int w = (ox0 + o) * 3 + j;
                        uint32_t hs = 0;
#pragma unroll
                        for (int ax = -AGG; ax <= AGG; ++ax) hs += rawb[w * PCS_COLS + tid + ax];
                        uint32_t* rp = ring + (slot * NW + w) * PCS_COLS + tid;
                        const uint32_t old = *rp;
                        *rp = hs;
                        vs[o][2 * j] += __byte_perm(hs, 0, 0x4140) - __byte_perm(old, 0, 0x4140);
                        vs[o][2 * j + 1] += __byte_perm(hs, 0, 0x4342) - __byte_perm(old, 0, 0x4342);
                    }
                }
            }
            const int yp = yq - AGG;
            if (yp >= y0) {
                uint8_t* op = Cb + (((size_t)yp * SX) * W + xq) * 16;
#pragma unroll
                for (int o = 0; o < OXG; ++o) {
                    if (ox0 + o < SX) {
                        // (u8)(1.0*s/wp + 0.5) by one fp16 fma per two labels: the u16 sum s < 1024 IS the fp16 subnormal s * 2^-24,
                        // and fp16(2^(24-k) / wp) * (s * 2^-24) + 2^(10-k) rounds to pattern (B | q), q = round(s / wp) (k = 4, B =
                        // 0x5400 for wp = 25; k = 5, B = 0x5000 for wp = 9; checked for every s <= 1023 — the derivation is in
                        // cost_epi.cu)
                        uint32_t nq[6];
#pragma unroll
                        for (int i = 0; i < 6; ++i)
                            asm("fma.rn.f16x2 %0, %1, %2, %3;" : "=r"(nq[i]) : "r"(vs[o][i]), "r"(WPX == 25 ? 0x791F791Fu : 0x7B1C7B1Cu), "r"(WPX == 25 ? 0x54005400u : 0x50005000u));
                        uint4 f;
                        f.x = __byte_perm(nq[0], 0xFFFFFFFFu, 0x2044) | padw[0];
                        f.y = __byte_perm(nq[1], nq[2], 0x6420) | padw[1];
                        f.z = __byte_perm(nq[3], nq[4], 0x6420) | padw[2];
                        f.w = __byte_perm(nq[5], 0xFFFFFFFFu, 0x4440) | padw[3];
                        *reinterpret_cast<uint4*>(op + (size_t)(ox0 + o) * W * 16) = f;
                    }
                }
            }
        }
        slot = slot + 1 == T ? 0 : slot + 1;
        __syncthreads();
    }
}

// pixels whose aggregation window does not carry one prior (bit patterns compared): appended to the pair's list.  The list is
// filled from both ends: pixels for which every tap of pyd_cost_px_kernel is certainly inside both images (window and the
// displaced window with a margin for the rounding of the prior) from the front, count[2 * pair]; the others from the back,
// count[2 * pair + 1] — so that the 32 lanes of a warp of the list kernel take its check-free path together, which 32 pixels
// appended in arrival order almost never do (one lane near the image border is enough).
__global__ void pyd_uniform_kernel(const double* __restrict__ preMv, int mvW, int mvH, int W, int H, int agg, int rx, int ry,
                                   uint32_t* __restrict__ list, uint32_t* __restrict__ count)
{
    const size_t N = (size_t)W * H;
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int pair = blockIdx.y;
    bool flag = false, inner = false;
    if (i < N) {
        const int y = (int)(i / W), x = (int)(i - (size_t)y * W);
        const long long* mx = reinterpret_cast<const long long*>(preMv) + (size_t)pair * 2 * mvW * mvH;
        const long long* my = mx + (size_t)mvW * mvH;
        const long long cx = mx[(size_t)mvW * y + x], cy = my[(size_t)mvW * y + x];
        for (int ay = -agg; ay <= agg; ++ay) {
            const int yy = y + ay;
            if (yy < 0 || yy >= H) continue;
            for (int ax = -agg; ax <= agg; ++ax) {
                const int xx = x + ax;
                if (xx < 0 || xx >= W) continue;
                flag |= mx[(size_t)mvW * yy + xx] != cx || my[(size_t)mvW * yy + xx] != cy;
            }
        }
        // conservative: the sample coordinates are (int)(s + mv + 0.5), s within rx + agg (ry + agg) of the pixel
        const double fx = __longlong_as_double(cx), fy = __longlong_as_double(cy);
        const double mxr = (double)(rx + agg + 2), myr = (double)(ry + agg + 2);
        inner = x >= agg && x + agg < W && y >= agg && y + agg < H &&
                (double)x + fx - mxr >= 0.0 && (double)x + fx + mxr <= (double)(W - 1) &&
                (double)y + fy - myr >= 0.0 && (double)y + fy + myr <= (double)(H - 1);     // false for NaN
    }
    // one append per block and class: the block's flagged pixels (128 consecutive pixels of a row) stay together in the list, so a
    // warp of the list kernel gathers from one or two row segments instead of wherever concurrently running warps appended
    __shared__ uint32_t wc[2][4], wb[2];
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    const unsigned mi = __ballot_sync(0xffffffffu, flag && inner), mo = __ballot_sync(0xffffffffu, flag && !inner);
    if (lane == 0) { wc[0][wib] = (uint32_t)__popc(mi); wc[1][wib] = (uint32_t)__popc(mo); }
    __syncthreads();
    if (threadIdx.x < 2) {
        const uint32_t t = wc[threadIdx.x][0] + wc[threadIdx.x][1] + wc[threadIdx.x][2] + wc[threadIdx.x][3];
        wb[threadIdx.x] = t ? atomicAdd(count + 2 * pair + threadIdx.x, t) : 0u;
    }
    __syncthreads();
    uint32_t bi = wb[0], bo = wb[1];
    for (int w = 0; w < wib; ++w) { bi += wc[0][w]; bo += wc[1][w]; }
    const unsigned below = (1u << lane) - 1u;
    if (flag && inner) list[(size_t)pair * N + bi + __popc(mi & below)] = (uint32_t)i;
    if (flag && !inner) list[(size_t)pair * N + (N - 1) - (bo + __popc(mo & below))] = (uint32_t)i;
}

static size_t pyd_cost_px_smem(int agg, int rx, int ry, int pitch)
{
    const int T = 2 * agg + 1, Sx = 2 * rx + 1, Sy = 2 * ry + 1, SX2 = Sx + T - 1, SY2 = Sy + T - 1;
    const size_t PT = pitch ? pitch : Sx * Sy;
    const size_t per_warp = (size_t)(SX2 + SY2) * 32 * 4 + (((size_t)32 * PT + 15) & ~(size_t)15);
    return per_warp * PYC_WARPS;
}

int launch_pyd_cost(fsgm_ctx* c, int n, const uint32_t* cen1, const uint32_t* cen2, int W, int H,
                    const double* preMv, int mvW, int mvH, int agg, int rx, int ry, uint8_t* C, int pitch, int soa,
                    const uint32_t* list, const uint32_t* list_count)
{
    StageScope ss(c, ST_PYD_COST);
    if (2 * (rx + agg) + 1 > PYD_MAXS + 32 || 2 * (ry + agg) + 1 > PYD_MAXS + 32)
        return fail(c, FSGM_ERR_DOMAIN, "search + aggregation window too large");
    const size_t N = (size_t)W * H;
    // lane = pixel kernel for the aggregation windows in use (5x5, 3x3) when its per-warp staging fits; else one warp per pixel
    const size_t smem = (agg == 1 || agg == 2) ? pyd_cost_px_smem(agg, rx, ry, pitch) : 0;
    if (pitch && !(smem && smem <= 100 * 1024)) return fail(c, FSGM_ERR_DOMAIN, "padded cost volume needs the lane = pixel cost kernel");
    if (smem && smem <= 100 * 1024) {
        const int jobs = ((W + 31) / 32) * H + (list ? 1 : 0);
        dim3 grid((unsigned)((jobs + PYC_WARPS - 1) / PYC_WARPS), n);
        if (agg == 2) {
            FSGM_CUDA(c, cudaFuncSetAttribute(pyd_cost_px_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            pyd_cost_px_kernel<2><<<grid, PYC_WARPS * 32, smem, c->stream>>>(cen1, cen2, W, H, preMv, mvW, mvH, rx, ry, pitch, soa, C, list, list_count);
        } else {
            FSGM_CUDA(c, cudaFuncSetAttribute(pyd_cost_px_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            pyd_cost_px_kernel<1><<<grid, PYC_WARPS * 32, smem, c->stream>>>(cen1, cen2, W, H, preMv, mvW, mvH, rx, ry, pitch, soa, C, list, list_count);
        }
        FSGM_LAUNCHED(c);
        return FSGM_OK;
    }
    dim3 grid((unsigned)((N + 7) / 8), n);
    pyd_cost_kernel<<<grid, 256, 0, c->stream>>>(cen1, cen2, W, H, preMv, mvW, mvH, agg, rx, ry, C);
    FSGM_LAUNCHED(c);
    return FSGM_OK;
}


template <int AGG, int SX>
static int pcs_launch(fsgm_ctx* c, int n, const uint32_t* cen1, const uint32_t* cen2, int W, int H, const double* preMv, int mvW, int mvH,
                      int ry, uint8_t* C)
{
    constexpr int T = 2 * AGG + 1;
    constexpr int G = SX >= 9 ? 3 : SX >= 5 ? 2 : 1;
    const size_t smem = (size_t)(T + 1) * 3 * SX * PCS_COLS * 4;
    auto kern = pyd_cost_sep_kernel<AGG, SX, G>;
    FSGM_CUDA(c, cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const int TW = PCS_COLS - 2 * AGG;
    dim3 grid((W + TW - 1) / TW, (H + PCS_ROWS - 1) / PCS_ROWS, n);
    kern<<<grid, PCS_COLS * G, smem, c->stream>>>(cen1, cen2, W, H, preMv, mvW, mvH, ry, C);
    FSGM_LAUNCHED(c);
    return FSGM_OK;
}

// cost volume in the [y][label column][x][16-byte frame] layout: separable box filter everywhere + exact recomputation of the
// pixels whose aggregation window sees more than one prior.  list: n*N u32, count: 2*n u32 (scratch).
int launch_pyd_cost_sep(fsgm_ctx* c, int n, const uint32_t* cen1, const uint32_t* cen2, int W, int H,
                        const double* preMv, int mvW, int mvH, int agg, int rx, int ry, uint8_t* C, uint32_t* list, uint32_t* count)
{
    const size_t N = (size_t)W * H;
    {
        StageScope ss(c, ST_PYD_COST);
        FSGM_CUDA(c, cudaMemsetAsync(count, 0, (size_t)n * 8, c->stream));
        pyd_uniform_kernel<<<dim3((unsigned)((N + 127) / 128), n), 128, 0, c->stream>>>(preMv, mvW, mvH, W, H, agg, rx, ry, list, count);
        FSGM_LAUNCHED(c);
        const int Sx = 2 * rx + 1;
#define PCS_GO(A, S) FSGM_TRY((pcs_launch<A, S>(c, n, cen1, cen2, W, H, preMv, mvW, mvH, ry, C)))
        if (agg == 2) {
            switch (Sx) { case 1: PCS_GO(2, 1); break; case 3: PCS_GO(2, 3); break; case 5: PCS_GO(2, 5); break;
                          case 7: PCS_GO(2, 7); break; case 9: PCS_GO(2, 9); break; case 11: PCS_GO(2, 11); break;
                          default: return fail(c, FSGM_ERR_DOMAIN, "separable pyd cost: window width must be odd and <= 11"); }
        } else if (agg == 1) {
            switch (Sx) { case 1: PCS_GO(1, 1); break; case 3: PCS_GO(1, 3); break; case 5: PCS_GO(1, 5); break;
                          case 7: PCS_GO(1, 7); break; case 9: PCS_GO(1, 9); break; case 11: PCS_GO(1, 11); break;
                          default: return fail(c, FSGM_ERR_DOMAIN, "separable pyd cost: window width must be odd and <= 11"); }
        } else return fail(c, FSGM_ERR_DOMAIN, "separable pyd cost: aggregation radius 1 or 2");
#undef PCS_GO
    }
    return launch_pyd_cost(c, n, cen1, cen2, W, H, preMv, mvW, mvH, agg, rx, ry, C, 16 * (2 * rx + 1), 1, list, count);
}

int launch_pyd_sweeps(fsgm_ctx* c, int n, const uint8_t* C, const uint8_t* I1, const double* preMv, int mvW, int mvH,
                      int W, int H, int Sx, int Sy, int P1, int P2, int adaptive, const int* dirs, int n_dirs, uint8_t* const* Lvols,
                      int pitch)
{
    StageScope ss(c, ST_PYD_SWEEP);
    const int D = Sx * Sy;
    if (D > PYD_MAXD || Sx > PYD_MAXS || Sy > PYD_MAXS) return fail(c, FSGM_ERR_DOMAIN, "search window too large (<= 1024 labels)");
    PydSweepParams p{};
    p.C = C; p.I1 = I1; p.preMv = preMv; p.n_dirs = n_dirs; p.W = W; p.H = H; p.Sx = Sx; p.Sy = Sy; p.mvW = mvW; p.mvH = mvH;
    p.P1 = P1; p.P2 = P2; p.adaptive = adaptive; p.pitch = pitch;
    p.line_start[0] = 0;
    for (int k = 0; k < n_dirs; ++k) {
        p.dir[k] = dirs[k]; p.L[k] = Lvols[k];
        p.line_start[k + 1] = p.line_start[k] + (dir_dy(dirs[k]) == 0 ? H : W);
    }
    dim3 grid((p.line_start[n_dirs] + PYD_WARPS - 1) / PYD_WARPS, n);
    // cost values of this variant are <= 24 (mean of 5x5 Hamming distances <= 23 and the constant 5)
    const bool fast = P1 >= 0 && P2 >= 0 && 25 + P1 + P2 <= 255 && 50 + P2 <= 255;
    if (pitch && !(fast && Sy <= 12 && Sx <= 16 && pitch == 16 * Sx)) return fail(c, FSGM_ERR_DOMAIN, "padded volumes need the vector mode of the pyd sweep");
#define PYD_GO(NJV)                                                                                   \
    do { if (fast) pyd_sweep_kernel<NJV, true><<<grid, PYD_WARPS * 32, 0, c->stream>>>(p);             \
         else pyd_sweep_kernel<NJV, false><<<grid, PYD_WARPS * 32, 0, c->stream>>>(p); } while (0)
    if (D <= 128) PYD_GO(4); else if (D <= 256) PYD_GO(8); else if (D <= 512) PYD_GO(16); else PYD_GO(32);
#undef PYD_GO
    FSGM_LAUNCHED(c);
    return FSGM_OK;
}

// ------------------------------------------------------------------------------------------------
// WTA + per-axis subpixel: one warp per pixel group, weighted sum of the direction volumes
// ------------------------------------------------------------------------------------------------
struct PydWtaParams {
    const uint8_t* L[8]; int weight[8]; int n_dirs;
    int W, H, Sx, Sy, subpixel;
    uint16_t* Sp16; uint32_t* bestD; uint32_t* minC; double* mvSub;
};

__global__ void __launch_bounds__(PYD_WARPS * 32)
pyd_wta_kernel(const PydWtaParams prm)
{
    __shared__ uint16_t sums[PYD_WARPS][PYD_MAXD];
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    const int Sy = prm.Sy, Sx = prm.Sx, D = Sx * Sy, R = prm.n_dirs;
    const size_t N = (size_t)prm.W * prm.H, vol = blockIdx.y * N * D;
    const size_t p = (size_t)blockIdx.x * PYD_WARPS + wib;
    if (p >= N) return;
    uint16_t* s = sums[wib];
    uint32_t key = 0xFFFFFFFFu;
    for (int d = lane; d < D; d += 32) {
        // all (up to 8) direction volumes are requested before the first is used: the kernel is bound by load latency
        uint32_t v[8];
#pragma unroll
        for (int r = 0; r < 8; ++r) v[r] = r < R ? (uint32_t)__ldg(prm.L[r] + vol + p * D + d) : 0u;
        uint32_t a = 0;
#pragma unroll
        for (int r = 0; r < 8; ++r) a += r < R ? prm.weight[r] * v[r] : 0u;
        s[d] = (uint16_t)a;
        if (prm.Sp16) prm.Sp16[vol + p * D + d] = (uint16_t)a;
        key = min(key, (a << 16) | (uint32_t)d);
    }
    if (R == 0) key = 0;                               // totalPass == 0: Sp stays zero, argmin = 0
    key = __reduce_min_sync(0xffffffffu, key);
    __syncwarp();
    if (lane != 0) return;
    const uint32_t idx = key & 0xFFFFu, best = key >> 16;
    const size_t gp = blockIdx.y * N + p;
    prm.bestD[gp] = idx;
    prm.minC[gp] = best;
    double sx = 0.0, sy = 0.0;
    if (prm.subpixel && R > 0) {
        const int lx = idx / Sy, ly = idx - lx * Sy;
        const double c0 = (double)best;
        if (ly > 0 && ly < Sy - 1) {
            const double a = s[idx - 1], b = s[idx + 1];
            sy = (b < a) ? __ddiv_rn(__ddiv_rn(__dsub_rn(b, a), __dsub_rn(c0, a)), 2.0)
                         : __ddiv_rn(__ddiv_rn(__dsub_rn(b, a), __dsub_rn(c0, b)), 2.0);
        }
        if (lx > 0 && lx < Sx - 1) {
            const double a = s[idx - Sy], b = s[idx + Sy];
            sx = (b < a) ? __ddiv_rn(__ddiv_rn(__dsub_rn(b, a), __dsub_rn(c0, a)), 2.0)
                         : __ddiv_rn(__ddiv_rn(__dsub_rn(b, a), __dsub_rn(c0, b)), 2.0);
        }
    }
    prm.mvSub[blockIdx.y * 2 * N + p] = sx;
    prm.mvSub[blockIdx.y * 2 * N + N + p] = sy;
}

int launch_pyd_wta(fsgm_ctx* c, int n, uint8_t* const* Lvols, const int* weights, int n_dirs, int W, int H, int Sx, int Sy,
                   int subpixel, uint16_t* Sp16, uint32_t* bestD, uint32_t* minC, double* mvSub)
{
    StageScope ss(c, ST_PYD_WTA);
    PydWtaParams p{};
    for (int k = 0; k < n_dirs; ++k) { p.L[k] = Lvols[k]; p.weight[k] = weights[k]; }
    p.n_dirs = n_dirs; p.W = W; p.H = H; p.Sx = Sx; p.Sy = Sy; p.subpixel = subpixel;
    p.Sp16 = Sp16; p.bestD = bestD; p.minC = minC; p.mvSub = mvSub;
    const size_t N = (size_t)W * H;
    dim3 grid((unsigned)((N + PYD_WARPS - 1) / PYD_WARPS), n);
    pyd_wta_kernel<<<grid, PYD_WARPS * 32, 0, c->stream>>>(p);
    FSGM_LAUNCHED(c);
    return FSGM_OK;
}

}  // namespace fsgm
