// Pyramidal neighbour-guided variant — replaces calc_cost(), sgm_step(), sgm2d() and subpixel_refine() of the
// reference's calc_pyd_cost_sgm_ng.cpp (:370-446, :39-78, :101-306, :308-368).
//
// Labels are explicit candidates: 9 prior-flow hints (3x3 grid, stride 8, dy outer, clamped to the prior map,
// :390-396) x (2r+1)^2 integer offsets (offx outer, :399-400).  Entry d = h*S*S + ox*S + oy carries
// mv = ((int)(mvx_h + offx), (int)(mvy_h + offy)) (:433-434) — stored here factored as X[h][ox], Y[h][oy]
// (2*9*S ints per pixel instead of 2*D) — and cost = (int)(1.0*sum/winPixels + 0.5) with the sample at
// ((int)((offx+x1)+mvx), (int)((offy+y1)+mvy)): NO +0.5 (:417-418), constant 5 outside the image.
//
// Aggregation: 2 passes x {horizontal, vertical} (:120-122), each step an O(D^2) compatibility search against
// the previous pixel's candidates: identical mv -> same-label term (the LAST match wins, :62-63), |dmv| <= 2 in
// both axes -> P1 term (:64-65), else P2.  Costs are int in the reference; the running minimum and the three
// terms are unsigned char (:45-56, :74).  This stage is integer-issue-bound, not HBM-bound.
#include "fsgm_internal.h"

namespace fsgm {

constexpr int NG_WARPS = 4;
constexpr int PYDNG_MAXR = 3;                       // (2r+1)^2*9 <= 441 candidates
constexpr int PYDNG_MAXS = 2 * PYDNG_MAXR + 1;
constexpr int PYDNG_MAXD = 9 * PYDNG_MAXS * PYDNG_MAXS;
constexpr int PYDNG_MAXAGG = 4;
constexpr int PYDNG_TAB = 2 * (PYDNG_MAXR + PYDNG_MAXAGG) + 1;

__device__ __forceinline__ int ng_d2i(double v)
{
    if (!(v > -2147483649.0 && v < 2147483648.0)) return INT_MIN;
    return __double2int_rz(v);
}

// ------------------------------------------------------------------------------------------------
// candidate volume: one warp per pixel
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(NG_WARPS * 32)
pydng_cost_kernel(const uint32_t* __restrict__ cen1, const uint32_t* __restrict__ cen2, int W, int H,
                  const double* __restrict__ preMv, int mvW, int mvH, int agg, int r,
                  uint8_t* __restrict__ cost, int* __restrict__ XY)
{
    __shared__ int fx[NG_WARPS][9][PYDNG_TAB], fy[NG_WARPS][9][PYDNG_TAB];
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    const size_t N = (size_t)W * H;
    const size_t p = (size_t)blockIdx.x * NG_WARPS + wib;
    if (p >= N) return;
    const int pair = blockIdx.y;
    const int x = (int)(p % W), y = (int)(p / W);
    const int S = 2 * r + 1, SS = S * S, D = 9 * SS, T = 2 * (r + agg) + 1;
    const double* mvxP = preMv + (size_t)pair * 2 * mvW * mvH;
    const double* mvyP = mvxP + (size_t)mvW * mvH;
    int* xy = XY + (pair * N + p) * (size_t)(18 * S);
    // per hint: sample-coordinate tables over s = off + a, and the entry mv tables X[h][ox], Y[h][oy]
    for (int i = lane; i < 9 * T; i += 32) {
        const int h = i / T, s = i - h * T;
        const int yn = min(max(y + (h / 3 - 1) * 8, 0), mvH - 1), xn = min(max(x + (h % 3 - 1) * 8, 0), mvW - 1);
        const double mvx = mvxP[(size_t)mvW * yn + xn], mvy = mvyP[(size_t)mvW * yn + xn];
        int vx = ng_d2i(__dadd_rn((double)(s - r - agg + x), mvx));
        int vy = ng_d2i(__dadd_rn((double)(s - r - agg + y), mvy));
        fx[wib][h][s] = (vx < 0 || vx > W - 1) ? -1 : vx;
        fy[wib][h][s] = (vy < 0 || vy > H - 1) ? -1 : vy;
    }
    for (int i = lane; i < 9 * S; i += 32) {
        const int h = i / S, o = i - h * S - r;
        const int yn = min(max(y + (h / 3 - 1) * 8, 0), mvH - 1), xn = min(max(x + (h % 3 - 1) * 8, 0), mvW - 1);
        xy[i] = ng_d2i(__dadd_rn(mvxP[(size_t)mvW * yn + xn], (double)o));
        xy[9 * S + i] = ng_d2i(__dadd_rn(mvyP[(size_t)mvW * yn + xn], (double)o));
    }
    __syncwarp();
    const uint32_t* c1 = cen1 + pair * N;
    const uint32_t* c2 = cen2 + pair * N;
    const int wp = (2 * agg + 1) * (2 * agg + 1);
    uint8_t* out = cost + (pair * N + p) * D;
    for (int d = lane; d < D; d += 32) {
        const int h = d / SS, rem = d - h * SS, ox = rem / S, oy = rem - ox * S;
        uint32_t s = 0;
        for (int ay = -agg; ay <= agg; ++ay) {
            const int y1 = y + ay;
            for (int ax = -agg; ax <= agg; ++ax) {
                const int x1 = x + ax;
                int hh = 5;
                if (y1 >= 0 && y1 < H && x1 >= 0 && x1 < W) {
                    const int x2 = fx[wib][h][ox + ax + agg], y2 = fy[wib][h][oy + ay + agg];
                    if (x2 >= 0 && y2 >= 0) hh = __popc(__ldg(c1 + (size_t)W * y1 + x1) ^ __ldg(c2 + (size_t)W * y2 + x2));
                }
                s += hh;
            }
        }
        out[d] = (uint8_t)((2 * s + wp) / (2 * wp));        // == (int)(1.0*s/wp + 0.5), see pyd.cu
    }
}

// ------------------------------------------------------------------------------------------------
// sweep: one warp per scanline; the previous pixel's expanded candidates and path costs live in smem
// ------------------------------------------------------------------------------------------------
struct NgSweepParams {
    const uint8_t* cost; const int* XY;
    int16_t* L[4]; int dir[4]; int line_start[5]; int n_dirs;
    int W, H, S, P1, P2;
};

template <int NJ>
__global__ void __launch_bounds__(NG_WARPS * 32)
pydng_sweep_kernel(const NgSweepParams prm)
{
    __shared__ int ex[NG_WARPS][2][NJ * 32], ey[NG_WARPS][2][NJ * 32], lc[NG_WARPS][2][NJ * 32];
    __shared__ uchar4 gr4[NG_WARPS][9 * 6];          // S == 3: row-interval minima (rows 0, 1, 2) of the previous pixel's nine candidate grids
    __shared__ int4 gm[NG_WARPS][9];                 //         smallest P1 term, its cell, second smallest
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    const int gw = blockIdx.x * NG_WARPS + wib;
    if (gw >= prm.line_start[prm.n_dirs]) return;
    int k = 0;
    while (gw >= prm.line_start[k + 1]) ++k;
    const int line = gw - prm.line_start[k], r = prm.dir[k];
    const int dx = dir_dx(r), dy = dir_dy(r);
    const int W = prm.W, H = prm.H, S = prm.S, SS = S * S, D = 9 * SS;
    const size_t N = (size_t)W * H;
    const uint8_t* __restrict__ Cb = prm.cost + blockIdx.y * N * D;
    const int* __restrict__ XYb = prm.XY + blockIdx.y * N * (size_t)(18 * S);
    int16_t* __restrict__ Lb = prm.L[k] + blockIdx.y * N * D;

    int x, y, len;
    if (dy == 0) { y = line; x = dx > 0 ? 0 : W - 1; len = W; }
    else         { x = line; y = dy > 0 ? 0 : H - 1; len = H; }

    int lh[NJ], lox[NJ], loy[NJ];
#pragma unroll
    for (int j = 0; j < NJ; ++j) {
        int d = lane + 32 * j; lh[j] = d / SS; int rem = d - lh[j] * SS; lox[j] = rem / S; loy[j] = rem - lox[j] * S;
    }
    uint32_t M = 0;
    int cur = 0;
    for (int t = 0; t < len; ++t) {
        const size_t pix = (size_t)y * W + x;
        const int* xy = XYb + pix * (size_t)(18 * S);
        int mx[NJ], my[NJ], cc[NJ];
#pragma unroll
        for (int j = 0; j < NJ; ++j) {
            const int d = lane + 32 * j;
            if (d < D) {
                mx[j] = __ldg(xy + lh[j] * S + lox[j]);
                my[j] = __ldg(xy + 9 * S + lh[j] * S + loy[j]);
                cc[j] = __ldg(Cb + pix * D + d);
            } else { mx[j] = my[j] = cc[j] = 0; }
        }
        int* exn = ex[wib][cur]; int* eyn = ey[wib][cur]; int* lcn = lc[wib][cur];
        const int* exp_ = ex[wib][cur ^ 1]; const int* eyp = ey[wib][cur ^ 1]; const int* lcp = lc[wib][cur ^ 1];
        int newL[NJ];
        if (t == 0) {
#pragma unroll
            for (int j = 0; j < NJ; ++j) newL[j] = cc[j];
            M = 0;
        } else {
            const uint32_t far_ = (M + (uint32_t)prm.P2) & 0xFFu;
            uint32_t same[NJ], near_[NJ];
#pragma unroll
            for (int j = 0; j < NJ; ++j) { same[j] = far_; near_[j] = far_; }
            // S == 3: the previous pixel's candidates are nine 3 x 3 grids of consecutive flow vectors (when (int)(mv + off) has
            // no duplicate at zero: checked).  For a candidate (mx, my) and a grid with corner (X0, Y0) the cell of EQUAL flow is
            // (ex, ey) = (mx - X0, my - Y0) and the cells within +-2 are the x-interval [ex-2, ex+2] x the y-interval, clipped to
            // the grid — the whole grid whenever the equal cell lies inside it.  So per grid: the smallest and second-smallest
            // P1 term with the position of the smallest (equal cell inside: the minimum over the grid WITHOUT that cell), and
            // the minima over the six x-intervals per grid row (equal cell outside: at most three of them).  9 grid queries per
            // candidate instead of 81 cell tests.
            bool grid_ok = false;
            if (S == 3) {
                bool reg = true;
                if (lane < 9) {
                    const int b = lane * 9;
                    reg = exp_[b + 3] == exp_[b] + 1 && exp_[b + 6] == exp_[b] + 2 && eyp[b + 1] == eyp[b] + 1 && eyp[b + 2] == eyp[b] + 2;
                }
                grid_ok = __all_sync(0xffffffffu, reg);
            }
            if (grid_ok) {
                // row-interval minima: entry (h, xi, oy), x-intervals [0,0] [0,1] [0,2] [1,1] [1,2] [2,2]
                for (int e = lane; e < 9 * 6; e += 32) {
                    const int h = e / 6, xi = e - h * 6;
                    const int a = xi < 3 ? 0 : xi < 5 ? 1 : 2, bnd = xi == 0 ? 0 : (xi == 1 || xi == 3) ? 1 : 2;
                    uint32_t v[3] = {255, 255, 255};
                    for (int ox = a; ox <= bnd; ++ox)
#pragma unroll
                        for (int oy = 0; oy < 3; ++oy) v[oy] = min(v[oy], ((uint32_t)lcp[h * 9 + ox * 3 + oy] + (uint32_t)prm.P1) & 0xFFu);
                    gr4[wib][e] = make_uchar4((unsigned char)v[0], (unsigned char)v[1], (unsigned char)v[2], 255);
                }
                if (lane < 9) {
                    uint32_t k1 = 0xFFFFFFFFu, k2 = 0xFFFFFFFFu;
                    for (int cidx = 0; cidx < 9; ++cidx) {
                        const uint32_t key = ((((uint32_t)lcp[lane * 9 + cidx] + (uint32_t)prm.P1) & 0xFFu) << 8) | (uint32_t)cidx;
                        if (key < k1) { k2 = k1; k1 = key; } else if (key < k2) k2 = key;
                    }
                    gm[wib][lane] = make_int4((int)(k1 >> 8), (int)(k1 & 0xFFu), (int)(k2 >> 8), 0);
                }
                __syncwarp();
#pragma unroll
                for (int j = 0; j < NJ; ++j) {
                    for (int h = 0; h < 9; ++h) {
                        const int ex = mx[j] - exp_[h * 9], ey = my[j] - eyp[h * 9];
                        const uint32_t ux = (uint32_t)(ex + 2), uy = (uint32_t)(ey + 2);
                        // warp-uniform branches only: a form is evaluated when some lane needs it, and selected per lane
                        const bool valid = ux <= 6u && uy <= 6u;             // something of this grid within +-2
                        const bool inside = valid && ux - 2u <= 2u && uy - 2u <= 2u;   // the cell of equal flow lies in the grid
                        if (!__any_sync(0xffffffffu, valid)) continue;
                        if (__any_sync(0xffffffffu, inside)) {
                            const int cell = inside ? ex * 3 + ey : 0;
                            const int4 g = gm[wib][h];
                            const uint32_t cs = (uint32_t)lcp[h * 9 + cell] & 0xFFu;
                            const uint32_t nin = (uint32_t)(cell == g.y ? g.z : g.x);
                            same[j] = inside ? cs : same[j];                 // later grids overwrite: the last match wins (:62-63)
                            near_[j] = min(near_[j], inside ? nin : 255u);
                        }
                        if (__any_sync(0xffffffffu, valid && !inside)) {
                            const int xi = ux == 0 ? 0 : ux == 1 ? 1 : ux <= 4 ? 2 : ux == 5 ? 4 : 5;
                            const uint32_t rw = (valid && !inside) ? *reinterpret_cast<const uint32_t*>(&gr4[wib][h * 6 + xi]) : 0xFFFFFFFFu;
                            // rows [uy-4, uy] of the three: drop row 0 when uy > 4, row 1 when uy > 5 or uy < 1, row 2 when uy < 2
                            const uint32_t r0 = uy <= 4u ? (rw & 0xFFu) : 255u;
                            const uint32_t r1 = (uy >= 1u && uy <= 5u) ? ((rw >> 8) & 0xFFu) : 255u;
                            const uint32_t r2 = uy >= 2u ? ((rw >> 16) & 0xFFu) : 255u;
                            near_[j] = min(near_[j], min(min(r0, r1), r2));
                        }
                    }
                }
            } else
            for (int d2 = 0; d2 < D; ++d2) {
                const int ax = exp_[d2], ay = eyp[d2];
                const uint32_t c2 = (uint32_t)lcp[d2];
                const uint32_t cs = c2 & 0xFFu, cn = (c2 + (uint32_t)prm.P1) & 0xFFu;
#pragma unroll
                for (int j = 0; j < NJ; ++j) {
                    // branch-free (data-dependent branches cost more than the work they skip)
                    const bool eq = (ax == mx[j]) & (ay == my[j]);
                    // |a-b| <= 2 in wrap-around int arithmetic, like abs(int - int) in the reference
                    const bool nr = ((uint32_t)(ax - mx[j] + 2) <= 4u) & ((uint32_t)(ay - my[j] + 2) <= 4u);
                    same[j] = eq ? cs : same[j];
                    near_[j] = min(near_[j], (nr & !eq) ? cn : 0xFFu);
                }
            }
            uint32_t m = 255;
#pragma unroll
            for (int j = 0; j < NJ; ++j) {
                const uint32_t best = min(min(far_, same[j]), near_[j]);
                newL[j] = cc[j] + (int)best - (int)M;
                if (lane + 32 * j < D) m = min(m, (uint32_t)newL[j] & 0xFFu);
            }
            M = __reduce_min_sync(0xffffffffu, m);
        }
#pragma unroll
        for (int j = 0; j < NJ; ++j) {
            const int d = lane + 32 * j;
            if (d < D) {
                exn[d] = mx[j]; eyn[d] = my[j]; lcn[d] = newL[j];
                Lb[pix * D + d] = (int16_t)newL[j];
            }
        }
        __syncwarp();
        cur ^= 1;
        if (dy == 0) x += dx; else y += dy;
    }
}

// ------------------------------------------------------------------------------------------------
// WTA over the summed path costs (unsigned compare, first minimum, :281-299) -> flow = winning entry's mv
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(NG_WARPS * 32)
pydng_wta_kernel(const int16_t* L0, const int16_t* L1, const int16_t* L2, const int16_t* L3, const int* __restrict__ XY,
                 int W, int H, int S, uint32_t* __restrict__ Sp32, uint32_t* __restrict__ minC, double* __restrict__ flow)
{
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    const int SS = S * S, D = 9 * SS;
    const size_t N = (size_t)W * H, vol = blockIdx.y * N * D;
    const size_t p = (size_t)blockIdx.x * NG_WARPS + wib;
    if (p >= N) return;
    unsigned long long key = ~0ull;
    for (int d = lane; d < D; d += 32) {
        const size_t i = vol + p * D + d;
        const uint32_t a = (uint32_t)((int)L0[i] + (int)L1[i] + (int)L2[i] + (int)L3[i]);
        if (Sp32) Sp32[i] = a;
        const unsigned long long kk = ((unsigned long long)a << 32) | (uint32_t)d;
        key = kk < key ? kk : key;
    }
    for (int o = 16; o; o >>= 1) {
        unsigned long long other = __shfl_xor_sync(0xffffffffu, key, o);
        key = other < key ? other : key;
    }
    if (lane != 0) return;
    const int d = (int)(key & 0xFFFFFFFFu);
    const int h = d / SS, rem = d - h * SS, ox = rem / S, oy = rem - ox * S;
    const int* xy = XY + (blockIdx.y * N + p) * (size_t)(18 * S);
    minC[blockIdx.y * N + p] = (uint32_t)(key >> 32);
    flow[blockIdx.y * 2 * N + p] = (double)xy[h * S + ox];
    flow[blockIdx.y * 2 * N + N + p] = (double)xy[9 * S + h * S + oy];
}

// census-based subpixel (:308-368); note the early exits: when the x refinement is skipped so is y (:337-338)
__global__ void pydng_subpixel_kernel(double* __restrict__ flow, const uint32_t* __restrict__ cen1,
                                      const uint32_t* __restrict__ cen2, int W, int H)
{
    const size_t N = (size_t)W * H;
    const size_t p = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= N) return;
    const int x = (int)(p % W), y = (int)(p / W);
    double* fxp = flow + blockIdx.y * 2 * N + p;
    double* fyp = fxp + N;
    const uint32_t* c2 = cen2 + blockIdx.y * N;
    const uint32_t c1 = cen1[blockIdx.y * N + p];
    const int tx = ng_d2i(__dadd_rn(*fxp, (double)x)), ty = ng_d2i(__dadd_rn(*fyp, (double)y));
    if (!(tx > 1 && tx < W - 1 && ty > 1 && ty < H - 1)) return;
    const double c0 = __popc(c1 ^ c2[(size_t)ty * W + tx]);
    double a = __popc(c1 ^ c2[(size_t)ty * W + tx - 1]), b = __popc(c1 ^ c2[(size_t)ty * W + tx + 1]);
    if (c0 >= a || c0 >= b) return;
    double s = (b < a) ? __ddiv_rn(__ddiv_rn(__dsub_rn(b, a), __dsub_rn(c0, a)), 2.0)
                       : __ddiv_rn(__ddiv_rn(__dsub_rn(b, a), __dsub_rn(c0, b)), 2.0);
    *fxp = __dadd_rn(*fxp, s);
    a = __popc(c1 ^ c2[(size_t)(ty - 1) * W + tx]); b = __popc(c1 ^ c2[(size_t)(ty + 1) * W + tx]);
    if (c0 >= a || c0 >= b) return;
    s = (b < a) ? __ddiv_rn(__ddiv_rn(__dsub_rn(b, a), __dsub_rn(c0, a)), 2.0)
                : __ddiv_rn(__ddiv_rn(__dsub_rn(b, a), __dsub_rn(c0, b)), 2.0);
    *fyp = __dadd_rn(*fyp, s);
}

int launch_pydng(fsgm_ctx* c, int n, const uint32_t* cen1, const uint32_t* cen2, int W, int H,
                 const double* preMv, int mvW, int mvH, int r, int agg, int subpixel, int P1, int P2,
                 uint8_t* cost, int* XY, int16_t* const* L, uint32_t* Sp32, uint32_t* minC, double* flow)
{
    if (r < 0 || r > PYDNG_MAXR) return fail(c, FSGM_ERR_DOMAIN, "halfSearchWinSize must be in 0..3");
    if (agg < 0 || agg > PYDNG_MAXAGG) return fail(c, FSGM_ERR_DOMAIN, "aggregation radius must be in 0..4");
    const size_t N = (size_t)W * H;
    const int S = 2 * r + 1, D = 9 * S * S;
    {
        StageScope ss(c, ST_PYDNG_COST);
        dim3 grid((unsigned)((N + NG_WARPS - 1) / NG_WARPS), n);
        pydng_cost_kernel<<<grid, NG_WARPS * 32, 0, c->stream>>>(cen1, cen2, W, H, preMv, mvW, mvH, agg, r, cost, XY);
        FSGM_LAUNCHED(c);
    }
    {
        StageScope ss(c, ST_PYDNG_SWEEP);
        NgSweepParams p{};
        p.cost = cost; p.XY = XY; p.n_dirs = 4; p.W = W; p.H = H; p.S = S; p.P1 = P1; p.P2 = P2;
        const int dirs[4] = {0, 1, 4, 5};                 // L1, L3 forward then reversed (:120-122)
        p.line_start[0] = 0;
        for (int k = 0; k < 4; ++k) {
            p.dir[k] = dirs[k]; p.L[k] = L[k];
            p.line_start[k + 1] = p.line_start[k] + (dir_dy(dirs[k]) == 0 ? H : W);
        }
        dim3 grid((p.line_start[4] + NG_WARPS - 1) / NG_WARPS, n);
        if (D <= 32) pydng_sweep_kernel<1><<<grid, NG_WARPS * 32, 0, c->stream>>>(p);
        else if (D <= 96) pydng_sweep_kernel<3><<<grid, NG_WARPS * 32, 0, c->stream>>>(p);
        else if (D <= 256) pydng_sweep_kernel<8><<<grid, NG_WARPS * 32, 0, c->stream>>>(p);
        else pydng_sweep_kernel<14><<<grid, NG_WARPS * 32, 0, c->stream>>>(p);
        FSGM_LAUNCHED(c);
    }
    {
        StageScope ss(c, ST_PYDNG_WTA);
        dim3 grid((unsigned)((N + NG_WARPS - 1) / NG_WARPS), n);
        pydng_wta_kernel<<<grid, NG_WARPS * 32, 0, c->stream>>>(L[0], L[1], L[2], L[3], XY, W, H, S, Sp32, minC, flow);
        FSGM_LAUNCHED(c);
        if (subpixel) {
            dim3 g2((unsigned)((N + 255) / 256), n);
            pydng_subpixel_kernel<<<g2, 256, 0, c->stream>>>(flow, cen1, cen2, W, H);
            FSGM_LAUNCHED(c);
        }
    }
    return FSGM_OK;
}

}  // namespace fsgm
