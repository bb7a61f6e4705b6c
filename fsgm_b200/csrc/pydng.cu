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
constexpr int PYDNG_MAXR = 5;                       // (2r+1)^2*9 <= 1089 candidates (the reference's own search windows are r = 1, 2)
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
// AGGT / RT: aggregation and search radius as compile-time constants (the reference's own settings), or -1: run-time values.
// Fast path (the pixel's window and all 9 T^2 displaced samples lie inside the image, T = 2 (r + agg) + 1 <= 9): a candidate's
// samples depend only on (hint, ox + ax, oy + ay), so the warp gathers the 9 T^2 census words once into shared memory
// (81 x 25 = 2025 loads -> 441 at r = 1) and every Hamming tap is LDS + XOR + POPC + ADD at a fixed offset.
constexpr int PYDNG_FAST_T = 9;
template <int AGGT, int RT>
__global__ void __launch_bounds__(NG_WARPS * 32)
pydng_cost_kernel(const uint32_t* __restrict__ cen1, const uint32_t* __restrict__ cen2, int W, int H,
                  const double* __restrict__ preMv, int mvW, int mvH, int agg_, int r_,
                  uint8_t* __restrict__ cost, int* __restrict__ XY)
{
    __shared__ int fx[NG_WARPS][9][PYDNG_TAB], fy[NG_WARPS][9][PYDNG_TAB];
    __shared__ uint32_t smp[NG_WARPS][9 * PYDNG_FAST_T * PYDNG_FAST_T];
    __shared__ uint32_t c1s[NG_WARPS][(2 * PYDNG_MAXAGG + 1) * (2 * PYDNG_MAXAGG + 1)];
    const int agg = AGGT >= 0 ? AGGT : agg_, r = RT >= 0 ? RT : r_;
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
    bool all_in = true;
    for (int i = lane; i < 9 * T; i += 32) {
        const int h = i / T, s = i - h * T;
        const int yn = min(max(y + (h / 3 - 1) * 8, 0), mvH - 1), xn = min(max(x + (h % 3 - 1) * 8, 0), mvW - 1);
        const double mvx = mvxP[(size_t)mvW * yn + xn], mvy = mvyP[(size_t)mvW * yn + xn];
        int vx = ng_d2i(__dadd_rn((double)(s - r - agg + x), mvx));
        int vy = ng_d2i(__dadd_rn((double)(s - r - agg + y), mvy));
        const bool bx = vx < 0 || vx > W - 1, by = vy < 0 || vy > H - 1;
        fx[wib][h][s] = bx ? -1 : vx;
        fy[wib][h][s] = by ? -1 : vy;
        all_in = all_in && !bx && !by;
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
    if (T <= PYDNG_FAST_T && x >= agg && x + agg < W && y >= agg && y + agg < H && __all_sync(0xffffffffu, all_in)) {
        const int K = 2 * agg + 1;
        for (int i = lane; i < 9 * T * T; i += 32) {
            const int h = i / (T * T), rem = i - h * T * T, sy = rem / T, sx = rem - sy * T;
            smp[wib][i] = __ldg(c2 + (size_t)W * fy[wib][h][sy] + fx[wib][h][sx]);
        }
        for (int i = lane; i < wp; i += 32) c1s[wib][i] = __ldg(c1 + (size_t)W * (y - agg + i / K) + (x - agg + i % K));
        __syncwarp();
        if (AGGT >= 0) {
            constexpr int KC = AGGT >= 0 ? 2 * AGGT + 1 : 1;
            uint32_t c1r[KC * KC];
#pragma unroll
            for (int i = 0; i < KC * KC; ++i) c1r[i] = c1s[wib][i];
            for (int d = lane; d < D; d += 32) {
                const int h = d / SS, rem = d - h * SS, ox = rem / S, oy = rem - ox * S;
                const uint32_t* sp = smp[wib] + h * T * T + oy * T + ox;
                uint32_t s = 0;
#pragma unroll
                for (int ay = 0; ay < KC; ++ay)
#pragma unroll
                    for (int ax = 0; ax < KC; ++ax) s += __popc(c1r[ay * KC + ax] ^ sp[ay * T + ax]);
                out[d] = (uint8_t)((2 * s + wp) / (2 * wp));
            }
        } else {
            for (int d = lane; d < D; d += 32) {
                const int h = d / SS, rem = d - h * SS, ox = rem / S, oy = rem - ox * S;
                const uint32_t* sp = smp[wib] + h * T * T + oy * T + ox;
                uint32_t s = 0;
                for (int ay = 0; ay < K; ++ay)
                    for (int ax = 0; ax < K; ++ax) s += __popc(c1s[wib][ay * K + ax] ^ sp[ay * T + ax]);
                out[d] = (uint8_t)((2 * s + wp) / (2 * wp));
            }
        }
        return;
    }
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
    int grid_tables;          // 0: every step by the cell-by-cell search (A/B and parity of the tabulated form)
};

// Shared memory of one warp (dynamic): the previous / current pixel's expanded candidates (x, y, path cost), and — when the
// previous pixel's nine candidate grids are REGULAR (consecutive flow vectors: (int)(mv + off) has no duplicate at zero) — the
// tabulated compatibility search.  For a candidate (mx, my) and a grid with corner (X0, Y0) the cell of EQUAL flow is
// (ex, ey) = (mx - X0, my - Y0), and the cells within +-2 (the P1 term, :64-65) are [ex-2, ex+2] x [ey-2, ey+2] clipped to the
// grid, minus that cell.  The answer depends only on (ux, uy) = (ex + 2, ey + 2) in 0..S+3 (anything else: no compatible cell),
// so it is tabulated per grid: T[h][ux * TW + uy], low half = smallest P1 term, bits 16-23 = cost of the equal cell, bit 24 =
// there is one; row / column S + 4 = 0xFF (written once).  The table is built separably (window minima along y with and
// without the centre, then along x), and the search is ONE look-up per candidate and grid: 9 instead of 9 S^2 entry tests.
template <int S> struct PydngSmem {
    static constexpr int SS = S * S, D = 9 * SS, NJ = (D + 31) / 32, NE = NJ * 32, TV = S + 4;
    static constexpr int TW = S + 6;                                // table pitch: TV entries + the invalid one, odd (S is odd): the
                                                                    // builders of neighbouring columns / grids hit different banks
    static constexpr int TVP = (TV + 1) / 2;                        // u16 pairs per window-minimum row
    static constexpr int VP = 2 * TVP + 1;                          // words per (grid, column) row pair, odd
    static constexpr int ints_per_warp = 6 * NE + 9 * TW * TW + 9 * S * VP;
    static constexpr size_t bytes = (size_t)NG_WARPS * ints_per_warp * 4;
};

template <int S>
__global__ void __launch_bounds__(NG_WARPS * 32)
pydng_sweep_kernel(const NgSweepParams prm)
{
    using SM = PydngSmem<S>;
    constexpr int SS = SM::SS, D = SM::D, NJ = SM::NJ, NE = SM::NE, TW = SM::TW, TV = SM::TV, TVP = SM::TVP, VP = SM::VP;
    extern __shared__ __align__(16) int pydng_smem[];
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    int* base = pydng_smem + wib * SM::ints_per_warp;
    int* exb = base; int* eyb = base + 2 * NE; int* lcb = base + 4 * NE;
    uint32_t* T = reinterpret_cast<uint32_t*>(base + 6 * NE);
    uint32_t* V = reinterpret_cast<uint32_t*>(base + 6 * NE + 9 * TW * TW);
    const int gw = blockIdx.x * NG_WARPS + wib;
    if (gw >= prm.line_start[prm.n_dirs]) return;
    int k = 0;
    while (gw >= prm.line_start[k + 1]) ++k;
    const int line = gw - prm.line_start[k], r = prm.dir[k];
    const int dx = dir_dx(r), dy = dir_dy(r);
    const int W = prm.W, H = prm.H;
    const size_t N = (size_t)W * H;
    const uint8_t* __restrict__ Cb = prm.cost + blockIdx.y * N * D;
    const int* __restrict__ XYb = prm.XY + blockIdx.y * N * (size_t)(18 * S);
    int16_t* __restrict__ Lb = prm.L[k] + blockIdx.y * N * D;

    int x, y, len;
    if (dy == 0) { y = line; x = dx > 0 ? 0 : W - 1; len = W; }
    else         { x = line; y = dy > 0 ? 0 : H - 1; len = H; }
    const int pstep = dy == 0 ? dx : dy * W;

    int offx[NJ], offy[NJ];
#pragma unroll
    for (int j = 0; j < NJ; ++j) {
        const int d = lane + 32 * j, h = d / SS, rem = d - h * SS, ox = rem / S, oy = rem - ox * S;
        offx[j] = h * S + ox; offy[j] = 9 * S + h * S + oy;
    }
    for (int i = lane; i < 9 * TW * TW; i += 32) T[i] = 0xFFu;
    __syncwarp();

    auto fetch = [&](size_t pix, int (&mx)[NJ], int (&my)[NJ], int (&cc)[NJ], bool& reg) {
        const int* xy = XYb + pix * (size_t)(18 * S);
#pragma unroll
        for (int j = 0; j < NJ; ++j) {
            if (lane + 32 * j < D) { mx[j] = __ldg(xy + offx[j]); my[j] = __ldg(xy + offy[j]); cc[j] = __ldg(Cb + pix * D + lane + 32 * j); }
            else { mx[j] = my[j] = cc[j] = 0; }
        }
        // regular grids: X[h][o] = X[h][0] + o and Y[h][o] = Y[h][0] + o
        bool ok = true;
        for (int i = lane; i < 9 * S; i += 32) {
            const int h = i / S, o = i - h * S;
            ok = ok && __ldg(xy + i) == __ldg(xy + h * S) + o && __ldg(xy + 9 * S + i) == __ldg(xy + 9 * S + h * S) + o;
        }
        reg = ok;
    };

    uint32_t M = 0;
    int cur = 0;
    size_t pix = (size_t)y * W + x;
    int mxn[NJ], myn[NJ], ccn[NJ]; bool regn;
    fetch(pix, mxn, myn, ccn, regn);
    bool grid_ok = false;                          // of the previous pixel
    constexpr bool PF = S <= 7;                    // larger windows: the second register set would spill
    for (int t = 0; t < len; ++t) {
        int mx[NJ], my[NJ], cc[NJ];
        if (!PF && t > 0) fetch(pix, mxn, myn, ccn, regn);
#pragma unroll
        for (int j = 0; j < NJ; ++j) { mx[j] = mxn[j]; my[j] = myn[j]; cc[j] = ccn[j]; }
        const bool reg_cur = __all_sync(0xffffffffu, regn);
        if (PF && t + 1 < len) fetch((size_t)((long long)pix + pstep), mxn, myn, ccn, regn);  // one step ahead
        int* exn = exb + cur * NE; int* eyn = eyb + cur * NE; int* lcn = lcb + cur * NE;
        const int* exp_ = exb + (cur ^ 1) * NE; const int* eyp = eyb + (cur ^ 1) * NE; const int* lcp = lcb + (cur ^ 1) * NE;
        int newL[NJ];
        if (t == 0) {
#pragma unroll
            for (int j = 0; j < NJ; ++j) newL[j] = cc[j];
            M = 0;
        } else {
            const uint32_t far_ = (M + (uint32_t)prm.P2) & 0xFFu;
            uint32_t same[NJ], near_[NJ];
            if (grid_ok) {
                // (1) per grid column (h, ox): minima over the y-window [uy-4, uy] clipped to the grid, with and without its
                //     centre oy = uy - 2, for every uy — all index ranges are compile-time; rows of u16 pairs
                for (int i = lane; i < 9 * S; i += 32) {
                    const int* col = lcp + i * S;                       // i = h * S + ox
                    uint32_t a[S];
#pragma unroll
                    for (int oy = 0; oy < S; ++oy) a[oy] = ((uint32_t)col[oy] + (uint32_t)prm.P1) & 0xFFu;
                    uint32_t v[2 * TVP], vx[2 * TVP];
#pragma unroll
                    for (int uy = 0; uy < 2 * TVP; ++uy) {
                        v[uy] = 255u; vx[uy] = 255u;
#pragma unroll
                        for (int q = 0; q < 5; ++q) {
                            const int oy = uy - 4 + q;
                            if (uy < TV && oy >= 0 && oy < S) { v[uy] = min(v[uy], a[oy]); if (q != 2) vx[uy] = min(vx[uy], a[oy]); }
                        }
                    }
#pragma unroll
                    for (int w = 0; w < TVP; ++w) {
                        V[i * VP + w] = v[2 * w] | (v[2 * w + 1] << 16);
                        V[i * VP + TVP + w] = vx[2 * w] | (vx[2 * w + 1] << 16);
                    }
                }
                __syncwarp();
                // (2) per grid and ux: minimum over the x-window [ux-4, ux] of those rows (the centre column ux - 2 without its
                //     centre cell), u16x2 at a time; the equal cell's own cost where it exists
                for (int i = lane; i < 9 * TV; i += 32) {
                    const int h = i / TV, ux = i - h * TV;
                    uint32_t m[TVP];
#pragma unroll
                    for (int w = 0; w < TVP; ++w) m[w] = 0x00FF00FFu;
#pragma unroll
                    for (int q = 0; q < 5; ++q) {
                        const int ox = ux - 4 + q;
                        if (ox >= 0 && ox < S) {
                            const uint32_t* row = V + (h * S + ox) * VP + (q == 2 ? TVP : 0);
#pragma unroll
                            for (int w = 0; w < TVP; ++w) m[w] = __vminu2(m[w], row[w]);
                        }
                    }
                    const int cx = ux - 2;
                    const bool xin = cx >= 0 && cx < S;
                    const int* ccol = lcp + h * SS + (xin ? cx : 0) * S;
                    uint32_t* t = T + h * TW * TW + ux * TW;
#pragma unroll
                    for (int uy = 0; uy < TV; ++uy) {
                        uint32_t e = (uy & 1) ? m[uy >> 1] >> 16 : m[uy >> 1] & 0xFFFFu;
                        if (uy >= 2 && uy - 2 < S && xin) e |= (((uint32_t)ccol[uy - 2] & 0xFFu) << 16) | 0x01000000u;
                        t[uy] = e;
                    }
                }
                __syncwarp();
                uint32_t acc[NJ], se[NJ];
#pragma unroll
                for (int j = 0; j < NJ; ++j) { acc[j] = far_; se[j] = 0; }
#pragma unroll
                for (int h = 0; h < 9; ++h) {
                    const int X0 = exp_[h * SS] - 2, Y0 = eyp[h * SS] - 2;
                    const uint32_t* Th = T + h * TW * TW;
#pragma unroll
                    for (int j = 0; j < NJ; ++j) {
                        const uint32_t ux = min((uint32_t)(mx[j] - X0), (uint32_t)(S + 4)), uy = min((uint32_t)(my[j] - Y0), (uint32_t)(S + 4));
                        const uint32_t e = Th[ux * TW + uy];
                        acc[j] = __vminu2(acc[j], e);
                        se[j] = e >= 0x01000000u ? e : se[j];              // later grids overwrite: the last match wins (:62-63)
                    }
                }
#pragma unroll
                for (int j = 0; j < NJ; ++j) { near_[j] = acc[j] & 0xFFFFu; same[j] = se[j] ? (se[j] >> 16) & 0xFFu : far_; }
            } else {
#pragma unroll
                for (int j = 0; j < NJ; ++j) { same[j] = far_; near_[j] = far_; }
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
        grid_ok = reg_cur && prm.grid_tables;
        pix = (size_t)((long long)pix + pstep);
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

template <int S>
static int launch_sweep_one(fsgm_ctx* c, dim3 grid, const NgSweepParams& p)
{
    constexpr size_t bytes = PydngSmem<S>::bytes;
    static bool attr_set = false;
    if (bytes > 48 * 1024 && !attr_set) {
        FSGM_CUDA(c, cudaFuncSetAttribute(pydng_sweep_kernel<S>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes));
        attr_set = true;
    }
    pydng_sweep_kernel<S><<<grid, NG_WARPS * 32, bytes, c->stream>>>(p);
    FSGM_LAUNCHED(c);
    return FSGM_OK;
}

static int launch_sweep_S(fsgm_ctx* c, int S, dim3 grid, const NgSweepParams& p)
{
    switch (S) {
    case 1: return launch_sweep_one<1>(c, grid, p);
    case 3: return launch_sweep_one<3>(c, grid, p);
    case 5: return launch_sweep_one<5>(c, grid, p);
    case 7: return launch_sweep_one<7>(c, grid, p);
    case 9: return launch_sweep_one<9>(c, grid, p);
    default: return launch_sweep_one<11>(c, grid, p);
    }
}

int launch_pydng(fsgm_ctx* c, int n, const uint32_t* cen1, const uint32_t* cen2, int W, int H,
                 const double* preMv, int mvW, int mvH, int r, int agg, int subpixel, int P1, int P2,
                 uint8_t* cost, int* XY, int16_t* const* L, uint32_t* Sp32, uint32_t* minC, double* flow)
{
    if (r < 0 || r > PYDNG_MAXR) return fail(c, FSGM_ERR_DOMAIN, "halfSearchWinSize must be in 0..5");
    if (agg < 0 || agg > PYDNG_MAXAGG) return fail(c, FSGM_ERR_DOMAIN, "aggregation radius must be in 0..4");
    const size_t N = (size_t)W * H;
    const int S = 2 * r + 1, D = 9 * S * S;
    {
        StageScope ss(c, ST_PYDNG_COST);
        dim3 grid((unsigned)((N + NG_WARPS - 1) / NG_WARPS), n);
        if (agg == 2 && r == 1) pydng_cost_kernel<2, 1><<<grid, NG_WARPS * 32, 0, c->stream>>>(cen1, cen2, W, H, preMv, mvW, mvH, agg, r, cost, XY);
        else if (agg == 2 && r == 2) pydng_cost_kernel<2, 2><<<grid, NG_WARPS * 32, 0, c->stream>>>(cen1, cen2, W, H, preMv, mvW, mvH, agg, r, cost, XY);
        else pydng_cost_kernel<-1, -1><<<grid, NG_WARPS * 32, 0, c->stream>>>(cen1, cen2, W, H, preMv, mvW, mvH, agg, r, cost, XY);
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
        p.grid_tables = c->pydng_generic ? 0 : 1;
        FSGM_TRY(launch_sweep_S(c, S, grid, p));
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
