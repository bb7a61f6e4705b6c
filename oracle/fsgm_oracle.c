/* fsgm_oracle.c — plain-C restatement of fSGM's hot path.  TEST INFRASTRUCTURE ONLY.
 *
 * This is the CPU checker ("port" oracle) for the CUDA library in fsgm_b200/csrc.  It is a
 * restatement, in different structure, of the arithmetic in the reference's four MEX files;
 * every function cites the reference lines it follows.  It is pinned against the reference's
 * own C++ compiled in place (oracle/_ref, see Makefile) by tests/test_oracle_cpu.py and by the
 * golden fixtures under tests/golden/ (generated from oracle/_ref by tests/golden/make_golden.py).
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
 * load this library.  The product (fsgm_b200/) never links, loads or calls it.
 *
 * Structural difference from the reference (deliberate — it doubles as a proof that the
 * decomposition the GPU uses is legitimate): the reference walks the image once per pass in
 * raster order and advances four ring-buffered paths per pixel; here every direction r is an
 * independent sweep producing its own volume L_r, and Sp = sum_r L_r.  The two are equal because
 * the paths never read each other (calc_cost_sgm.cpp:182-232).
 *
 * Storage types are the reference's: PathCost/CostType = unsigned char with mod-256 truncation
 * wherever the reference stores or std::min<PathCost>()s an int (common.h:4-8).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

typedef uint8_t u8;
typedef uint32_t u32;

static inline int clampi(int v, int lo, int hi) { return v < lo ? lo : (v > hi ? hi : v); } /* common.h:12 */

/* x86-64 double -> int conversion as gcc emits it (cvttsd2si): out-of-range and NaN give INT_MIN. */
static inline int d2i(double v)
{
    if (!(v > -2147483649.0 && v < 2147483648.0)) return INT32_MIN;
    return (int)v;
}
/* x86-64 double -> unsigned: gcc converts through a 64-bit signed cvttsd2si and keeps the low word. */
static inline u32 d2u(double v)
{
    if (!(v > -9223372036854775809.0 && v < 9223372036854775808.0)) return 0u;
    return (u32)(int64_t)v;
}

/* ------------------------------------------------------------------------------------------
 * a-1  5x5 census (common.cpp:3-27): window scanned dy-outer/dx-inner, replicate border,
 * bit = (neighbour >= centre), "code += bit; code <<= 1" after every tap including the last,
 * so tap k (0..24) lands in bit 25-k and bit 0 is always 0.
 * ---------------------------------------------------------------------------------------- */
void orc_census(const u8* img, u32* cen, int W, int H)
{
    for (int y = 0; y < H; ++y)
        for (int x = 0; x < W; ++x) {
            const u8 c = img[(size_t)y * W + x];
            u32 code = 0;
            for (int k = 0; k < 25; ++k) {
                int yy = clampi(y + k / 5 - 2, 0, H - 1), xx = clampi(x + k % 5 - 2, 0, W - 1);
                code |= (u32)(img[(size_t)yy * W + xx] >= c) << (25 - k);
            }
            cen[(size_t)y * W + x] = code;
        }
}

static inline int popc(u32 v) { return __builtin_popcount(v); }

/* ------------------------------------------------------------------------------------------
 * a-2  epipolar matching cost (calc_cost_sgm.cpp:319-412).
 * raw[p][d] = popcount(cen1[p] ^ cen2[q(p,d)]),  q = clamp(round(Pd0 - 1 + (O*vz(d))*u))
 * with vz(d) = r/(1-r), r = 1.0*d/(D+1)*vMax (:339,360-361), products left-associated (:365-366),
 * C `round` (half away from zero) then double->int (:371-372), clamp (:374-375).
 * C[p][d] = (u8)(1.0*sum5x5(raw)/25 + 0.5) with replicate border over the cost volume (:387-404).
 * ---------------------------------------------------------------------------------------- */
void orc_epi_cost_raw(const u32* cen1, const u32* cen2, int W, int H, int D, double vMax,
                      const double* Pd0, const double* dirn, const double* O, u8* raw)
{
    const size_t N = (size_t)W * H;
    const double n = D + 1;
    double* vz = (double*)malloc(sizeof(double) * D);
    for (int d = 0; d < D; ++d) { double r = 1.0 * d / n * vMax; vz[d] = r / (1 - r); }
    for (size_t p = 0; p < N; ++p) {
        const double bx = Pd0[p] - 1, by = Pd0[N + p] - 1, ux = dirn[p], uy = dirn[N + p], off = O[p];
        for (int d = 0; d < D; ++d) {
            double ox = off * vz[d] * ux, oy = off * vz[d] * uy;
            int x2 = clampi(d2i(round(bx + ox)), 0, W - 1);
            int y2 = clampi(d2i(round(by + oy)), 0, H - 1);
            raw[p * D + d] = (u8)popc(cen1[p] ^ cen2[(size_t)y2 * W + x2]);
        }
    }
    free(vz);
}

void orc_box5(const u8* raw, int W, int H, int D, u8* C)
{
    /* separable running sums would be faster; clarity wins here */
    for (int y = 0; y < H; ++y)
        for (int x = 0; x < W; ++x) {
            u8* out = C + ((size_t)y * W + x) * D;
            for (int d = 0; d < D; ++d) {
                unsigned s = 0;
                for (int k = 0; k < 25; ++k) {
                    int yy = clampi(y + k / 5 - 2, 0, H - 1), xx = clampi(x + k % 5 - 2, 0, W - 1);
                    s += raw[((size_t)yy * W + xx) * D + d];
                }
                out[d] = (u8)d2i(1.0 * s / 25 + 0.5);
            }
        }
}

/* ------------------------------------------------------------------------------------------
 * a-3  one 1-D path step (calc_cost_sgm.cpp:33-66), u8 semantics spelled out.
 * Lpre has D entries, preMin is the stored minimum slot Lpre[D] (0 at a path start, :154).
 * Returns the new minimum (starts from 255, :39).
 * ---------------------------------------------------------------------------------------- */
static inline u8 u8min(u8 a, u8 b) { return a < b ? a : b; }

static u8 step1d(u8* L, const u8* Lpre, u8 preMin, const u8* C, int D, int P1, int P2)
{
    u8 newMin = 255;
    const u8 far_ = (u8)(preMin + P2);
    for (int d = 0; d < D; ++d) {
        u8 best = far_;
        if (d > 0) best = u8min(best, (u8)(Lpre[d - 1] + P1));
        if (d < D - 1) best = u8min(best, (u8)(Lpre[d + 1] + P1));
        best = u8min(best, Lpre[d]);
        L[d] = (u8)((C[d] + best) - preMin);
        newMin = u8min(newMin, L[d]);
    }
    return newMin;
}

/* adaptive P2 (calc_cost_sgm.cpp:68-72 thr 25 [compiled off]; calc_pyd_cost_sgm.cpp:91-95 and ng :101-105 thr 50) */
static inline int adaptP2(int P2, int cur, int pre, int thr) { return abs(cur - pre) > thr ? P2 / 8 : P2; }

/* The eight sweep directions in the reference's order of appearance:
 * pass 0: L1 (+1,0)  L3 (0,+1)  L2 (+1,+1)  L4 (-1,+1)      (calc_cost_sgm.cpp:141-148)
 * pass 1: the same four negated                               (:115-123)
 * A path (re)starts wherever the predecessor pixel p - r falls outside the image, which is
 * exactly the reference's x==xstart / y==ystart / x==xend-xstep tests (:152-180). */
static const int DIRS[8][2] = { {1, 0}, {0, 1}, {1, 1}, {-1, 1}, {-1, 0}, {0, -1}, {-1, -1}, {1, -1} };

/* which of the 8 directions are active: passes in {1,2}, diag in {0,1} */
static int dir_enabled(int r, int passes, int diag)
{
    if (r >= 4 && passes < 2) return 0;
    if ((r & 3) >= 2 && !diag) return 0;
    return 1;
}

/* a-4 (sweep part)  one direction over the whole image; L_out[N][D]; Lmin scratch [N]. */
void orc_sweep1d(const u8* C, const u8* I1, int W, int H, int D, int P1, int P2, int adaptive_thr,
                 int r, u8* L_out)
{
    const int dx = DIRS[r][0], dy = DIRS[r][1];
    const size_t N = (size_t)W * H;
    u8* Lmin = (u8*)malloc(N);
    const int forward = (dy > 0) || (dy == 0 && dx > 0);
    for (size_t i = 0; i < N; ++i) {
        size_t p = forward ? i : N - 1 - i;
        int x = (int)(p % W), y = (int)(p / W), xp = x - dx, yp = y - dy;
        u8* L = L_out + p * D;
        const u8* Cp = C + p * D;
        if (xp < 0 || xp >= W || yp < 0 || yp >= H) {
            memcpy(L, Cp, D);
            Lmin[p] = 0;                                    /* quirk: not min(C) (:154,:164) */
        } else {
            size_t q = (size_t)yp * W + xp;
            int p2 = adaptive_thr > 0 ? adaptP2(P2, I1[p], I1[q], adaptive_thr) : P2;
            Lmin[p] = step1d(L, L_out + q * D, Lmin[q], Cp, D, P1, p2);
        }
    }
    free(Lmin);
}

/* a-4  aggregated volume Sp[N][D] = sum over enabled directions (calc_cost_sgm.cpp:227-232). */
void orc_epi_aggregate(const u8* C, const u8* I1, int W, int H, int D, int P1, int P2,
                       int passes, int diag, int adaptive, u32* Sp)
{
    const size_t V = (size_t)W * H * D;
    u8* L = (u8*)malloc(V);
    memset(Sp, 0, V * sizeof(u32));
    for (int r = 0; r < 8; ++r) {
        if (!dir_enabled(r, passes, diag)) continue;
        orc_sweep1d(C, I1, W, H, D, P1, P2, adaptive ? 25 : 0, r, L);
        for (size_t i = 0; i < V; ++i) Sp[i] += L[i];
    }
    free(L);
}

/* a-4 (WTA + subpixel, calc_cost_sgm.cpp:259-308): first minimum (strict <); refinement only for
 * 1 < idx < D — so label 1 is never refined and idx == D-1 reads Sp[p][D], i.e. the NEXT pixel's
 * label 0 (for the very last pixel the reference reads past its buffer; here that value is 0,
 * matching the zero-padded shim allocation the _ref build uses). */
void orc_epi_wta(const u32* Sp, int W, int H, int D, int subpixel, u32* bestD, u32* minC)
{
    const size_t N = (size_t)W * H;
    for (size_t p = 0; p < N; ++p) {
        const u32* s = Sp + p * D;
        u32 best = s[0], idx = 0;
        for (int d = 1; d < D; ++d) if (s[d] < best) { best = s[d]; idx = d; }
        minC[p] = best;
        if (!subpixel) { bestD[p] = idx; continue; }
        if (idx > 1 && idx < (u32)D) {
            double c_1 = s[idx - 1], c = s[idx];
            double c1 = (idx + 1 < (u32)D) ? s[idx + 1] : (p + 1 < N ? Sp[(p + 1) * D] : 0.0);
            double sub = idx;
            if (c1 < c_1) sub = sub + (c1 - c_1) / (c - c_1) / 2.0;
            else          sub = sub + (c1 - c_1) / (c - c1) / 2.0;
            bestD[p] = d2u(sub * 256);
        } else
            bestD[p] = idx * 256;
    }
}

/* a-5  vz-index label (Q24.8) -> pixel disparity (Q24.8), calc_cost_sgm.cpp:414-426, n = D+1 (:593). */
void orc_vz_to_disp(u32* bestD, int W, int H, const double* O, double vMax, int D)
{
    const size_t N = (size_t)W * H;
    const int n = D + 1;
    for (size_t p = 0; p < N; ++p) {
        double d = (double)bestD[p] / 256;
        double r = d / n * vMax;
        double vz = r / (1 - r);
        bestD[p] = d2u((O[p] * vz) * 256);
    }
}

/* N3  forward/backward consistency (calc_cost_sgm.cpp:488-536) with calc_disp_from_first (:429-486): the call the
 * reference ships commented out (:589-590).  D1 = x256 label map before the vz conversion, n = D+1.
 * Restated as two gathers-by-scatter over a "largest offer" map instead of the reference's in-place update rule
 * ("== INVALID || D2 < D1" keeps the maximum, so the result does not depend on the visiting order as long as
 * D1 < INVALID_DISPARITY = 512<<8). */
#define ORC_INVALID_DISPARITY (512u << 8)
static inline double fb_disp(u32 D1, double O, double vMax, int n, int use_vzind)
{
    double d = (double)D1 / 256.0;
    if (use_vzind) { double r = d / n * vMax; d = O * (r / (1 - r)); }
    return d;
}
void orc_fb_check(const u32* D1, int W, int H, const double* Pd0, const double* dirn, const double* O,
                  double vMax, int n, int thr, int use_vzind, u8* conf, u32* D2)
{
    const size_t N = (size_t)W * H;
    long long* top = (long long*)malloc(N * sizeof(long long));          /* -1 = nothing landed here */
    for (size_t i = 0; i < N; ++i) top[i] = -1;
    for (size_t i = 0; i < N; ++i) {
        const double d = fb_disp(D1[i], O[i], vMax, n, use_vzind);
        const int px = d2i((Pd0[i] - 1) + d * dirn[i]), py = d2i((Pd0[N + i] - 1) + d * dirn[N + i]);   /* truncation, :459-460 */
        for (int k = 0; k < 4; ++k) {
            const long long tx = (long long)px + (k & 1), ty = (long long)py + (k >> 1);
            if (tx < 0 || tx >= W || ty < 0 || ty >= H) continue;
            if (top[ty * W + tx] < (long long)D1[i]) top[ty * W + tx] = D1[i];
        }
    }
    for (size_t i = 0; i < N; ++i) D2[i] = top[i] < 0 ? ORC_INVALID_DISPARITY : (u32)top[i];
    free(top);
    for (size_t i = 0; i < N; ++i) {
        const double d = fb_disp(D1[i], O[i], vMax, n, use_vzind);
        const int px = d2i(round((Pd0[i] - 1) + d * dirn[i])), py = d2i(round((Pd0[N + i] - 1) + d * dirn[N + i]));   /* :516-517 */
        u8 ok = 1;
        if (px < 0 || px > W - 1 || py < 0 || py > H - 1) ok = 0;
        else if (D2[(size_t)py * W + px] == ORC_INVALID_DISPARITY) ok = 0;
        else if (abs((int)D1[i] - (int)D2[(size_t)py * W + px]) > thr) ok = 0;
        conf[i] = ok;
    }
}

/* a-6 with the commented-out call switched on: labels, check, then the vz conversion (:581-594) */
void orc_epi_fb(const u8* I1, const u8* I2, int W, int H, int D, double vMax,
                const double* Pd0, const double* dirn, const double* O, int P1, int P2, int paths, int thr,
                u32* bestD, u32* minC, u8* conf, u32* bestD2)
{
    const size_t N = (size_t)W * H, V = N * D;
    u32 *cen1 = (u32*)malloc(N * 4), *cen2 = (u32*)malloc(N * 4), *Sp = (u32*)malloc(V * 4);
    u8 *raw = (u8*)malloc(V), *C = (u8*)malloc(V);
    orc_census(I1, cen1, W, H);
    orc_census(I2, cen2, W, H);
    orc_epi_cost_raw(cen1, cen2, W, H, D, vMax, Pd0, dirn, O, raw);
    orc_box5(raw, W, H, D, C);
    orc_epi_aggregate(C, I1, W, H, D, P1, P2, 2, paths == 8, 0, Sp);
    orc_epi_wta(Sp, W, H, D, 1, bestD, minC);
    orc_fb_check(bestD, W, H, Pd0, dirn, O, vMax, D + 1, thr, 1, conf, bestD2);
    orc_vz_to_disp(bestD, W, H, O, vMax, D);
    free(cen1); free(cen2); free(Sp); free(raw); free(C);
}

/* a-6  whole gateway (calc_cost_sgm.cpp:539-598); paths = 4 (as shipped, :102-104) or 8.
 * Any of the stage outputs may be NULL. */
void orc_epi(const u8* I1, const u8* I2, int W, int H, int D, double vMax,
             const double* Pd0, const double* dirn, const double* O, int P1, int P2, int paths,
             u32* bestD, u32* minC, u32* cen1_o, u32* cen2_o, u8* raw_o, u8* C_o, u32* Sp_o)
{
    const size_t N = (size_t)W * H, V = N * D;
    u32 *cen1 = (u32*)malloc(N * 4), *cen2 = (u32*)malloc(N * 4), *Sp = (u32*)malloc(V * 4);
    u8 *raw = (u8*)malloc(V), *C = (u8*)malloc(V);
    orc_census(I1, cen1, W, H);
    orc_census(I2, cen2, W, H);
    orc_epi_cost_raw(cen1, cen2, W, H, D, vMax, Pd0, dirn, O, raw);
    orc_box5(raw, W, H, D, C);
    orc_epi_aggregate(C, I1, W, H, D, P1, P2, 2, paths == 8, 0, Sp);
    orc_epi_wta(Sp, W, H, D, 1, bestD, minC);
    orc_vz_to_disp(bestD, W, H, O, vMax, D);
    if (cen1_o) memcpy(cen1_o, cen1, N * 4);
    if (cen2_o) memcpy(cen2_o, cen2, N * 4);
    if (raw_o) memcpy(raw_o, raw, V);
    if (C_o) memcpy(C_o, C, V);
    if (Sp_o) memcpy(Sp_o, Sp, V * 4);
    free(cen1); free(cen2); free(Sp); free(raw); free(C);
}

/* ==========================================================================================
 * pyramidal 2-D-window variant (calc_pyd_cost_sgm.cpp)
 * ======================================================================================== */

/* a-7  cost over a (2rx+1)x(2ry+1) window centred on the prior flow (calc_pyd_cost_sgm.cpp:374-437).
 * Labels: offx outer, offy inner (:392-393).  The prior mv is taken at the CENTRE pixel for the
 * whole aggregation window (:388-389).  y2 = (int)(1.0*(offy+y1) + mvy + 0.5) truncates toward
 * zero (:415-416).  Out-of-image current OR reference sample contributes the constant 5 (:382,405-421). */
void orc_pyd_cost(const u32* cen1, const u32* cen2, int W, int H, const double* preMv, int mvW, int mvH,
                  int agg, int rx, int ry, u8* C)
{
    const double *mvxP = preMv, *mvyP = preMv + (size_t)mvW * mvH;
    const int wp = (2 * agg + 1) * (2 * agg + 1);
    const int D = (2 * rx + 1) * (2 * ry + 1);
    for (int y = 0; y < H; ++y)
        for (int x = 0; x < W; ++x) {
            const double mvx = mvxP[(size_t)mvW * y + x], mvy = mvyP[(size_t)mvW * y + x];
            u8* out = C + ((size_t)y * W + x) * D;
            for (int d = 0; d < D; ++d) {
                const int offx = d / (2 * ry + 1) - rx, offy = d % (2 * ry + 1) - ry;
                unsigned s = 0;
                for (int ay = -agg; ay <= agg; ++ay)
                    for (int ax = -agg; ax <= agg; ++ax) {
                        int y1 = y + ay, x1 = x + ax;
                        if (y1 < 0 || y1 > H - 1 || x1 < 0 || x1 > W - 1) { s += 5; continue; }
                        int y2 = d2i(1.0 * (offy + y1) + mvy + 0.5);
                        int x2 = d2i(1.0 * (offx + x1) + mvx + 0.5);
                        if (y2 < 0 || y2 > H - 1 || x2 < 0 || x2 > W - 1) { s += 5; continue; }
                        s += popc(cen1[(size_t)W * y1 + x1] ^ cen2[(size_t)W * y2 + x2]);
                    }
                out[d] = (u8)d2i((1.0 * s / wp) + 0.5);
            }
        }
}

/* a-8  2-D-label path step (calc_pyd_cost_sgm.cpp:34-89).  (ddx,ddy) = mv(cur) - mv(prev on path).
 * The predecessor label of (sx,sy) is (xpre,ypre) = ((int)(sx+ddx+0.5), (int)(sy+ddy+0.5)) (:46-47);
 * same-label term only if it lies inside the window (:56-59); P1 term = min over the 5x5
 * neighbourhood of (xpre,ypre), centre excluded, inside-window only (:61-76). */
static u8 step2d(u8* L, const u8* Lpre, u8 preMin, const u8* C, double ddx, double ddy,
                 int Sx, int Sy, int P1, int P2)
{
    u8 newMin = 255;
    const u8 far_ = (u8)(preMin + P2);
    for (int sx = 0; sx < Sx; ++sx)
        for (int sy = 0; sy < Sy; ++sy) {
            int ypre = d2i(sy + ddy + 0.5), xpre = d2i(sx + ddx + 0.5);
            u8 same = far_, near_ = far_;
            if (xpre >= 0 && xpre < Sx && ypre >= 0 && ypre < Sy) same = Lpre[xpre * Sy + ypre];
            for (int k = -2; k <= 2; ++k)
                for (int m = -2; m <= 2; ++m) {
                    if (!k && !m) continue;
                    /* long arithmetic: xpre may be INT_MIN after an out-of-range conversion */
                    long ty = (long)ypre + k, tx = (long)xpre + m;
                    if (tx >= 0 && tx < Sx && ty >= 0 && ty < Sy)
                        near_ = u8min(near_, (u8)(Lpre[tx * Sy + ty] + P1));
                }
            u8 best = u8min(u8min(far_, same), near_);
            int d = sx * Sy + sy;
            L[d] = (u8)((C[d] + best) - preMin);
            newMin = u8min(newMin, L[d]);
        }
    return newMin;
}

void orc_sweep2d(const u8* C, const u8* I1, int W, int H, const double* preMv, int mvW, int mvH,
                 int Sx, int Sy, int P1, int P2, int adaptive, int r, u8* L_out)
{
    const int dx = DIRS[r][0], dy = DIRS[r][1], D = Sx * Sy;
    const size_t N = (size_t)W * H;
    const double *mvxP = preMv, *mvyP = preMv + (size_t)mvW * mvH;
    u8* Lmin = (u8*)malloc(N);
    const int forward = (dy > 0) || (dy == 0 && dx > 0);
    for (size_t i = 0; i < N; ++i) {
        size_t p = forward ? i : N - 1 - i;
        int x = (int)(p % W), y = (int)(p / W), xp = x - dx, yp = y - dy;
        u8* L = L_out + p * D;
        const u8* Cp = C + p * D;
        if (xp < 0 || xp >= W || yp < 0 || yp >= H) {
            memcpy(L, Cp, D);
            Lmin[p] = 0;
        } else {
            size_t q = (size_t)yp * W + xp;
            /* prior difference along the path, indexed with the mv map's own stride (:213-254) */
            double ddx = mvxP[(size_t)y * mvW + x] - mvxP[(size_t)yp * mvW + xp];
            double ddy = mvyP[(size_t)y * mvW + x] - mvyP[(size_t)yp * mvW + xp];
            int p2 = adaptive ? adaptP2(P2, I1[p], I1[q], 50) : P2;
            Lmin[p] = step2d(L, L_out + q * D, Lmin[q], Cp, ddx, ddy, Sx, Sy, P1, p2);
        }
    }
    free(Lmin);
}

/* a-9  sgm2d: sum of sweeps, WTA, per-axis subpixel (calc_pyd_cost_sgm.cpp:114-372). */
void orc_pyd_sgm(const u8* C, const u8* I1, int W, int H, const double* preMv, int mvW, int mvH,
                 int Sx, int Sy, int P1, int P2, int subpixel, int diag, int passes, int adaptive,
                 u32* bestD, u32* minC, double* mvSub, u32* Sp)
{
    const int D = Sx * Sy;
    const size_t N = (size_t)W * H, V = N * D;
    u8* L = (u8*)malloc(V);
    memset(Sp, 0, V * 4);
    for (int r = 0; r < 8; ++r) {
        /* totalPass is a loop bound (:142): 0 passes leaves Sp at zero; >2 repeats pass 1 */
        int reps = 0;
        if (r < 4) reps = passes >= 1; else reps = passes >= 2 ? passes - 1 : 0;
        if ((r & 3) >= 2 && !diag) reps = 0;
        if (!reps) continue;
        orc_sweep2d(C, I1, W, H, preMv, mvW, mvH, Sx, Sy, P1, P2, adaptive, r, L);
        for (size_t i = 0; i < V; ++i) Sp[i] += (u32)reps * L[i];
    }
    free(L);
    memset(mvSub, 0, N * 2 * sizeof(double));
    for (size_t p = 0; p < N; ++p) {
        const u32* s = Sp + p * D;
        u32 best = s[0], idx = 0;
        for (int d = 1; d < D; ++d) if (s[d] < best) { best = s[d]; idx = d; }
        minC[p] = best; bestD[p] = idx;
        if (!subpixel) continue;
        double c0 = s[idx];
        int lx = idx / Sy, ly = idx % Sy;
        if (ly > 0 && ly < Sy - 1) {                                   /* :336-347 */
            double a = s[idx - 1], b = s[idx + 1];
            mvSub[N + p] = (b < a) ? (b - a) / (c0 - a) / 2.0 : (b - a) / (c0 - b) / 2.0;
        }
        if (lx > 0 && lx < Sx - 1) {                                   /* :349-360 */
            double a = s[idx - Sy], b = s[idx + Sy];
            mvSub[p] = (b < a) ? (b - a) / (c0 - a) / 2.0 : (b - a) / (c0 - b) / 2.0;
        }
    }
}

/* a-10  gateway (calc_pyd_cost_sgm.cpp:439-510). */
void orc_pyd(const u8* I1, const u8* I2, int W, int H, const double* preMv, int mvW, int mvH,
             int rx, int ry, int agg, int subpixel, int P1, int P2, int diag, int passes, int adaptive,
             u32* bestD, u32* minC, double* mvSub, u32* cen1_o, u32* cen2_o, u8* C_o, u32* Sp_o)
{
    const int Sx = 2 * rx + 1, Sy = 2 * ry + 1, D = Sx * Sy;
    const size_t N = (size_t)W * H, V = N * D;
    u32 *cen1 = (u32*)malloc(N * 4), *cen2 = (u32*)malloc(N * 4), *Sp = (u32*)malloc(V * 4);
    u8* C = (u8*)malloc(V);
    orc_census(I1, cen1, W, H);
    orc_census(I2, cen2, W, H);
    orc_pyd_cost(cen1, cen2, W, H, preMv, mvW, mvH, agg, rx, ry, C);
    orc_pyd_sgm(C, I1, W, H, preMv, mvW, mvH, Sx, Sy, P1, P2, subpixel, diag, passes, adaptive, bestD, minC, mvSub, Sp);
    if (cen1_o) memcpy(cen1_o, cen1, N * 4);
    if (cen2_o) memcpy(cen2_o, cen2, N * 4);
    if (C_o) memcpy(C_o, C, V);
    if (Sp_o) memcpy(Sp_o, Sp, V * 4);
    free(cen1); free(cen2); free(Sp); free(C);
}

/* ==========================================================================================
 * explicit-candidate ("neighbour guided") variants
 * ======================================================================================== */
typedef struct { int mvx, mvy, cost; } Cand;                       /* calc_cost_sgm_ng.cpp:39-44 */

/* O(D^2) compatibility search shared by a-13 and a-16: for candidate (mvx,mvy) find
 *   same = cost of the LAST previous entry with identical mv (the reference overwrites, :71-72)
 *   near = min over previous entries with |dmv| <= 2 in both axes, not identical, of cost + P1 (:73-74)
 * all in u8. */
static inline u8 compat_best(const Cand* Lpre, int D, int mvx, int mvy, u8 far_, int P1)
{
    u8 same = far_, near_ = far_;
    for (int k = 0; k < D; ++k) {
        int ax = Lpre[k].mvx, ay = Lpre[k].mvy;
        if (ax == mvx && ay == mvy) same = (u8)Lpre[k].cost;
        else if (abs(mvx - ax) <= 2 && abs(mvy - ay) <= 2) near_ = u8min(near_, (u8)(Lpre[k].cost + P1));
    }
    return u8min(u8min(far_, same), near_);
}

/* a-13  ng path step (calc_cost_sgm_ng.cpp:46-98).  L has D+2 entries; the two extra slots hold the
 * running top-2 (strict <, stable shift).  Only their .cost is reset to 255 first (:54-55): the mv
 * fields stay whatever they were. */
static void ng_step(Cand* L, const Cand* Lpre, const Cand* C, int D, int P1, int P2)
{
    const u8 preMin = (u8)Lpre[D].cost;
    const u8 far_ = (u8)(preMin + P2);
    L[D].cost = 255; L[D + 1].cost = 255;
    for (int d = 0; d < D; ++d) {
        u8 best = compat_best(Lpre, D, C[d].mvx, C[d].mvy, far_, P1);
        L[d].cost = (C[d].cost + best) - preMin;            /* int, NOT truncated (:80) */
        L[d].mvx = C[d].mvx; L[d].mvy = C[d].mvy;
        if (L[d].cost < L[D].cost) { L[D + 1] = L[D]; L[D] = L[d]; }
        else if (L[d].cost < L[D + 1].cost) L[D + 1] = L[d];
    }
}

/* a-12  candidates for one pixel (calc_cost_sgm_ng.cpp:122-186).  hints[l] points at the two top
 * slots of direction l's *current* ring entry, i.e. stale data from an earlier pixel (:276-277).
 * 2 x rand() per direction, in order mvx then mvy (:148-149). */
static void ng_candidates(Cand* out, int x, int y, const u32* cen1, const u32* cen2, int W, int H,
                          const Cand* const hints[4])
{
    int c = 0;
    for (int l = 0; l < 4; ++l)
        for (int i = 0; i < 3; ++i) {
            int hx, hy;
            if (i < 2) { hx = hints[l][i].mvx; hy = hints[l][i].mvy; }
            else { hx = rand() % 256 - 128; hy = rand() % 128 - 64; }
            for (int oy = -1; oy <= 1; ++oy)
                for (int ox = -1; ox <= 1; ++ox) {
                    unsigned s = 0;
                    for (int ay = -2; ay <= 2; ++ay)
                        for (int ax = -2; ax <= 2; ++ax) {
                            int y1 = clampi(y + ay, 0, H - 1), x1 = clampi(x + ax, 0, W - 1);
                            /* int adds wrap like the hardware does; hints are small in practice */
                            int y2 = clampi((int)((u32)(oy + y1) + (u32)hy), 0, H - 1);
                            int x2 = clampi((int)((u32)(ox + x1) + (u32)hx), 0, W - 1);
                            s += popc(cen1[(size_t)W * y1 + x1] ^ cen2[(size_t)W * y2 + x2]);
                        }
                    out[c].cost = d2i(1.0 * s / 25 + 0.5);
                    out[c].mvx = (int)((u32)hx + (u32)ox);
                    out[c].mvy = (int)((u32)hy + (u32)oy);
                    ++c;
                }
        }
}

/* a-14  calc_cost_sgm_ng.cpp:188-419.  One forward raster pass, 4 directions, adaptive P2 (thr 50).
 * The raster dependence is intrinsic (hints come from the ring buffers), so unlike the fixed-label
 * variants this one cannot be restated per direction.  Ring buffers: L1 two slots swapped per pixel
 * and never reset per row (:247-248,:365-367); L2/L3/L4 two rows swapped per row (:371-385); all
 * zero-initialised (:204-208).  Uses libc rand(): caller seeds (srand) first. */
void orc_ng(const u8* I1, const u8* I2, int W, int H, int P1, int P2, unsigned seed,
            u32* minC, double* flow, int32_t* Cent_o, u32* Sp_o)
{
    enum { D = 108, E = D + 2 };
    const size_t N = (size_t)W * H;
    u32 *cen1 = (u32*)malloc(N * 4), *cen2 = (u32*)malloc(N * 4);
    orc_census(I1, cen1, W, H);
    orc_census(I2, cen2, W, H);
    Cand* Cv = (Cand*)malloc(N * D * sizeof(Cand));
    u32* Sp = (u32*)calloc(N * D, 4);
    Cand* ring1 = (Cand*)calloc(2 * E, sizeof(Cand));
    Cand* rows[3];                                   /* index 0: L2, 1: L3, 2: L4; each 2 rows */
    for (int k = 0; k < 3; ++k) rows[k] = (Cand*)calloc((size_t)2 * W * E, sizeof(Cand));
    srand(seed);
    int cur1 = 1;                                    /* L1: Pre = slot 0, Cur = slot 1 (:247-248) */
    for (int y = 0; y < H; ++y) {
        const int curRow = (y + 1) & 1, preRow = y & 1;   /* row rings start Pre=0,Cur=1 and swap per row */
        for (int x = 0; x < W; ++x) {
            const size_t p = (size_t)y * W + x;
            Cand* L1c = ring1 + (size_t)cur1 * E;
            const Cand* L1p = ring1 + (size_t)(cur1 ^ 1) * E;
            Cand* L2c = rows[0] + ((size_t)curRow * W + x) * E;
            Cand* L3c = rows[1] + ((size_t)curRow * W + x) * E;
            Cand* L4c = rows[2] + ((size_t)curRow * W + x) * E;
            Cand* Cp = Cv + p * D;
            const Cand* hints[4] = { L1c + D, L2c + D, L3c + D, L4c + D };
            ng_candidates(Cp, x, y, cen1, cen2, W, H, hints);
            const int startX = (x == 0), startY = (y == 0), startR = (x == W - 1);
            if (startX) { memcpy(L1c, Cp, D * sizeof(Cand)); L1c[D].cost = 0; }
            if (startX || startY) { memcpy(L2c, Cp, D * sizeof(Cand)); L2c[D].cost = 0; }
            if (startY) { memcpy(L3c, Cp, D * sizeof(Cand)); L3c[D].cost = 0; }
            if (startY || startR) { memcpy(L4c, Cp, D * sizeof(Cand)); L4c[D].cost = 0; }
            const int cur = I1[p];
            if (!startX) ng_step(L1c, L1p, Cp, D, P1, adaptP2(P2, cur, I1[p - 1], 50));
            if (!startY) ng_step(L3c, rows[1] + ((size_t)preRow * W + x) * E, Cp, D, P1, adaptP2(P2, cur, I1[p - W], 50));
            if (!startX && !startY)
                ng_step(L2c, rows[0] + ((size_t)preRow * W + x - 1) * E, Cp, D, P1, adaptP2(P2, cur, I1[p - W - 1], 50));
            if (!startR && !startY)
                ng_step(L4c, rows[2] + ((size_t)preRow * W + x + 1) * E, Cp, D, P1, adaptP2(P2, cur, I1[p - W + 1], 50));
            for (int d = 0; d < D; ++d)
                Sp[p * D + d] += (u32)(L1c[d].cost + L3c[d].cost) + (u32)(L2c[d].cost + L4c[d].cost);
            cur1 ^= 1;
        }
    }
    for (size_t p = 0; p < N; ++p) {                                     /* :389-407 */
        const u32* s = Sp + p * D;
        u32 best = s[0], idx = 0;
        for (int d = 1; d < D; ++d) if (s[d] < best) { best = s[d]; idx = d; }
        minC[p] = best;
        flow[p] = Cv[p * D + idx].mvx; flow[N + p] = Cv[p * D + idx].mvy;
    }
    if (Cent_o) memcpy(Cent_o, Cv, N * D * sizeof(Cand));
    if (Sp_o) memcpy(Sp_o, Sp, N * D * 4);
    free(cen1); free(cen2); free(Cv); free(Sp); free(ring1);
    for (int k = 0; k < 3; ++k) free(rows[k]);
}

/* a-15  pyd_ng candidate volume (calc_pyd_cost_sgm_ng.cpp:370-446): hints = prior mv on a 3x3 grid
 * with stride 8 (dy outer, dx inner, clamped to the mv map, :390-396), each expanded by (2r+1)^2
 * offsets (offx outer, :399-400).  NO +0.5 in the coordinates (:417-418); const cost 5 outside;
 * entry mv = (int)(mv + off) truncated (:433-434). */
void orc_pydng_cost(const u32* cen1, const u32* cen2, int W, int H, const double* preMv, int mvW, int mvH,
                    int agg, int r, Cand* Cv)
{
    const double *mvxP = preMv, *mvyP = preMv + (size_t)mvW * mvH;
    const int wp = (2 * agg + 1) * (2 * agg + 1), S = 2 * r + 1, D = 9 * S * S;
    for (int y = 0; y < H; ++y)
        for (int x = 0; x < W; ++x) {
            Cand* out = Cv + ((size_t)y * W + x) * D;
            int d = 0;
            for (int hy = -8; hy <= 8; hy += 8)
                for (int hx = -8; hx <= 8; hx += 8) {
                    int yn = clampi(y + hy, 0, mvH - 1), xn = clampi(x + hx, 0, mvW - 1);
                    double mvx = mvxP[(size_t)mvW * yn + xn], mvy = mvyP[(size_t)mvW * yn + xn];
                    for (int offx = -r; offx <= r; ++offx)
                        for (int offy = -r; offy <= r; ++offy) {
                            unsigned s = 0;
                            for (int ay = -agg; ay <= agg; ++ay)
                                for (int ax = -agg; ax <= agg; ++ax) {
                                    int y1 = y + ay, x1 = x + ax;
                                    if (y1 < 0 || y1 > H - 1 || x1 < 0 || x1 > W - 1) { s += 5; continue; }
                                    int y2 = d2i((offy + y1) + mvy), x2 = d2i((offx + x1) + mvx);
                                    if (y2 < 0 || y2 > H - 1 || x2 < 0 || x2 > W - 1) { s += 5; continue; }
                                    s += popc(cen1[(size_t)W * y1 + x1] ^ cen2[(size_t)W * y2 + x2]);
                                }
                            out[d].cost = d2i((1.0 * s / wp) + 0.5);
                            out[d].mvx = d2i(mvx + offx);
                            out[d].mvy = d2i(mvy + offy);
                            ++d;
                        }
                }
        }
}

/* a-16  pyd_ng sweeps (calc_pyd_cost_sgm_ng.cpp:39-78, :101-306): 2 passes, horizontal + vertical
 * only (:120-122), no adaptive P2; running min kept as u8 (:45,:74,:77); L.cost is an int. */
static void pydng_sweep(const Cand* Cv, int W, int H, int D, int P1, int P2, int r, int* Lc /*[N][D]*/)
{
    const int dx = DIRS[r][0], dy = DIRS[r][1];
    const size_t N = (size_t)W * H;
    u8* Lmin = (u8*)malloc(N);
    Cand* pre = (Cand*)malloc(sizeof(Cand) * D);
    const int forward = (dy > 0) || (dy == 0 && dx > 0);
    for (size_t i = 0; i < N; ++i) {
        size_t p = forward ? i : N - 1 - i;
        int x = (int)(p % W), y = (int)(p / W), xp = x - dx, yp = y - dy;
        const Cand* Cp = Cv + p * D;
        int* L = Lc + p * D;
        if (xp < 0 || xp >= W || yp < 0 || yp >= H) {
            for (int d = 0; d < D; ++d) L[d] = Cp[d].cost;
            Lmin[p] = 0;
        } else {
            size_t q = (size_t)yp * W + xp;
            for (int d = 0; d < D; ++d) { pre[d].mvx = Cv[q * D + d].mvx; pre[d].mvy = Cv[q * D + d].mvy; pre[d].cost = Lc[q * D + d]; }
            const u8 preMin = Lmin[q], far_ = (u8)(preMin + P2);
            u8 m = 255;
            for (int d = 0; d < D; ++d) {
                u8 best = compat_best(pre, D, Cp[d].mvx, Cp[d].mvy, far_, P1);
                L[d] = (Cp[d].cost + best) - preMin;
                m = u8min(m, (u8)L[d]);
            }
            Lmin[p] = m;
        }
    }
    free(Lmin); free(pre);
}

/* a-17  census-based subpixel (calc_pyd_cost_sgm_ng.cpp:308-368).  Note the `continue`s: when the x
 * refinement is skipped the y refinement is skipped too (:337-338). */
void orc_pydng_subpixel(double* flow, const u32* cen1, const u32* cen2, int W, int H)
{
    const size_t N = (size_t)W * H;
    for (int y = 0; y < H; ++y)
        for (int x = 0; x < W; ++x) {
            size_t p = (size_t)y * W + x;
            u32 c1 = cen1[p];
            int tx = d2i(flow[p] + x), ty = d2i(flow[N + p] + y);
            if (!(tx > 1 && tx < W - 1 && ty > 1 && ty < H - 1)) continue;
            double c0 = popc(c1 ^ cen2[(size_t)ty * W + tx]);
            double a = popc(c1 ^ cen2[(size_t)ty * W + tx - 1]), b = popc(c1 ^ cen2[(size_t)ty * W + tx + 1]);
            if (c0 >= a || c0 >= b) continue;
            flow[p] += (b < a) ? (b - a) / (c0 - a) / 2.0 : (b - a) / (c0 - b) / 2.0;
            a = popc(c1 ^ cen2[(size_t)(ty - 1) * W + tx]); b = popc(c1 ^ cen2[(size_t)(ty + 1) * W + tx]);
            if (c0 >= a || c0 >= b) continue;
            flow[N + p] += (b < a) ? (b - a) / (c0 - a) / 2.0 : (b - a) / (c0 - b) / 2.0;
        }
}

/* a-18  gateway (calc_pyd_cost_sgm_ng.cpp:448-523): r = halfSearchWinSize, agg = (int)aggSize/2 (:488-490). */
void orc_pydng(const u8* I1, const u8* I2, int W, int H, const double* preMv, int mvW, int mvH,
               int r, int aggSize, int subpixel, int P1, int P2,
               u32* minC, double* flow, int32_t* Cent_o, u32* Sp_o)
{
    const int S = 2 * r + 1, D = 9 * S * S, agg = aggSize / 2;
    const size_t N = (size_t)W * H, V = N * D;
    u32 *cen1 = (u32*)malloc(N * 4), *cen2 = (u32*)malloc(N * 4);
    orc_census(I1, cen1, W, H);
    orc_census(I2, cen2, W, H);
    Cand* Cv = (Cand*)malloc(V * sizeof(Cand));
    orc_pydng_cost(cen1, cen2, W, H, preMv, mvW, mvH, agg, r, Cv);
    u32* Sp = (u32*)calloc(V, 4);
    int* L = (int*)malloc(V * sizeof(int));
    static const int order[4] = { 0, 1, 4, 5 };       /* L1, L3 forward; L1, L3 backward */
    for (int k = 0; k < 4; ++k) {
        pydng_sweep(Cv, W, H, D, P1, P2, order[k], L);
        for (size_t i = 0; i < V; ++i) Sp[i] += (u32)L[i];
    }
    for (size_t p = 0; p < N; ++p) {                                     /* :281-299 */
        const u32* s = Sp + p * D;
        u32 best = s[0], idx = 0;
        for (int d = 1; d < D; ++d) if (s[d] < best) { best = s[d]; idx = d; }
        minC[p] = best;
        flow[p] = Cv[p * D + idx].mvx; flow[N + p] = Cv[p * D + idx].mvy;
    }
    if (subpixel) orc_pydng_subpixel(flow, cen1, cen2, W, H);
    if (Cent_o) memcpy(Cent_o, Cv, V * sizeof(Cand));
    if (Sp_o) memcpy(Sp_o, Sp, V * 4);
    free(cen1); free(cen2); free(Cv); free(Sp); free(L);
}
