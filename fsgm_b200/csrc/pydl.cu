// Lane = path aggregation for the pyramidal 2-D-window variant — sgm2d() / sgm_step() of the reference
// (calc_pyd_cost_sgm.cpp:114-372, :34-89) with ONE THREAD PER PATH: a lane owns a scanline of one direction and keeps the whole
// (2rx+1) x (2ry+1) label grid of its path to itself, so the 5x5 label-neighbourhood minimum of the step (:61-76) never crosses
// lanes: the y-window runs along the packed axis of u16x2 registers (one PRMT + two 3-input minima per register), the x-window
// across registers of neighbouring label columns (two 3-input minima per register).  About 40 warp instructions per pixel and
// direction against ~180 for the warp-per-pixel kernels (pydv.cu, pyd.cu), whose x-window goes through shared memory.
//
// Volumes (cost, per-direction L) use the layout [pair][y][label column sx][x][16 bytes]: a label column is a 16-byte frame
// (two pad bytes, Sy labels, pad bytes up to 16, every pad 255), and the frames of one label column of 32 neighbouring pixels
// are 512 contiguous bytes — a warp's 128-bit loads and stores are fully coalesced in the six non-horizontal directions.
//
// Prior shift (:46-47, :213-254): the predecessor label of (sx, sy) is ((int)(sx + ddx + 0.5), (int)(sy + ddy + 0.5)) with
// (ddx, ddy) the difference of the prior flow between the pixel and its predecessor on the path.  Truncation toward zero makes
// that map, per axis, a shift by k with the labels whose argument is negative shifted by k + 1:  a(s) = s + k + [s < t].  A
// pre-pass (pydl_desc_kernel) evaluates the reference's expression for every label row / column in fp64, verifies that form and
// stores (kx, ky, tx, ty) per pixel and direction; a step whose map is not of that form (rounding anomalies of fractional priors)
// is flagged and redone label by label (pl_generic_step).  The previous row's path costs live in shared memory as
// [word][lane] (bank = lane: conflict-free for any per-lane shift); a shifted column is five 32-bit loads + four funnel shifts
// + a validity mask, columns outside the window read an all-255 column.  Pads and 255s never win: in the no-wrap parameter
// domain the far term M + P2 is at most 255 - P1.
//
// Applies to: Sx, Sy <= 11, parameters in the no-wrap domain (P1, P2 >= 0, 25 + P1 + P2 <= 255, 50 + P2 <= 255), every enabled
// direction counted once (totalPass 1 or 2).  Everything else takes the one-warp-per-scanline kernels of pyd.cu.
#include "fsgm_internal.h"

namespace fsgm {

constexpr int PL_WARPS = 2;                 // warps are independent; the block size only sets the shared-memory granularity (8 warps per SM)
// (ncu r2I, 32 pairs at level 0: 8 warps resident per SM — at 188 registers a sub-partition's 16 K registers hold two warps — and
// the stall samples sit on shared memory: short scoreboard 18 %, MIO queue 15 %, i.e. on the ~140 LDS.32 / STS.32 / LDGSTS a step
// issues, which the [word][lane] state layout keeps 32 bits wide.  More resident warps do not help: register caps of 157-164 with
// the maximum shared-memory carve-out, 10-11 warps per SM, gave 783-796 pairs/s against 805-807 uncapped.  Neither does halving the
// instruction count: a [word pair][lane][2] layout with 64-bit accesses (3 LDS.64 + a clamp-mode funnel shift for odd word offsets
// instead of 5 LDS.32, 2 STS.64 instead of 4 STS.32; equally conflict free) passed every parity test and ran at 806 against 811.)
#ifndef FSGM_PL_DUP_ALWAYS
#define FSGM_PL_DUP_ALWAYS 0
#endif
constexpr bool PL_DUP_ALWAYS = FSGM_PL_DUP_ALWAYS != 0;   // 1: the truncation-duplicate selects run unconditionally (straight-line step)

struct PlParams {
    const uint8_t* C;
    uint8_t* L[8];
    const uint32_t* desc[8];
    const uint8_t* I1;
    const double* preMv;
    int mvW, mvH;
    int dir[8];
    int chunk_start[9];
    int n_dirs;
    int W, H, Sy, P1, P2, adaptive, n_pairs;
};

__device__ __forceinline__ uint32_t pl_lds32(uint32_t a) { uint32_t v; asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(a)); return v; }
__device__ __forceinline__ uint32_t pl_lds8(uint32_t a) { uint32_t v; asm volatile("ld.shared.u8 %0, [%1];" : "=r"(v) : "r"(a)); return v; }
__device__ __forceinline__ void pl_sts32(uint32_t a, uint32_t v) { asm volatile("st.shared.u32 [%0], %1;" ::"r"(a), "r"(v) : "memory"); }
__device__ __forceinline__ void pl_sts8(uint32_t a, uint32_t v) { asm volatile("st.shared.u8 [%0], %1;" ::"r"(a), "r"(v) : "memory"); }
// volatile: the compiler keeps these where they are written (requests issued early, far from their first use)
__device__ __forceinline__ uint4 pl_ldg128(const void* p)
{
    uint4 v; asm volatile("ld.global.nc.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p)); return v;
}
__device__ __forceinline__ uint32_t pl_ldg32(const void* p) { uint32_t v; asm volatile("ld.global.nc.u32 %0, [%1];" : "=r"(v) : "l"(p)); return v; }
__device__ __forceinline__ uint4 pl_lds128(uint32_t a)
{
    uint4 v; asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(a)); return v;
}
// asynchronous global -> shared copies (LDGSTS): no register, no scoreboard; one wait per step
__device__ __forceinline__ void pl_cp16(uint32_t dst, const void* src) { asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src) : "memory"); }
__device__ __forceinline__ void pl_cp4(uint32_t dst, const void* src) { asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(dst), "l"(src) : "memory"); }
__device__ __forceinline__ void pl_cp_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
__device__ __forceinline__ void pl_cp_wait() { asm volatile("cp.async.wait_group 0;" ::: "memory"); }
__device__ __forceinline__ uint32_t pl_min3(uint32_t a, uint32_t b, uint32_t c) { return __vimin3_u16x2(a, b, c); }

__device__ __forceinline__ int pl_x86_d2i(double v)
{
    if (!(v > -2147483649.0 && v < 2147483648.0)) return INT_MIN;
    return __double2int_rz(v);
}

// One step label by label, straight from the reference's expressions (:40-88); used for the flagged steps only.
// src_s / dst_s: shared-space byte addresses of word 0 of the lane's previous / new state ([word][lane] words of 4 bytes).
__device__ __noinline__ uint32_t pl_generic_step(uint32_t src_s, uint32_t dst_s, const uint8_t* Cpix, uint8_t* Lpix, size_t colstride,
                                                 int SX, int Sy, double ddx, double ddy, uint32_t M, int P1, int P2, bool active)
{
    const uint32_t far_ = M + (uint32_t)P2;
    uint32_t mn = 0xFFFFu;
    auto S = [&](int cx, int cy) { return pl_lds8(src_s + (uint32_t)((cx * 4 + ((cy + 2) >> 2)) * 128 + ((cy + 2) & 3))); };
    for (int sx = 0; sx < SX; ++sx) {
        const int xp = min(max(pl_x86_d2i(__dadd_rn(__dadd_rn((double)sx, ddx), 0.5)), -8), 64);
        for (int sy = 0; sy < Sy; ++sy) {
            const int yp = min(max(pl_x86_d2i(__dadd_rn(__dadd_rn((double)sy, ddy), 0.5)), -8), 64);
            uint32_t best = far_;
            if ((unsigned)xp < (unsigned)SX && (unsigned)yp < (unsigned)Sy) best = min(best, S(xp, yp));
            for (int m = -2; m <= 2; ++m) {
                const int tx = xp + m;
                if ((unsigned)tx >= (unsigned)SX) continue;
                for (int k = -2; k <= 2; ++k) {
                    const int ty = yp + k;
                    if ((unsigned)ty >= (unsigned)Sy || (m == 0 && k == 0)) continue;
                    best = min(best, S(tx, ty) + (uint32_t)P1);
                }
            }
            const uint32_t l = (uint32_t)Cpix[(size_t)sx * colstride + 2 + sy] + best - M;
            pl_sts8(dst_s + (uint32_t)((sx * 4 + ((sy + 2) >> 2)) * 128 + ((sy + 2) & 3)), l);
            if (active) Lpix[(size_t)sx * colstride + 2 + sy] = (uint8_t)l;
            mn = min(mn, l);
        }
    }
    return mn;
}

template <int SX>
__global__ void __launch_bounds__(PL_WARPS * 32, 4)
pydl_sweep_kernel(const PlParams prm)
{
    constexpr int STW = (SX + 1) * 4;                       // words of one state buffer: SX label columns + one all-255 column
    constexpr int WW = 4 + STW + STW + 8;                   // slack | buffer 0 | buffer 1 | slack: a shifted read overruns its column by up to 4 words
                                                            // in front and 8 behind; between the buffers it lands in the other buffer (masked bytes)
    extern __shared__ __align__(128) uint32_t pl_smem[];
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    // blocks are numbered scanline-chunk-major, pair-minor: the long horizontal sweeps of EVERY pair are scheduled first
    const int pair = blockIdx.x % prm.n_pairs;
    const int gw = (blockIdx.x / prm.n_pairs) * PL_WARPS + wib;
    if (gw >= prm.chunk_start[prm.n_dirs]) return;          // warps are independent: no block-level synchronisation below
    int k = 0;
    while (gw >= prm.chunk_start[k + 1]) ++k;
    const int r = prm.dir[k], dx = dir_dx(r), dy = dir_dy(r);
    const int W = prm.W, H = prm.H, Sy = prm.Sy;
    const int lines = dy == 0 ? H : W, len = dy == 0 ? W : H;
    const int line_raw = (gw - prm.chunk_start[k]) * 32 + lane;
    const bool active = line_raw < lines;                   // lanes past the last scanline repeat it and store nothing
    const int line = min(line_raw, lines - 1);
    const size_t N = (size_t)W * H;
    const uint8_t* __restrict__ Cb = prm.C + (size_t)pair * N * (SX * 16);
    uint8_t* __restrict__ Lb = prm.L[k] + (size_t)pair * N * (SX * 16);
    const uint32_t* __restrict__ desc = prm.desc[k] + (size_t)pair * N;
    const uint8_t* __restrict__ Ib = prm.I1 + (size_t)pair * N;
    const size_t colstride = (size_t)W * 16;

    // per warp: WW state words x 32 lanes | cost-column stage [SX][32 lanes][16 B] | descriptor stage [2][32 lanes]
    constexpr int PW = WW * 32 + SX * 128 + 64;             // words per warp
    const uint32_t warp_s = (uint32_t)__cvta_generic_to_shared(pl_smem + (size_t)wib * PW);
    const uint32_t st_s = warp_s + lane * 4;
    for (int w = 0; w < WW; ++w) pl_sts32(st_s + w * 128, 0xFFFFFFFFu);         // lane-private words: no synchronisation needed
    const uint32_t cst_s = warp_s + WW * 128 + lane * 16, dst_stage_s = warp_s + WW * 128 + SX * 512 + lane * 4;

    // pad bytes of a frame (rows outside [0, Sy)) and the per-register row masks
    uint32_t padw[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        uint32_t m = 0;
#pragma unroll
        for (int b = 0; b < 4; ++b) { const int row = 4 * j + b - 2; if (row < 0 || row >= Sy) m |= 0xFFu << (8 * b); }
        padw[j] = m;
    }
    const uint32_t P1P1 = (uint32_t)prm.P1 * 0x10001u;

    int x, y;
    if (dy == 0) { y = line; x = dx > 0 ? 0 : W - 1; }
    else         { x = line; y = dy > 0 ? 0 : H - 1; }
    uint32_t M = 0;
    int cur = 0;
    int iprev = 0;
    uint32_t ppix = 0;
    // A step's cost columns and descriptor are copied into the stage asynchronously (LDGSTS) during the step before it: the slot
    // of cost column c is requested again right after the step has read it, one commit group per column, so that at every read
    // exactly SX - 1 younger groups may still be in flight (cp.async.wait_group SX-1).  The descriptor rides with column SX-1.
    {
        const uint8_t* cn = Cb + ((size_t)((uint32_t)y * (uint32_t)SX) * W + x) * 16;
#pragma unroll
        for (int cc = SX - 1; cc >= 0; --cc) { pl_cp16(cst_s + (uint32_t)(cc * 512), cn + (size_t)cc * colstride); pl_cp_commit(); }
    }

    for (int t = 0; t < len; ++t) {
        const uint32_t pix = (uint32_t)y * (uint32_t)W + (uint32_t)x;
        const bool start = (t == 0) || (dy != 0 && dx != 0 && x == (dx > 0 ? 0 : W - 1));
        // the descriptor of this step arrived with the first cost column of the previous step's requests
        asm volatile("cp.async.wait_group %0;" ::"n"(SX - 1) : "memory");
        const uint32_t d = start ? 0u : pl_lds32(dst_stage_s + (uint32_t)((t & 1) * 128));
        int nx = x + dx, ny = y + dy;
        if (dy != 0) nx = nx < 0 ? W - 1 : (nx >= W ? 0 : nx);
        const bool more = t + 1 < len;
        const uint8_t* cnx = Cb + ((size_t)((uint32_t)(more ? ny : y) * (uint32_t)SX) * W + (more ? nx : x)) * 16;
        const uint32_t* dnx = desc + (uint32_t)(more ? ny : y) * (uint32_t)W + (uint32_t)(more ? nx : x);
        int P2 = prm.P2;
        if (prm.adaptive) {
            const int icur = Ib[pix];
            if (!start && abs(icur - iprev) > 50) P2 = P2 / 8;
            iprev = icur;
        }
        // path start (:152-180): L = C and the stored minimum is 0 — far term 0 makes every label's best 0
        const uint32_t MM = start ? 0u : M * 0x10001u;
        const uint32_t far2 = start ? 0u : (M + (uint32_t)P2) * 0x10001u;
        const int kx = (int)(int8_t)(d & 0xFFu), ky = (int)(int8_t)((d >> 8) & 0xFFu);
        const int tx = (int)((d >> 16) & 15u), ty = (int)((d >> 20) & 15u);
        const bool gen = (d >> 24) & 1u;
        const int kyc = min(max(ky, -13), 13);
        const int q = kyc >> 2;
        const uint32_t rr8 = (uint32_t)(kyc & 3) * 8u;
        // frame byte b of a shifted column holds source row b - 2 + ky: valid iff that row is a label row
        uint32_t mk[4];
        {
            const int lo = 2 - ky, hi = lo + Sy;
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const int nlow = __vimin_s32_relu(lo - 4 * j, 4), nhigh = __vimin_s32_relu(4 * j + 4 - hi, 4);
                mk[j] = __funnelshift_lc(0xFFFFFFFFu, 0u, 8 * nlow) | __funnelshift_rc(0u, 0xFFFFFFFFu, 8 * nhigh);
            }
        }
        const bool anydup = __any_sync(0xffffffffu, (tx | ty) != 0);
        uint32_t dm[5];
#pragma unroll
        for (int i = 0; i < 5; ++i) dm[i] = (2 * i < ty ? 0xFFFFu : 0u) | (2 * i + 1 < ty ? 0xFFFF0000u : 0u);

        const uint32_t src0_s = st_s + (uint32_t)((4 + (cur ^ 1) * STW) * 128), dst_s = st_s + (uint32_t)((4 + cur * STW) * 128);
        const uint32_t src_s = src0_s + (uint32_t)(q * 128);
        const uint32_t ccol_s = cst_s;
        const uint8_t* cpix = Cb + ((size_t)((uint32_t)y * (uint32_t)SX) * W + x) * 16;
        uint8_t* lpix = Lb + ((size_t)((uint32_t)y * (uint32_t)SX) * W + x) * 16;

        uint32_t Y[5][6], EC[3][6], bprev[6];
        uint32_t mm = 0xFFFFFFFFu;
#pragma unroll
        for (int i = 0; i < 6; ++i) bprev[i] = 0;
        // the five words of a source column are requested two iterations before they are used (and so, in program order, before
        // the stores of the columns finished in between: the hardware keeps shared-memory accesses of a thread in order)
        uint32_t wq[3][5];
        uint4 cnext = make_uint4(0, 0, 0, 0), ccur = cnext;
        auto fetch = [&](int it_) {
            const int cp_ = SX + 1 - it_;
            const uint32_t ce = min((uint32_t)(cp_ + kx), (uint32_t)SX);
            const uint32_t a = src_s + ce * 512u;
#pragma unroll
            for (int j = 0; j < 5; ++j) wq[it_ % 3][j] = pl_lds32(a + 128 * j);
        };
        fetch(0); fetch(1);
#pragma unroll
        for (int it = 0; it < SX + 4; ++it) {
            const int cp = SX + 1 - it;                                 // shifted-grid column whose y-window minima are built now
            if (it + 2 < SX + 4) fetch(it + 2);
            // ---- source column cp + kx (anything outside the window: the all-255 column), shifted by ky rows --------------------
            {
                const uint32_t w0 = wq[it % 3][0], w1 = wq[it % 3][1], w2 = wq[it % 3][2], w3 = wq[it % 3][3], w4 = wq[it % 3][4];
                const uint32_t b0 = __funnelshift_r(w0, w1, rr8) | mk[0], b1 = __funnelshift_r(w1, w2, rr8) | mk[1],
                               b2 = __funnelshift_r(w2, w3, rr8) | mk[2], b3 = __funnelshift_r(w3, w4, rr8) | mk[3];
                uint32_t E[8], O[7];                                    // E[i]: rows (2i-2, 2i-1); O[i]: rows (2i-1, 2i)
                E[0] = __byte_perm(b0, 0, 0x4140); E[1] = __byte_perm(b0, 0, 0x4342);
                E[2] = __byte_perm(b1, 0, 0x4140); E[3] = __byte_perm(b1, 0, 0x4342);
                E[4] = __byte_perm(b2, 0, 0x4140); E[5] = __byte_perm(b2, 0, 0x4342);
                E[6] = __byte_perm(b3, 0, 0x4140); E[7] = __byte_perm(b3, 0, 0x4342);
#pragma unroll
                for (int i = 0; i < 7; ++i) O[i] = __byte_perm(E[i], E[i + 1], 0x5432);
                // rows (2i, 2i+1): windows [2i-2, 2i+2] and [2i-1, 2i+3]
#pragma unroll
                for (int i = 0; i < 6; ++i) {
                    Y[(cp + 2) % 5][i] = pl_min3(pl_min3(E[i], O[i], E[i + 1]), O[i + 1], E[i + 2]);
                    EC[(cp + 2) % 3][i] = E[i + 1];
                }
            }
            const int c = cp + 2;                                       // output column of this iteration
            // cost column c - 1 is read one iteration ahead; its slot is then requested for the next step (the last step re-requests
            // its own columns: harmless, and the group count stays uniform)
            if (c >= 1 && c <= SX) {
                if (c < SX) asm volatile("cp.async.wait_group %0;" ::"n"(SX - 1) : "memory");
                ccur = cnext;
                cnext = pl_lds128(ccol_s + (uint32_t)((c - 1) * 512));
                pl_cp16(ccol_s + (uint32_t)((c - 1) * 512), cnx + (size_t)(c - 1) * colstride);
                if (c == SX) pl_cp4(dst_stage_s + (uint32_t)(((t + 1) & 1) * 128), dnx);
                pl_cp_commit();
            } else if (c == 0) ccur = cnext;
            if (c >= 0 && c < SX) {
                const uint4 ccol = ccur;
                uint32_t bst[6];
#pragma unroll
                for (int i = 0; i < 6; ++i) {
                    const uint32_t xm = pl_min3(pl_min3(Y[0][i], Y[1][i], Y[2][i]), Y[3][i], Y[4][i]);
                    bst[i] = pl_min3(far2, xm + P1P1, EC[(cp + 4) % 3][i]);
                }
                if (PL_DUP_ALWAYS || anydup) {
                    // rows below ty take the next row's value, columns below tx the next column's (truncation toward zero)
#pragma unroll
                    for (int i = 0; i < 5; ++i) {
                        const uint32_t up = __byte_perm(bst[i], bst[i + 1], 0x5432);
                        bst[i] = (bst[i] & ~dm[i]) | (up & dm[i]);
                    }
                    const bool dupx = c < tx;
#pragma unroll
                    for (int i = 0; i < 6; ++i) { const uint32_t own = bst[i]; bst[i] = dupx ? bprev[i] : own; bprev[i] = own; }
                }
                uint32_t cu[6], l[6];
                cu[0] = __byte_perm(ccol.x, 0, 0x4342); cu[1] = __byte_perm(ccol.y, 0, 0x4140); cu[2] = __byte_perm(ccol.y, 0, 0x4342);
                cu[3] = __byte_perm(ccol.z, 0, 0x4140); cu[4] = __byte_perm(ccol.z, 0, 0x4342); cu[5] = __byte_perm(ccol.w, 0, 0x4140);
#pragma unroll
                for (int i = 0; i < 6; ++i) l[i] = cu[i] + bst[i] - MM;          // every candidate >= M: no borrow between the halves
                mm = pl_min3(mm, l[0], l[1]); mm = pl_min3(mm, l[2], l[3]); mm = pl_min3(mm, l[4], l[5]);
                uint4 o;
                o.x = __byte_perm(l[0], 0xFFFFFFFFu, 0x2044);
                o.y = __byte_perm(l[1], l[2], 0x6420) | padw[1];
                o.z = __byte_perm(l[3], l[4], 0x6420) | padw[2];
                o.w = __byte_perm(l[5], 0xFFFFFFFFu, 0x4420) | padw[3];
                o.x |= padw[0];
                const uint32_t da = dst_s + (uint32_t)c * 512u;
                pl_sts32(da, o.x); pl_sts32(da + 128, o.y); pl_sts32(da + 256, o.z); pl_sts32(da + 384, o.w);
                if (active) *reinterpret_cast<uint4*>(lpix + (size_t)c * colstride) = o;
            }
        }
        uint32_t m = min(mm & 0xFFFFu, mm >> 16);
        if (__any_sync(0xffffffffu, gen)) {
            if (gen) {
                const double* mvx = prm.preMv + (size_t)pair * 2 * prm.mvW * prm.mvH;
                const double* mvy = mvx + (size_t)prm.mvW * prm.mvH;
                const uint32_t py = ppix / (uint32_t)W, px = ppix - py * (uint32_t)W;
                const double ddx = __dsub_rn(mvx[(size_t)y * prm.mvW + x], mvx[(size_t)py * prm.mvW + px]);
                const double ddy = __dsub_rn(mvy[(size_t)y * prm.mvW + x], mvy[(size_t)py * prm.mvW + px]);
                m = pl_generic_step(src0_s, dst_s, cpix, lpix, colstride, SX, Sy, ddx, ddy, M, prm.P1, P2, active);
            }
        }
        M = start ? 0u : m;
        cur ^= 1;
        ppix = pix;
        x = nx; y = ny;
    }
    pl_cp_wait();
}

// ---- shift descriptors ---------------------------------------------------------------------------------------------------------
struct PlDescParams { uint32_t* out[8]; int dir[8]; int n_dirs; };

// a(s) = (int)((s + dd) + 0.5) for s = 0..S-1 as the reference evaluates it; what matters is the column it selects, clamped to
// [-3, S+2] (three or more outside the window: no neighbour inside).  Returns true when a(s) = s + k + [s < t].
__device__ __forceinline__ bool pl_analyse(double dd, int S, int* k_out, int* t_out)
{
    // integer-valued difference n (what the pyramid driver produces: priors are 2 x integer): the argument s + n + 0.5 truncates
    // to s + n where it is positive and to s + n + 1 where it is negative
    if (dd == rint(dd) && fabs(dd) <= 64.0) {
        const int n = (int)dd;
        int k = n, t = min(max(-n, 0), S);
        if (t == S) { k = n + 1; t = 0; }
        *k_out = min(max(k, -15), 15); *t_out = t;
        return true;
    }
    int a[11];
#pragma unroll
    for (int s = 0; s < 11; ++s)
        a[s] = s < S ? min(max(pl_x86_d2i(__dadd_rn(__dadd_rn((double)s, dd), 0.5)), -40), 60) : 0;
    int last = 0;
#pragma unroll
    for (int s = 0; s < 11; ++s) if (s == S - 1) last = a[s];
    const int k = min(max(last - (S - 1), -15), 15);
    int t = 0;                                                 // one past the last label that does not follow the plain shift
#pragma unroll
    for (int s = 0; s < 11; ++s)
        if (s < S && min(max(s + k, -3), S + 2) != min(max(a[s], -3), S + 2)) t = s + 1;
    bool ok = true;
#pragma unroll
    for (int s = 0; s < 11; ++s)
        if (s < S) ok &= min(max(s + k + (s < t ? 1 : 0), -3), S + 2) == min(max(a[s], -3), S + 2);
    *k_out = k; *t_out = t;
    return ok;
}

__global__ void pydl_desc_kernel(const double* __restrict__ preMv, int mvW, int mvH, int W, int H, int Sx, int Sy, int force_generic,
                                 const PlDescParams prm)
{
    const size_t N = (size_t)W * H;
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= N) return;
    const int y = (int)(i / W), x = (int)(i - (size_t)y * W);
    const double* mvx = preMv + (size_t)blockIdx.y * 2 * mvW * mvH;
    const double* mvy = mvx + (size_t)mvW * mvH;
    const double cx = mvx[(size_t)y * mvW + x], cy = mvy[(size_t)y * mvW + x];
    for (int k = 0; k < prm.n_dirs; ++k) {
        const int r = prm.dir[k];
        const int px = x - dir_dx(r), py = y - dir_dy(r);
        uint32_t d = 0;
        if (px >= 0 && px < W && py >= 0 && py < H) {
            const double ddx = __dsub_rn(cx, mvx[(size_t)py * mvW + px]), ddy = __dsub_rn(cy, mvy[(size_t)py * mvW + px]);
            if (!(ddx == 0.0 && ddy == 0.0) || force_generic) {
                int kx, tx, ky, ty;
                const bool okx = pl_analyse(ddx, Sx, &kx, &tx), oky = pl_analyse(ddy, Sy, &ky, &ty);
                d = ((uint32_t)kx & 0xFFu) | (((uint32_t)ky & 0xFFu) << 8) | ((uint32_t)tx << 16) | ((uint32_t)ty << 20);
                if (!(okx && oky) || force_generic) d |= 1u << 24;
            }
        }
        prm.out[k][blockIdx.y * N + i] = d;
    }
}

// ---- winner-take-all + per-axis parabola (:298-360), lane = pixel ----------------------------------------------------------
struct PlWtaParams {
    const uint8_t* L[8]; int n_dirs;
    int W, H, Sy, subpixel;
    uint32_t* bestD; uint32_t* minC; double* mvSub;
};

template <int SX>
__global__ void __launch_bounds__(128)
pydl_wta_kernel(const PlWtaParams prm)
{
    const int W = prm.W, Sy = prm.Sy, R = prm.n_dirs;
    const size_t N = (size_t)W * prm.H;
    const size_t p = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= N) return;
    const int y = (int)(p / W), x = (int)(p - (size_t)y * W);
    const size_t colstride = (size_t)W * 16;
    const size_t base = (size_t)blockIdx.y * N * (SX * 16) + ((size_t)y * SX * W + x) * 16;
    uint32_t padh[6];
#pragma unroll
    for (int i = 0; i < 6; ++i) padh[i] = (2 * i < Sy ? 0u : 0xFFFFu) | (2 * i + 1 < Sy ? 0u : 0xFFFF0000u);
    // first minimum in label order d = sx * Sy + sy (:298-314).  Sums are below 8 * 255 < 4096, so (sum << 4 | row) fits a
    // 16-bit half and its minimum is the column's smallest sum at its first row; columns are then compared as
    // (sum << 8 | column << 4 | row).  One pass, nothing but the running key is kept.
    uint32_t key = 0xFFFFFFFFu;
#pragma unroll
    for (int c = 0; c < SX; ++c) {
        uint4 v[8];
#pragma unroll
        for (int k = 0; k < 8; ++k) v[k] = k < R ? __ldg(reinterpret_cast<const uint4*>(prm.L[k] + base + (size_t)c * colstride)) : make_uint4(0, 0, 0, 0);
        uint32_t s[6];
#pragma unroll
        for (int i = 0; i < 6; ++i) s[i] = 0;
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            if (k < R) {
                s[0] += __byte_perm(v[k].x, 0, 0x4342); s[1] += __byte_perm(v[k].y, 0, 0x4140); s[2] += __byte_perm(v[k].y, 0, 0x4342);
                s[3] += __byte_perm(v[k].z, 0, 0x4140); s[4] += __byte_perm(v[k].z, 0, 0x4342); s[5] += __byte_perm(v[k].w, 0, 0x4140);
            }
        }
        uint32_t kc = 0xFFFFFFFFu;
#pragma unroll
        for (int i = 0; i < 6; ++i) kc = __vminu2(kc, (s[i] * 16u + (uint32_t)((2 * i) | ((2 * i + 1) << 16))) | padh[i]);
        const uint32_t kr = min(kc & 0xFFFFu, kc >> 16);
        key = min(key, ((kr >> 4) << 8) | ((uint32_t)c << 4) | (kr & 15u));
    }
    const uint32_t best = key >> 8;
    const uint32_t idx = ((key >> 4) & 15u) * (uint32_t)Sy + (key & 15u);
    const size_t gp = blockIdx.y * N + p;
    prm.bestD[gp] = idx;
    prm.minC[gp] = best;
    double sx = 0.0, sy = 0.0;
    if (prm.subpixel) {
        const int lx = (int)idx / Sy, ly = (int)idx - lx * Sy;
        auto S = [&](int cx, int cy) {
            uint32_t a = 0;
            for (int k = 0; k < R; ++k) a += prm.L[k][base + (size_t)cx * colstride + 2 + cy];
            return (double)a;
        };
        const double c0 = (double)best;
        if (ly > 0 && ly < Sy - 1) {
            const double a = S(lx, ly - 1), b = S(lx, ly + 1);
            sy = (b < a) ? __ddiv_rn(__ddiv_rn(__dsub_rn(b, a), __dsub_rn(c0, a)), 2.0)
                         : __ddiv_rn(__ddiv_rn(__dsub_rn(b, a), __dsub_rn(c0, b)), 2.0);
        }
        if (lx > 0 && lx < SX - 1) {
            const double a = S(lx - 1, ly), b = S(lx + 1, ly);
            sx = (b < a) ? __ddiv_rn(__ddiv_rn(__dsub_rn(b, a), __dsub_rn(c0, a)), 2.0)
                         : __ddiv_rn(__ddiv_rn(__dsub_rn(b, a), __dsub_rn(c0, b)), 2.0);
        }
    }
    prm.mvSub[blockIdx.y * 2 * N + p] = sx;
    prm.mvSub[blockIdx.y * 2 * N + N + p] = sy;
}

// ---- host side -----------------------------------------------------------------------------------------------------------------
bool pydl_applicable(int Sx, int Sy, int P1, int P2, int n_dirs, const int* weights)
{
    if (Sx < 1 || Sy < 1 || Sx > 11 || Sy > 11 || !(Sx & 1) || n_dirs < 1) return false;
    for (int k = 0; k < n_dirs; ++k) if (weights[k] != 1) return false;
    return P1 >= 0 && P2 >= 0 && 25 + P1 + P2 <= 255 && 50 + P2 <= 255;
}

int launch_pydl_desc(fsgm_ctx* c, int n, const double* preMv, int mvW, int mvH, int W, int H, int Sx, int Sy, const int* dirs, int n_dirs,
                     uint32_t* const* out, int force_generic)
{
    StageScope ss(c, ST_PYD_SWEEP);
    PlDescParams p{};
    for (int k = 0; k < n_dirs; ++k) { p.out[k] = out[k]; p.dir[k] = dirs[k]; }
    p.n_dirs = n_dirs;
    const size_t N = (size_t)W * H;
    pydl_desc_kernel<<<dim3((unsigned)((N + 127) / 128), n), 128, 0, c->stream>>>(preMv, mvW, mvH, W, H, Sx, Sy, force_generic, p);
    FSGM_LAUNCHED(c);
    return FSGM_OK;
}

template <int SX>
static int pl_launch_sweep(fsgm_ctx* c, const PlParams& p, dim3 grid)
{
    constexpr size_t smem = (size_t)PL_WARPS * ((4 + 2 * ((SX + 1) * 4) + 8) * 32 + SX * 128 + 64) * 4;
    auto kern = pydl_sweep_kernel<SX>;
    FSGM_CUDA(c, cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    kern<<<grid, PL_WARPS * 32, smem, c->stream>>>(p);
    return FSGM_OK;
}

int launch_pydl_sweeps(fsgm_ctx* c, int n, const uint8_t* C, const uint8_t* I1, const double* preMv, int mvW, int mvH, int W, int H,
                       int Sx, int Sy, int P1, int P2, int adaptive, const int* dirs, int n_dirs, uint32_t* const* desc, uint8_t* const* Lvols)
{
    StageScope ss(c, ST_PYD_SWEEP);
    PlParams p{};
    p.C = C; p.I1 = I1; p.preMv = preMv; p.mvW = mvW; p.mvH = mvH; p.W = W; p.H = H; p.Sy = Sy; p.P1 = P1; p.P2 = P2; p.adaptive = adaptive; p.n_pairs = n;
    // the two horizontal directions run W steps against H for the others: their warps go first
    int order[8], m = 0;
    for (int k = 0; k < n_dirs; ++k) if (dir_dy(dirs[k]) == 0) order[m++] = k;
    for (int k = 0; k < n_dirs; ++k) if (dir_dy(dirs[k]) != 0) order[m++] = k;
    p.n_dirs = n_dirs;
    p.chunk_start[0] = 0;
    for (int j = 0; j < n_dirs; ++j) {
        const int k = order[j];
        p.dir[j] = dirs[k]; p.L[j] = Lvols[k]; p.desc[j] = desc[k];
        p.chunk_start[j + 1] = p.chunk_start[j] + ((dir_dy(dirs[k]) == 0 ? H : W) + 31) / 32;
    }
    dim3 grid((unsigned)(((p.chunk_start[n_dirs] + PL_WARPS - 1) / PL_WARPS) * n));
    switch (Sx) {
        case 1: FSGM_TRY(pl_launch_sweep<1>(c, p, grid)); break;
        case 3: FSGM_TRY(pl_launch_sweep<3>(c, p, grid)); break;
        case 5: FSGM_TRY(pl_launch_sweep<5>(c, p, grid)); break;
        case 7: FSGM_TRY(pl_launch_sweep<7>(c, p, grid)); break;
        case 9: FSGM_TRY(pl_launch_sweep<9>(c, p, grid)); break;
        case 11: FSGM_TRY(pl_launch_sweep<11>(c, p, grid)); break;
        default: return fail(c, FSGM_ERR_DOMAIN, "lane = path pyd sweep: window width must be odd and <= 11");
    }
    FSGM_LAUNCHED(c);
    return FSGM_OK;
}

int launch_pydl_wta(fsgm_ctx* c, int n, uint8_t* const* Lvols, int n_dirs, int W, int H, int Sx, int Sy, int subpixel,
                    uint32_t* bestD, uint32_t* minC, double* mvSub)
{
    StageScope ss(c, ST_PYD_WTA);
    PlWtaParams p{};
    for (int k = 0; k < n_dirs; ++k) p.L[k] = Lvols[k];
    p.n_dirs = n_dirs; p.W = W; p.H = H; p.Sy = Sy; p.subpixel = subpixel; p.bestD = bestD; p.minC = minC; p.mvSub = mvSub;
    const size_t N = (size_t)W * H;
    dim3 grid((unsigned)((N + 127) / 128), n);
    switch (Sx) {
        case 1: pydl_wta_kernel<1><<<grid, 128, 0, c->stream>>>(p); break;
        case 3: pydl_wta_kernel<3><<<grid, 128, 0, c->stream>>>(p); break;
        case 5: pydl_wta_kernel<5><<<grid, 128, 0, c->stream>>>(p); break;
        case 7: pydl_wta_kernel<7><<<grid, 128, 0, c->stream>>>(p); break;
        case 9: pydl_wta_kernel<9><<<grid, 128, 0, c->stream>>>(p); break;
        case 11: pydl_wta_kernel<11><<<grid, 128, 0, c->stream>>>(p); break;
        default: return fail(c, FSGM_ERR_DOMAIN, "lane = path pyd WTA: window width must be odd and <= 11");
    }
    FSGM_LAUNCHED(c);
    return FSGM_OK;
}

}  // namespace fsgm
