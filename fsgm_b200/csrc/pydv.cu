// Row-synchronous aggregation for the pyramidal 2-D-window variant — the fast path of sgm2d() (reference
// calc_pyd_cost_sgm.cpp:114-372, step :34-89) — built like the epipolar cluster kernel (vsweep.cu):
//
//   * the three directions that advance one image row per step ((0,+1) (+1,+1) (-1,+1) going down, their negatives going up)
//     need the same cost row at the same time: ONE thread-block cluster walks a pair row by row, every CTA owns a strip of
//     columns and keeps the three directions' previous-row path costs in shared memory, a diagonal path that leaves the strip
//     is handed to the neighbour CTA through distributed shared memory, the strip's cost row arrives by one 1-D bulk copy
//     (cp.async.bulk + mbarrier) two rows ahead;
//   * the down pass writes ONE byte per voxel, the sum of its three directions (3*(25+P2) <= 255); the up pass reads that
//     volume and the two horizontal directions' volumes, adds its own three directions and does winner-take-all in the
//     kernel: neither the six L volumes nor Sp reach HBM.  The per-axis parabola (:333-360) runs in a finalize kernel on
//     16-byte records (argmin, minimum, four neighbours).
//
// Label layout: the (2rx+1) x (2ry+1) window is a PADDED GRID of 16-byte columns (column sx at byte 16*sx: two pad bytes, Sy
// labels, pad bytes up to 16), every pad 255, in HBM (cost volume, byte volumes) and in shared memory alike.  A lane owns half a
// column (8 slots) as four u16x2 registers; the 5x5 label-neighbourhood minimum of the step (:61-76) is separable: the y-window
// comes from the lane's own 8 bytes + 2 on either side, the x-window from a u16 grid in per-warp shared memory (one STS.128 +
// five LDS.128).  Pads never win (255 >= far term), so there are no bounds tests.
//
// Prior shift: the predecessor label of (sx,sy) is ((int)(sx+ddx+0.5), (int)(sy+ddy+0.5)) with (ddx,ddy) the difference of the
// prior flow between the pixel and its predecessor on the path (:213-254).  A pre-pass marks, per pixel and direction, whether
// that difference is zero (always at the coarsest level, and wherever the 2x-upsampled prior is locally constant); only marked
// steps take the general form (coordinate tables, byte gathers from the same padded grids), a warp-uniform branch.
//
// Applies to: Sx, Sy <= 11 (the reference's 11 x 11 and BASELINE's 9 x 9 windows), 8 paths, 2 passes, no adaptive P2, parameters
// inside the no-wrap domain with 3*(25+P2) <= 255.  Everything else takes the one-warp-per-scanline kernels of pyd.cu.
#include "fsgm_internal.h"
#include <cooperative_groups.h>
#include <algorithm>

namespace cg = cooperative_groups;

namespace fsgm {

constexpr int PV_WARPS = 20;
constexpr int PV_RCOLS = 21;                 // columns of the per-warp y-window-minimum grid: Sx + 10 (5 pad columns on either side)
constexpr int PV_RGRID = PV_RCOLS * 16;      // u16 entries
constexpr int PV_WSCR = 3 * PV_RGRID * 2 + 13 * 16 * 2 + 64;     // per-warp scratch bytes: 3 R grids, the sum grid (Sx + 2 columns), tables

struct PvParams {
    const uint8_t* C;            // [n][H][W][PITCH] padded grid
    const uint8_t* H0;           // FINAL: the two horizontal directions' L volumes (padded grid)
    const uint8_t* H1;
    const uint8_t* S1in;         // FINAL: the down pass's byte volume
    uint8_t* S1out;              // !FINAL
    uint4* rec;                  // FINAL: [n][N] WTA records
    const uint8_t* flags;        // [n][N] bit r: direction r takes the general (shifted) step at this pixel
    const double* preMv;         // [n][2][mvH][mvW]
    int mvW, mvH;
    int W, H, Wk, Sx, Sy, P1, P2;
    int up;
};

__device__ __forceinline__ uint32_t pv_smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void pv_mbar_init(uint64_t* bar, int count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(pv_smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void pv_mbar_expect_tx(uint64_t* bar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(pv_smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void pv_mbar_wait(uint64_t* bar, uint32_t parity)
{
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "PVW_%=:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra PVD_%=;\n\t"
        "bra PVW_%=;\n\t"
        "PVD_%=:\n\t}" ::"r"(pv_smem_u32(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ void pv_tma_load_1d(void* dst, const void* src, uint32_t bytes, uint64_t* bar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(pv_smem_u32(dst)), "l"(src), "r"(bytes), "r"(pv_smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void pv_cluster_arrive() { asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory"); }
__device__ __forceinline__ void pv_cluster_arrive_relaxed() { asm volatile("barrier.cluster.arrive.relaxed.aligned;" ::: "memory"); }
__device__ __forceinline__ void pv_cluster_wait() { asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory"); }

__device__ __forceinline__ uint32_t pv_lds32(uint32_t a) { uint32_t v; asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(a)); return v; }
__device__ __forceinline__ uint32_t pv_lds8(uint32_t a) { uint32_t v; asm volatile("ld.shared.u8 %0, [%1];" : "=r"(v) : "r"(a)); return v; }
__device__ __forceinline__ uint2 pv_lds64(uint32_t a) { uint2 v; asm volatile("ld.shared.v2.u32 {%0, %1}, [%2];" : "=r"(v.x), "=r"(v.y) : "r"(a)); return v; }
__device__ __forceinline__ uint4 pv_lds128(uint32_t a)
{
    uint4 v; asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(a)); return v;
}
__device__ __forceinline__ void pv_sts32(uint32_t a, uint32_t v) { asm volatile("st.shared.u32 [%0], %1;" ::"r"(a), "r"(v)); }
__device__ __forceinline__ void pv_sts64(uint32_t a, uint2 v) { asm volatile("st.shared.v2.u32 [%0], {%1, %2};" ::"r"(a), "r"(v.x), "r"(v.y)); }
__device__ __forceinline__ void pv_sts128(uint32_t a, uint4 v)
{
    asm volatile("st.shared.v4.u32 [%0], {%1, %2, %3, %4};" ::"r"(a), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w));
}

__device__ __forceinline__ int pv_x86_d2i(double v)
{
    if (!(v > -2147483649.0 && v < 2147483648.0)) return INT_MIN;
    return __double2int_rz(v);
}
__device__ __forceinline__ uint32_t pv_min3(uint32_t a, uint32_t b, uint32_t c) { return __vimin3_u16x2(a, b, c); }

struct PvThread {
    uint32_t state_s, stmin_s, inbox_s, scr_s;     // shared-space addresses; state_s and the grids carry the lane's byte offset
    uint8_t* inbox_right; uint8_t* inbox_left;     // neighbours' inboxes (distributed shared memory, generic pointers)
    const uint8_t* H0_l; const uint8_t* H1_l; const uint8_t* S1in_l; uint8_t* S1out_l;
    const uint8_t* flags; const double* mvx; const double* mvy;
    uint4* rec;
    int Wk, Wk_max, W, H, xb, lane, vc, vh, Sx, Sy, PITCH, IBS, mvW, up;
    uint32_t P1P1, P2P2, padmask[4];
    uint2 bmask;                                   // 0xFF in the bytes of the lane's real slots
    bool vact;
    int base_d;
};

// One pixel of one row: the three directions, sum, output.  EDGE as in vsweep.cu (first row, first / last column of the strip).
template <bool FINAL, bool EDGE>
__device__ __forceinline__ void pv_pixel(const PvThread& th, uint32_t crow_s, int xl, int yy, int y, int par, int off, uint32_t pix,
                                         uint32_t fl, uint2 gH0, uint2 gH1, uint2 gS1, bool arrive)
{
    const int lane = th.lane, Wk = th.Wk, PITCH = th.PITCH;
    const bool vact = th.vact;
    constexpr uint32_t K255 = 0x00FF00FFu;
    // the pixel's cost: this lane's 8 grid slots
    const uint2 cw = pv_lds64(crow_s + xl * PITCH);
    const uint32_t cc[4] = {__byte_perm(cw.x, 0, 0x4140), __byte_perm(cw.x, 0, 0x4342), __byte_perm(cw.y, 0, 0x4140), __byte_perm(cw.y, 0, 0x4342)};
    // fl: the pixel's shift flags (warp-uniform, fetched one pixel ahead by the caller)
    uint32_t st[3], sm[3], src[3], Mv[3];
    bool restart[3], slow[3];
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        constexpr int dxs[3] = {0, 1, -1};
        const int dx = dxs[k];
        int slot = xl;
        if (dx > 0) { slot = xl - off; if (slot < 0) slot += Wk; }
        if (dx < 0) { slot = xl + off; if (slot >= Wk) slot -= Wk; }
        st[k] = th.state_s + (k * th.Wk_max + slot) * PITCH;
        sm[k] = th.stmin_s + (k * th.Wk_max + slot) * 4;
        restart[k] = false;
        src[k] = st[k];
        bool boxed = false;
        if (EDGE) {
            const int x = th.xb + xl;
            restart[k] = (yy == 0) || (dx > 0 && x == 0) || (dx < 0 && x == th.W - 1);
            const bool from_left = dx > 0 && xl == 0, from_right = dx < 0 && xl == Wk - 1;
            if (!restart[k] && (from_left || from_right)) {
                const uint32_t box = th.inbox_s + (uint32_t)((((yy - 1) & 1) * 2 + (from_right ? 1 : 0)) * th.IBS);
                Mv[k] = pv_lds32(box + 16 + PITCH + 4);
                src[k] = box + 16 + th.vc * 16 + 8 * th.vh;
                boxed = true;
            }
        }
        if (!boxed) Mv[k] = pv_lds32(sm[k]);
        // direction index in the flag byte: down pass k -> r = 1,2,3; up pass k -> r = 5,7,6 (fsgm_internal.h direction table)
        const int r = th.up ? (k == 0 ? 5 : k == 1 ? 7 : 6) : (k + 1);
        slow[k] = !restart[k] && ((fl >> r) & 1u);
    }
    uint32_t l[3][4];
    uint32_t Mn[3];
    // ---- y-window minima of the previous rows -> per-warp R grids (zero-shift form), all three directions before the barrier ----
    uint32_t Ec[3][4];                                               // the lane's own previous values (slots 0..7 as u16x2): the same-label term
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        if ((EDGE && restart[k]) || slow[k]) continue;
        const uint2 w = pv_lds64(src[k]);
        const uint32_t wl = pv_lds32(src[k] - 4), wr = pv_lds32(src[k] + 8);
        // slots -2,-1 of the lane's 8 (the other half's last two, or pads below slot -2 of the column) and slots 8,9
        uint32_t E[6];
        E[0] = th.vh ? __byte_perm(wl, 0, 0x4342) : K255;
        E[1] = __byte_perm(w.x, 0, 0x4140); E[2] = __byte_perm(w.x, 0, 0x4342);
        E[3] = __byte_perm(w.y, 0, 0x4140); E[4] = __byte_perm(w.y, 0, 0x4342);
        E[5] = th.vh ? K255 : __byte_perm(wr, 0, 0x4140);
        uint32_t O[5], rr[4];
#pragma unroll
        for (int i = 0; i < 5; ++i) O[i] = __byte_perm(E[i], E[i + 1], 0x5432);
#pragma unroll
        for (int i = 0; i < 4; ++i) rr[i] = pv_min3(pv_min3(E[i], O[i], E[i + 1]), O[i + 1], E[i + 2]);
        if (vact) pv_sts128(th.scr_s + k * (PV_RGRID * 2) + (5 * 16) * 2, make_uint4(rr[0], rr[1], rr[2], rr[3]));
#pragma unroll
        for (int i = 0; i < 4; ++i) Ec[k][i] = E[i + 1];
    }
    __syncwarp();
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        Mn[k] = 0;
        if (EDGE && restart[k]) {
#pragma unroll
            for (int i = 0; i < 4; ++i) l[k][i] = cc[i];             // L = C, stored minimum 0 (:152-180)
            continue;
        }
        const uint32_t MM = Mv[k] * 0x10001u;
        const uint32_t far2 = MM + th.P2P2;
        uint32_t mm2;
        if (!slow[k]) {
            const uint32_t rq = th.scr_s + k * (PV_RGRID * 2) + (3 * 16) * 2;      // column vc - 2 of the padded R grid
            const uint4 a0 = pv_lds128(rq), a1 = pv_lds128(rq + 32), a2 = pv_lds128(rq + 64), a3 = pv_lds128(rq + 96), a4 = pv_lds128(rq + 128);
            const uint32_t m5[4] = {pv_min3(pv_min3(a0.x, a1.x, a2.x), a3.x, a4.x), pv_min3(pv_min3(a0.y, a1.y, a2.y), a3.y, a4.y),
                                    pv_min3(pv_min3(a0.z, a1.z, a2.z), a3.z, a4.z), pv_min3(pv_min3(a0.w, a1.w, a2.w), a3.w, a4.w)};
            mm2 = 0xFFFFFFFFu;
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const uint32_t best = pv_min3(far2, m5[i] + th.P1P1, Ec[k][i]);        // every candidate >= M: no borrow below
                l[k][i] = __vminu2(cc[i] + best - MM, K255);                          // pad slots (cost 255) saturate back to 255
                mm2 = __vminu2(mm2, l[k][i]);
            }
        } else {
            // ---- general step: predecessor label ((int)(sx+ddx+0.5), (int)(sy+ddy+0.5)), :46-47 / :213-254 --------------------------
            constexpr int dxs[3] = {0, 1, -1};
            const int x = th.xb + xl, pxx = x - dxs[k], pyy = th.up ? y + 1 : y - 1;
            const uint32_t mi = (uint32_t)y * (uint32_t)th.mvW + (uint32_t)x, mp = (uint32_t)pyy * (uint32_t)th.mvW + (uint32_t)pxx;
            const double ddx = __dsub_rn(th.mvx[mi], th.mvx[mp]), ddy = __dsub_rn(th.mvy[mi], th.mvy[mp]);
            // this lane's column -> predecessor column; anything further than 3 outside behaves like 3 outside (no neighbour inside)
            const int xp = min(max(pv_x86_d2i(__dadd_rn(__dadd_rn((double)th.vc, ddx), 0.5)), -3), th.Sx + 2);
            const bool xin = (unsigned)xp < (unsigned)th.Sx;
            const uint32_t prow = src[k] - (th.vc * 16 + 8 * th.vh);             // start of the previous row's grid
            uint32_t rr16[8], cen[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                const int sy = 8 * th.vh - 2 + j;                                 // label row of slot j (pads: outside [0, Sy))
                const int yp = min(max(pv_x86_d2i(__dadd_rn(__dadd_rn((double)sy, ddy), 0.5)), -3), th.Sy + 2);
                const bool real = (unsigned)sy < (unsigned)th.Sy;
                // y-window minimum of the lane's OWN column at the predecessor row (pads 255 stand for "outside")
                const uint32_t q = prow + th.vc * 16 + 2 + yp;
                uint32_t r5 = min(min(min(pv_lds8(q - 2), pv_lds8(q - 1)), min(pv_lds8(q), pv_lds8(q + 1))), pv_lds8(q + 2));
                rr16[j] = (real && vact) ? r5 : 255u;
                cen[j] = (real && vact && xin) ? pv_lds8(prow + xp * 16 + 2 + yp) : 255u;
            }
            if (vact) pv_sts128(th.scr_s + k * (PV_RGRID * 2) + (5 * 16) * 2,
                                make_uint4(rr16[0] | (rr16[1] << 16), rr16[2] | (rr16[3] << 16), rr16[4] | (rr16[5] << 16), rr16[6] | (rr16[7] << 16)));
            __syncwarp();
            // x-window around the predecessor column of the grid just written
            const uint32_t rq = th.scr_s - (th.vc * 16 + 8 * th.vh) * 2 + k * (PV_RGRID * 2) + ((xp + 3) * 16 + 8 * th.vh) * 2;
            const uint4 a0 = pv_lds128(rq), a1 = pv_lds128(rq + 32), a2 = pv_lds128(rq + 64), a3 = pv_lds128(rq + 96), a4 = pv_lds128(rq + 128);
            const uint32_t m5[4] = {pv_min3(pv_min3(a0.x, a1.x, a2.x), a3.x, a4.x), pv_min3(pv_min3(a0.y, a1.y, a2.y), a3.y, a4.y),
                                    pv_min3(pv_min3(a0.z, a1.z, a2.z), a3.z, a4.z), pv_min3(pv_min3(a0.w, a1.w, a2.w), a3.w, a4.w)};
            mm2 = 0xFFFFFFFFu;
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const uint32_t ce = cen[2 * i] | (cen[2 * i + 1] << 16);
                // pads and out-of-window cells read 255; far2 <= 255 - P1 in the no-wrap domain, so 255 + P1 never wins and
                // every winning candidate is >= M
                const uint32_t best = pv_min3(far2, __vminu2(m5[i], K255) + th.P1P1, ce);
                l[k][i] = __vminu2(cc[i] + best - MM, K255);
                mm2 = __vminu2(mm2, l[k][i]);
            }
            __syncwarp();
        }
        uint32_t m = vact ? min(mm2 & 0xFFFFu, mm2 >> 16) : 0xFFFFu;
        Mn[k] = __reduce_min_sync(0xffffffffu, m);
    }
    __syncwarp();                                                        // everyone is done with the R grids of this pixel
    // ---- new state rows, hand-overs -------------------------------------------------------------------------------------------
    uint2 pw[3];
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        constexpr int dxs[3] = {0, 1, -1};
        const int dx = dxs[k];
        pw[k] = make_uint2(__byte_perm(l[k][0], l[k][1], 0x6420), __byte_perm(l[k][2], l[k][3], 0x6420));
        if (vact) pv_sts64(st[k], pw[k]);
        pv_sts32(sm[k], Mn[k]);                                          // every lane writes the same value
        if (EDGE) {
            if (dx > 0 && xl == Wk - 1 && th.inbox_right) {
                uint8_t* dst = th.inbox_right + ((size_t)par * 2 + 0) * th.IBS + 16;
                if (vact) *reinterpret_cast<uint2*>(dst + th.vc * 16 + 8 * th.vh) = pw[k];
                if (lane == 0) *reinterpret_cast<uint32_t*>(dst + PITCH + 4) = Mn[k];
            }
            if (dx < 0 && xl == 0 && th.inbox_left) {
                uint8_t* dst = th.inbox_left + ((size_t)par * 2 + 1) * th.IBS + 16;
                if (vact) *reinterpret_cast<uint2*>(dst + th.vc * 16 + 8 * th.vh) = pw[k];
                if (lane == 0) *reinterpret_cast<uint32_t*>(dst + PITCH + 4) = Mn[k];
            }
        }
    }
    if (EDGE && arrive) pv_cluster_arrive();
    // ---- sum of the three directions ------------------------------------------------------------------------------------------
    uint32_t acc[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) acc[i] = l[0][i] + l[1][i] + l[2][i];
    const size_t vox = (size_t)pix * PITCH;
    if (!FINAL) {
        // real slots: at most 3*(25+P2) <= 255; pad slots hold garbage (masked in the final pass)
        if (vact) *reinterpret_cast<uint2*>(th.S1out_l + vox) = make_uint2(__byte_perm(acc[0], acc[1], 0x6420), __byte_perm(acc[2], acc[3], 0x6420));
        return;
    }
    // both horizontal rows (pads masked off; byte-wise sum of real slots <= 2*(25+P2), no carry) and the down pass's byte row
    const uint32_t hx = (gH0.x & th.bmask.x) + (gH1.x & th.bmask.x), hy = (gH0.y & th.bmask.y) + (gH1.y & th.bmask.y);
    acc[0] += __byte_perm(hx, 0, 0x4140) + __byte_perm(gS1.x, 0, 0x4140);
    acc[1] += __byte_perm(hx, 0, 0x4342) + __byte_perm(gS1.x, 0, 0x4342);
    acc[2] += __byte_perm(hy, 0, 0x4140) + __byte_perm(gS1.y, 0, 0x4140);
    acc[3] += __byte_perm(hy, 0, 0x4342) + __byte_perm(gS1.y, 0, 0x4342);
#pragma unroll
    for (int i = 0; i < 4; ++i) acc[i] |= th.padmask[i];                 // pad slots never win
    // winner-take-all: first minimum in label order d = sx*Sy + sy (:298-314) via (sum << 16 | d)
    uint32_t key = 0xFFFFFFFFu;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        key = min(key, __byte_perm(acc[i], 2 * i, 0x1054));
        key = min(key, __byte_perm(acc[i], 2 * i + 1, 0x3254));
    }
    key = vact ? key + (uint32_t)th.base_d : 0xFFFFFFFFu;                // slot j of this lane is label base_d + j (pads cannot win)
    const uint32_t gs = th.scr_s + 3 * (PV_RGRID * 2) + 16 * 2;          // sum grid: one pad column in front (lane offset folded in)
    if (vact) pv_sts128(gs, make_uint4(acc[0], acc[1], acc[2], acc[3]));
    key = __reduce_min_sync(0xffffffffu, key);
    __syncwarp();
    if (lane == 0) {
        const uint32_t idx = key & 0xFFFFu, best = key >> 16;
        const uint32_t lx = idx / (uint32_t)th.Sy, ly = idx - lx * (uint32_t)th.Sy;
        // lane 0 is column 0, half 0: its folded offset is zero, so gs is the grid's column 0 / slot -2
        const uint32_t pos = gs + (lx * 16 + 2 + ly) * 2;
        uint32_t ay, by, ax, bx;
        asm volatile("ld.shared.u16 %0, [%1];" : "=r"(ay) : "r"(pos - 2));
        asm volatile("ld.shared.u16 %0, [%1];" : "=r"(by) : "r"(pos + 2));
        asm volatile("ld.shared.u16 %0, [%1];" : "=r"(ax) : "r"(pos - 32));
        asm volatile("ld.shared.u16 %0, [%1];" : "=r"(bx) : "r"(pos + 32));
        th.rec[pix] = make_uint4(idx | (best << 16), ay | (by << 16), ax | (bx << 16), 0u);
    }
    __syncwarp();
}

template <bool FINAL>
__global__ void __launch_bounds__(PV_WARPS * 32, 1)
pydv_kernel(const PvParams prm)
{
    cg::cluster_group cluster = cg::this_cluster();
    const int CS = (int)cluster.num_blocks(), rank = (int)cluster.block_rank();
    const int pair = blockIdx.x / CS;
    const int W = prm.W, H = prm.H, Wk_max = prm.Wk, Sx = prm.Sx, Sy = prm.Sy;
    const int PITCH = 16 * Sx, IBS = PITCH + 32;                            // inbox: 16 guard bytes (255), the row, 4 guard bytes, the minimum
    const int xb = rank * Wk_max, Wk = min(Wk_max, W - xb);
    const size_t N = (size_t)W * H;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;

    extern __shared__ __align__(128) unsigned char pv_smem[];
    uint8_t* guard0 = pv_smem;                                             // 16 bytes of 255 in front of the first grid
    uint8_t* cbuf = guard0 + 16;                                           // [2][Wk_max*PITCH]
    uint8_t* state = cbuf + 2 * (size_t)Wk_max * PITCH;                    // [3][Wk_max][PITCH]
    uint8_t* guard1 = state + 3 * (size_t)Wk_max * PITCH;                  // 16 bytes of 255 behind the last grid
    uint32_t* stmin = reinterpret_cast<uint32_t*>(guard1 + 16);            // [3][Wk_max]
    uint8_t* inbox = reinterpret_cast<uint8_t*>(stmin) + (((size_t)3 * Wk_max * 4 + 15) & ~(size_t)15);   // [2][2][IBS]
    uint8_t* scr = inbox + 4 * (size_t)IBS;                                // [PV_WARPS][PV_WSCR]
    uint64_t* bars = reinterpret_cast<uint64_t*>(scr + (size_t)PV_WARPS * PV_WSCR);

    const uint8_t* Cb = prm.C + pair * N * PITCH;
    const uint32_t row_bytes = (uint32_t)Wk * PITCH;
    auto row_y = [&](int yy) { return prm.up ? H - 1 - yy : yy; };

    // pads: every grid byte starts at 255 (state rows keep 255 in their pad slots from then on: cost pads are 255 and saturate)
    {
        uint32_t* p32 = reinterpret_cast<uint32_t*>(pv_smem);
        const size_t words = (size_t)(reinterpret_cast<uint8_t*>(bars) - pv_smem) / 4;
        for (size_t i = threadIdx.x; i < words; i += blockDim.x) p32[i] = 0xFFFFFFFFu;
    }
    __syncthreads();
    {   // the u16 grids of the per-warp scratch: pad entries are the VALUE 255 (P1 is added to them in 16-bit lanes)
        uint32_t* p32 = reinterpret_cast<uint32_t*>(scr);
        for (int i = threadIdx.x; i < PV_WARPS * PV_WSCR / 4; i += blockDim.x) p32[i] = 0x00FF00FFu;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        pv_mbar_init(&bars[0], 1);
        pv_mbar_init(&bars[1], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");       // the generic-proxy fill above precedes the bulk copies
    }
    cluster.sync();
    if (threadIdx.x == 0) {
        for (int yy = 0; yy < min(2, H); ++yy) {
            pv_mbar_expect_tx(&bars[yy], row_bytes);
            pv_tma_load_1d(cbuf + (size_t)yy * Wk_max * PITCH, Cb + ((size_t)row_y(yy) * W + xb) * PITCH, row_bytes, &bars[yy]);
        }
    }
    uint8_t* inbox_right = (rank + 1 < CS) ? cluster.map_shared_rank(inbox, rank + 1) : nullptr;
    uint8_t* inbox_left = (rank > 0) ? cluster.map_shared_rank(inbox, rank - 1) : nullptr;

    PvThread th;
    th.lane = lane; th.vc = lane >> 1; th.vh = lane & 1;
    th.vact = th.vc < Sx;
    const int loff = th.vc * 16 + 8 * th.vh;                               // this lane's byte offset inside a grid row
    const uint32_t cbuf_s = pv_smem_u32(cbuf) + loff;
    th.state_s = pv_smem_u32(state) + loff; th.stmin_s = pv_smem_u32(stmin); th.inbox_s = pv_smem_u32(inbox);
    th.scr_s = pv_smem_u32(scr) + warp * PV_WSCR + loff * 2;               // u16 grids: the lane's entry offset is 2 bytes per slot
    th.inbox_right = inbox_right; th.inbox_left = inbox_left;
    th.H0_l = prm.H0 ? prm.H0 + pair * N * PITCH + loff : nullptr;
    th.H1_l = prm.H1 ? prm.H1 + pair * N * PITCH + loff : nullptr;
    th.S1in_l = prm.S1in ? prm.S1in + pair * N * PITCH + loff : nullptr;
    th.S1out_l = prm.S1out ? prm.S1out + pair * N * PITCH + loff : nullptr;
    th.flags = prm.flags + pair * N;
    th.mvx = prm.preMv + (size_t)pair * 2 * prm.mvW * prm.mvH; th.mvy = th.mvx + (size_t)prm.mvW * prm.mvH;
    th.rec = prm.rec ? prm.rec + pair * N : nullptr;
    th.Wk = Wk; th.Wk_max = Wk_max; th.W = W; th.H = H; th.xb = xb; th.Sx = Sx; th.Sy = Sy; th.PITCH = PITCH; th.IBS = IBS;
    th.mvW = prm.mvW; th.up = prm.up;
    th.P1P1 = (uint32_t)prm.P1 * 0x10001u; th.P2P2 = (uint32_t)prm.P2 * 0x10001u;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int s0 = 8 * th.vh - 2 + 2 * i;                              // label rows of the register's two halves
        th.padmask[i] = (((unsigned)s0 < (unsigned)Sy) ? 0u : 0xFFFFu) | (((unsigned)(s0 + 1) < (unsigned)Sy) ? 0u : 0xFFFF0000u);
    }
    th.base_d = th.vc * Sy + 8 * th.vh - 2;
    th.bmask = make_uint2(~__byte_perm(th.padmask[0], th.padmask[1], 0x6420), ~__byte_perm(th.padmask[2], th.padmask[3], 0x6420));

    // pixel order inside a row, barrier protocol and the relief rule are those of vsweep_kernel (vsweep.cu)
    constexpr int HW = PV_WARPS / 2;
    const int wsub = warp % HW;
    const bool upper = warp >= HW;
    const int n_lo = (Wk + 1) / 2, cnt = upper ? Wk - n_lo : n_lo;
    const int xfirst = upper ? Wk - 1 : 0, xstep = upper ? -1 : 1;
    int lim = cnt, imax = 0x7FFFFFFF;
    {
        const int full = cnt / HW, rem = cnt % HW;
        if (full >= 1 && rem > 0) {
            if (wsub == 0) lim = full * HW;
            if (wsub == rem) { lim = cnt + HW; imax = full * HW; }
        }
    }
    auto xl_of = [&](int i) { return xfirst + xstep * min(i, imax); };

    int off = 0;
    for (int yy = 0; yy < H; ++yy) {
        const int y = row_y(yy), par = yy & 1;
        if (yy > 0) pv_cluster_wait();
        pv_mbar_wait(&bars[par], (uint32_t)((yy >> 1) & 1));
        const uint32_t crow_l = cbuf_s + (uint32_t)(par * Wk_max * PITCH);
        const uint32_t rowpix = (uint32_t)y * (uint32_t)W + (uint32_t)xb;
        if (wsub != 0 || lim == 0) pv_cluster_arrive_relaxed();
        // the three byte rows of the final pass are requested one pixel ahead
        uint2 g0 = make_uint2(0, 0), g1 = g0, g2 = g0;
        uint32_t gf = 0;
        auto fetch = [&](int i) {
            gf = __ldg(th.flags + rowpix + xl_of(i));
            if (FINAL && th.vact) {
                const size_t vox = (size_t)(rowpix + xl_of(i)) * PITCH;
                g0 = __ldg(reinterpret_cast<const uint2*>(th.H0_l + vox));
                g1 = __ldg(reinterpret_cast<const uint2*>(th.H1_l + vox));
                g2 = __ldg(reinterpret_cast<const uint2*>(th.S1in_l + vox));
            }
        };
        if (wsub < lim) fetch(wsub);
        for (int i = wsub; i < lim; i += HW) {
            const int xl = xl_of(i);
            const uint2 c0 = g0, c1 = g1, c2 = g2;
            const uint32_t cf = gf;
            if (i + HW < lim) fetch(i + HW);
            if (yy == 0) pv_pixel<FINAL, true>(th, crow_l, xl, yy, y, par, off, rowpix + xl, cf, c0, c1, c2, i == 0);
            else if (i == 0) pv_pixel<FINAL, true>(th, crow_l, xl, yy, y, par, off, rowpix + xl, cf, c0, c1, c2, true);
            else pv_pixel<FINAL, false>(th, crow_l, xl, yy, y, par, off, rowpix + xl, cf, c0, c1, c2, false);
        }
        if (++off == Wk) off = 0;
        __syncthreads();
        if (threadIdx.x == 0 && yy + 2 < H) {
            pv_mbar_expect_tx(&bars[par], row_bytes);
            pv_tma_load_1d(cbuf + (size_t)par * Wk_max * PITCH, Cb + ((size_t)row_y(yy + 2) * W + xb) * PITCH, row_bytes, &bars[par]);
        }
    }
    pv_cluster_wait();
}

// per pixel: bit r set when direction r's step at this pixel has a non-zero prior difference (:213-254) and must take the
// general form; a pixel whose predecessor lies outside the image starts a path and needs no flag
__global__ void pyd_shift_flags_kernel(const double* __restrict__ preMv, int mvW, int mvH, int W, int H, uint8_t* __restrict__ flags)
{
    const size_t N = (size_t)W * H;
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= N) return;
    const int y = (int)(i / W), x = (int)(i - (size_t)y * W);
    const double* mvx = preMv + (size_t)blockIdx.y * 2 * mvW * mvH;
    const double* mvy = mvx + (size_t)mvW * mvH;
    const double cx = mvx[(size_t)y * mvW + x], cy = mvy[(size_t)y * mvW + x];
    uint32_t f = 0;
#pragma unroll
    for (int r = 0; r < 8; ++r) {
        const int px = x - dir_dx(r), py = y - dir_dy(r);
        if (px < 0 || px >= W || py < 0 || py >= H) continue;
        const double ddx = __dsub_rn(cx, mvx[(size_t)py * mvW + px]), ddy = __dsub_rn(cy, mvy[(size_t)py * mvW + px]);
        if (!(ddx == 0.0 && ddy == 0.0)) f |= 1u << r;
    }
    flags[blockIdx.y * N + i] = (uint8_t)f;
}

// bestD, minC and the per-axis parabola (calc_pyd_cost_sgm.cpp:333-360) from the WTA records
__global__ void pydv_finalize_kernel(const uint4* __restrict__ rec, size_t N, int Sx, int Sy, int subpixel,
                                     uint32_t* __restrict__ bestD, uint32_t* __restrict__ minC, double* __restrict__ mvSub)
{
    const size_t p = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= N) return;
    const size_t gp = blockIdx.y * N + p;
    const uint4 r = rec[gp];
    const uint32_t idx = r.x & 0xFFFFu, best = r.x >> 16;
    bestD[gp] = idx;
    minC[gp] = best;
    double sx = 0.0, sy = 0.0;
    if (subpixel) {
        const int lx = (int)idx / Sy, ly = (int)idx - lx * Sy;
        const double c0 = (double)best;
        if (ly > 0 && ly < Sy - 1) {
            const double a = (double)(r.y & 0xFFFFu), b = (double)(r.y >> 16);
            sy = (b < a) ? __ddiv_rn(__ddiv_rn(__dsub_rn(b, a), __dsub_rn(c0, a)), 2.0)
                         : __ddiv_rn(__ddiv_rn(__dsub_rn(b, a), __dsub_rn(c0, b)), 2.0);
        }
        if (lx > 0 && lx < Sx - 1) {
            const double a = (double)(r.z & 0xFFFFu), b = (double)(r.z >> 16);
            sx = (b < a) ? __ddiv_rn(__ddiv_rn(__dsub_rn(b, a), __dsub_rn(c0, a)), 2.0)
                         : __ddiv_rn(__ddiv_rn(__dsub_rn(b, a), __dsub_rn(c0, b)), 2.0);
        }
    }
    mvSub[blockIdx.y * 2 * N + p] = sx;
    mvSub[blockIdx.y * 2 * N + N + p] = sy;
}

// ---- host side ----------------------------------------------------------------------------------------------
static size_t pv_smem_bytes(int Sx, int Wk)
{
    const size_t PITCH = 16 * (size_t)Sx;
    return 16 + 5 * (size_t)Wk * PITCH + 16 + (((size_t)3 * Wk * 4 + 15) & ~(size_t)15) + 4 * (PITCH + 32) + (size_t)PV_WARPS * PV_WSCR + 64;
}

bool pydv_applicable(int Sx, int Sy, int P1, int P2, int diag, int passes, int adaptive)
{
    return Sx >= 1 && Sy >= 1 && Sx <= 11 && Sy <= 11 && diag && passes == 2 && !adaptive &&
           P1 >= 0 && P2 >= 0 && 25 + P1 + P2 <= 255 && 50 + P2 <= 255 && 3 * (25 + P2) <= 255;
}

static bool pv_cluster_fits(int W, int Sx, int cs, int max_smem)
{
    const int Wk = (W + cs - 1) / cs;
    if (Wk < 2 || (cs - 1) * Wk >= W) return false;
    if (cs > 1 && W - (cs - 1) * Wk < 2) return false;
    return pv_smem_bytes(Sx, Wk) <= (size_t)max_smem;
}

static int pv_max_clusters(int cs, size_t smem)
{
    auto kern = pydv_kernel<true>;
    cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    cudaFuncSetAttribute(kern, cudaFuncAttributeNonPortableClusterSizeAllowed, 1);
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(cs * 64); cfg.blockDim = dim3(PV_WARPS * 32); cfg.dynamicSmemBytes = smem;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = cs; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
    cfg.attrs = at; cfg.numAttrs = 1;
    int n = -1;
    if (cudaOccupancyMaxActiveClusters(&n, kern, &cfg) != cudaSuccess) { cudaGetLastError(); return -1; }
    return n;
}

// Cluster size for n pairs of width W: the size with the shortest estimated batch time — waves of resident clusters x rows x
// pixel rounds per row (one round = one pixel per warp, each half of a strip walked by PV_WARPS/2 warps).  Larger clusters
// spread a small batch over more SMs; 0 = the path does not fit.
int pydv_pick_cluster(fsgm_ctx* c, int n, int W, int Sx, int forced)
{
    constexpr int MAXS = 227 * 1024;
    if (forced > 0) return pv_cluster_fits(W, Sx, forced, MAXS) ? forced : 0;
    int best = 0; double best_t = 0;
    for (int cs = 1; cs <= 16; ++cs) {
        if (!pv_cluster_fits(W, Sx, cs, MAXS)) continue;
        const int Wk = (W + cs - 1) / cs;
        int k = 0;
        const int key = cs * 4096 + Wk;                                    // occupancy depends on (cs, smem)
        for (auto& e : c->pv_occ) if (e.first == key * 16 + Sx) k = e.second;
        if (!k) { k = pv_max_clusters(cs, pv_smem_bytes(Sx, Wk)); c->pv_occ.push_back({key * 16 + Sx, k}); }
        if (k < 1) continue;
        const int rounds = ((Wk + 1) / 2 + PV_WARPS / 2 - 1) / (PV_WARPS / 2);
        const int waves = (n + k - 1) / k;
        const double t = (double)waves * (rounds + 0.6);                   // + the per-row barrier / hand-over overhead
        if (!best || t < best_t * 0.98) { best = cs; best_t = t; }
    }
    return best;
}

int launch_pyd_shift_flags(fsgm_ctx* c, int n, const double* preMv, int mvW, int mvH, int W, int H, uint8_t* flags)
{
    StageScope ss(c, ST_PYD_SWEEP);
    const size_t N = (size_t)W * H;
    pyd_shift_flags_kernel<<<dim3((unsigned)((N + 255) / 256), n), 256, 0, c->stream>>>(preMv, mvW, mvH, W, H, flags);
    FSGM_LAUNCHED(c);
    return FSGM_OK;
}

int launch_pydv(fsgm_ctx* c, int n, int cs, bool final_, const uint8_t* C, const uint8_t* H0, const uint8_t* H1, const uint8_t* S1in,
                uint8_t* S1out, uint4* rec, const uint8_t* flags, const double* preMv, int mvW, int mvH, int W, int H, int Sx, int Sy,
                int P1, int P2)
{
    StageScope ss(c, ST_PYD_SWEEP);
    PvParams p{};
    p.C = C; p.H0 = H0; p.H1 = H1; p.S1in = S1in; p.S1out = S1out; p.rec = rec; p.flags = flags; p.preMv = preMv;
    p.mvW = mvW; p.mvH = mvH; p.W = W; p.H = H; p.Wk = (W + cs - 1) / cs; p.Sx = Sx; p.Sy = Sy; p.P1 = P1; p.P2 = P2; p.up = final_ ? 1 : 0;
    const size_t smem = pv_smem_bytes(Sx, p.Wk);
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(cs * n); cfg.blockDim = dim3(PV_WARPS * 32); cfg.dynamicSmemBytes = smem; cfg.stream = c->stream;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = cs; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
    cfg.attrs = at; cfg.numAttrs = 1;
    if (final_) {
        auto kern = pydv_kernel<true>;
        FSGM_CUDA(c, cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
        FSGM_CUDA(c, cudaFuncSetAttribute(kern, cudaFuncAttributeNonPortableClusterSizeAllowed, 1));
        FSGM_CUDA(c, cudaLaunchKernelEx(&cfg, kern, p));
    } else {
        auto kern = pydv_kernel<false>;
        FSGM_CUDA(c, cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
        FSGM_CUDA(c, cudaFuncSetAttribute(kern, cudaFuncAttributeNonPortableClusterSizeAllowed, 1));
        FSGM_CUDA(c, cudaLaunchKernelEx(&cfg, kern, p));
    }
    c->launches++;
    return FSGM_OK;
}

int launch_pydv_finalize(fsgm_ctx* c, int n, const uint4* rec, int W, int H, int Sx, int Sy, int subpixel,
                         uint32_t* bestD, uint32_t* minC, double* mvSub)
{
    StageScope ss(c, ST_PYD_WTA);
    const size_t N = (size_t)W * H;
    pydv_finalize_kernel<<<dim3((unsigned)((N + 255) / 256), n), 256, 0, c->stream>>>(rec, N, Sx, Sy, subpixel, bestD, minC, mvSub);
    FSGM_LAUNCHED(c);
    return FSGM_OK;
}

}  // namespace fsgm
