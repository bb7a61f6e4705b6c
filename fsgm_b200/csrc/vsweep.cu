// Row-synchronous aggregation of the non-horizontal directions — the fast path for the sweeps of sgm()
// (reference calc_cost_sgm.cpp:114-257) when the image is narrow enough for its path state to live on chip.
//
// The generic kernel (aggregate.cu) runs every direction as independent scanlines: each direction reads the cost
// volume C once and writes its own u8 volume L_r, and a WTA kernel reads all of them back: 3 B per voxel and
// direction.  The three directions that advance one image row per step (down: (0,+1) (+1,+1) (-1,+1); up: their
// negatives) all need the SAME cost row at the same time, so here ONE thread-block cluster walks an image row by row:
//
//   * the image width is split over the CTAs of the cluster; each CTA keeps the L rows of its columns for the three
//     directions in shared memory (3 * Wk * D bytes), updated in place: a diagonal path's state stays in one slot
//     ((x - dx*row) mod Wk) while the pixel it belongs to slides across the strip;
//   * a path that leaves the strip is handed to the neighbour CTA through distributed shared memory (256 B + its
//     minimum per row and diagonal), one cluster barrier per row orders the exchange;
//   * the strip's cost row is fetched once per row by a 1-D bulk TMA copy (cp.async.bulk + mbarrier), two rows ahead;
//   * per pixel the three new L rows are summed in registers; the DOWN pass adds the two horizontal volumes and
//     writes one u16 sum; the UP pass adds that sum and does winner-take-all directly, so neither the six L volumes
//     nor Sp ever exist in HBM.
//
// HBM traffic per voxel: down pass 1 (C) + 2 (two horizontal u8 volumes) + 2 (u16 sum out); up pass 1 (C) + 2 (sum in);
// against 6 * 3 = 18 for the same six directions through the generic kernel + WTA.  Bound: integer issue
// (~55 instructions per pixel and direction), with the down pass close to the HBM roofline as well.
//
// Used only when: D in {64,128,256}, parameters inside the no-wrap domain, no adaptive P2, and the state fits in
// shared memory with a cluster of <= 8 CTAs (W*D <= ~320 k, e.g. 1242 x 256).  Everything else takes the generic path.
#include "fsgm_internal.h"
#include "sgm_step.cuh"
#include <cooperative_groups.h>
#include <algorithm>

namespace cg = cooperative_groups;

namespace fsgm {

#ifndef FSGM_VS_PDF
#define FSGM_VS_PDF 3
#endif
#ifndef FSGM_VS_DEFER
#define FSGM_VS_DEFER 0              // A/B: 1 = final pass consumes a pixel's global rows / runs its WTA after the NEXT pixel's direction steps.
                                     // Measured slower like every other way of moving the first use of the rows away from the row start:
                                     // 2060 -> 2021 pairs/s with a ring of 3 or 4 pixels, 2045 with a ring of 2 (r2x).
#endif
#ifndef FSGM_VS_RELIEF
#define FSGM_VS_RELIEF 1
#endif
constexpr int VS_WARPS = 20;      // measured at KITTI size (156 columns per CTA): 13/16/18/20/24/26 warps -> 14.9/13.1/13.5/12.7/13.3/14.2 ms per 30 pairs

struct VsParams {
    const uint8_t* C;            // [n][H][W][D]
    const uint8_t* addA;         // optional u8 volumes added to the sum (the horizontal directions)
    const uint8_t* addB;
    const uint16_t* Sin;         // optional u16 volume added to the sum (the other pass)
    uint16_t* Sout;              // !FINAL: u16 sum volume; FINAL: optional dump of the total Sp (stage parity), may be null
    uint32_t* minC;              // FINAL: winner-take-all outputs
    uint16_t* rec;               //        [n][N][4] = argmin, Sp[argmin-1], Sp[argmin+1] (don't-care at argmin 0 / D-1), Sp[0]
    int W, H, Wk, P1, P2;
    int up;                      // 0: rows 0..H-1 with directions (0,+1)(+1,+1)(-1,+1); 1: rows H-1..0, negated
    int fast;                    // the FAST operand configuration (see vs_fetch): Sin / Sout are BYTE volumes
};

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, int count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity)
{
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "WAIT_%=:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra DONE_%=;\n\t"
        "bra WAIT_%=;\n\t"
        "DONE_%=:\n\t}" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ void tma_load_1d(void* dst, const void* src, uint32_t bytes, uint64_t* bar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}

__device__ __forceinline__ void cluster_arrive() { asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory"); }
// for warps that publish nothing: a release would also wait for their global stores to reach L2
__device__ __forceinline__ void cluster_arrive_relaxed() { asm volatile("barrier.cluster.arrive.relaxed.aligned;" ::: "memory"); }
__device__ __forceinline__ void cluster_wait() { asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory"); }

template <int NREG> __device__ __forceinline__ void ld_row(const uint8_t* base, int lane, uint32_t (&w)[(NREG + 1) / 2])
{
    if (NREG == 1) w[0] = reinterpret_cast<const uint16_t*>(base)[lane];
    else if (NREG == 2) w[0] = reinterpret_cast<const uint32_t*>(base)[lane];
    else { uint2 v = reinterpret_cast<const uint2*>(base)[lane]; w[0] = v.x; w[1] = v.y; }
}
template <int NREG> __device__ __forceinline__ void st_row(uint8_t* base, int lane, const uint32_t (&w)[(NREG + 1) / 2])
{
    if (NREG == 1) reinterpret_cast<uint16_t*>(base)[lane] = (uint16_t)w[0];
    else if (NREG == 2) reinterpret_cast<uint32_t*>(base)[lane] = w[0];
    else reinterpret_cast<uint2*>(base)[lane] = make_uint2(w[0], w[1]);
}

// Shared-memory accesses of the pixel body go through 32-bit shared-space addresses: with generic pointers the compiler rebuilds
// the CTA's shared window base from %cluster_ctaid for every pixel (S2R + LEA in front of the first LDS of the dependency chain).
__device__ __forceinline__ uint32_t lds_u32(uint32_t a) { uint32_t v; asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(a)); return v; }
__device__ __forceinline__ void sts_u32(uint32_t a, uint32_t v) { asm volatile("st.shared.u32 [%0], %1;" ::"r"(a), "r"(v)); }
template <int NREG> __device__ __forceinline__ void lds_row(uint32_t a, uint32_t (&w)[(NREG + 1) / 2])
{
    if (NREG == 1) { uint16_t h; asm volatile("ld.shared.u16 %0, [%1];" : "=h"(h) : "r"(a)); w[0] = h; }
    else if (NREG == 2) asm volatile("ld.shared.u32 %0, [%1];" : "=r"(w[0]) : "r"(a));
    else asm volatile("ld.shared.v2.u32 {%0, %1}, [%2];" : "=r"(w[0]), "=r"(w[(NREG + 1) / 2 - 1]) : "r"(a));
}
template <int NREG> __device__ __forceinline__ void sts_row(uint32_t a, const uint32_t (&w)[(NREG + 1) / 2])
{
    if (NREG == 1) asm volatile("st.shared.u16 [%0], %1;" ::"r"(a), "h"((uint16_t)w[0]));
    else if (NREG == 2) asm volatile("st.shared.u32 [%0], %1;" ::"r"(a), "r"(w[0]));
    else asm volatile("st.shared.v2.u32 [%0], {%1, %2};" ::"r"(a), "r"(w[0]), "r"(w[(NREG + 1) / 2 - 1]));
}

template <int NREG>
struct VsThread {
    uint32_t state_s, stmin_s, inbox_s;            // shared-space addresses (state_s with the lane's byte offset folded in)
    uint8_t* inbox_right; uint8_t* inbox_left;     // neighbours' inboxes: distributed shared memory, generic pointers
    const uint8_t* addA_l; const uint8_t* addB_l; const uint16_t* Sin_l; uint16_t* Sout_l;
    uint32_t* minC; uint16_t* rec; uint16_t* ws;
    int Wk, Wk_max, W, xb, lane;
    uint32_t P1P1, P2P2, sel_lo, sel_hi;
    bool preadd;
};

// the rows a pixel needs from global memory (horizontal volumes / the other pass's sum), fetched two pixels ahead
template <int NREG>
struct VsGlobals { uint32_t a[(NREG + 1) / 2], b[(NREG + 1) / 2], s[NREG], s8[(NREG + 1) / 2]; };     // s8: the FAST byte volume

// FAST = the standard two-pass run (both horizontal volumes present, no Sp dump, 3*P2 <= 255, 8*(24+P2) < 1024), with the
// operand configuration known at compile time and the HBM traffic of the first pass cut to the minimum:
//   * first pass (!FINAL): NO global loads at all.  It writes ONE BYTE per voxel, sum_k (L_k - C): every L_k - C is
//     min(b - M, P2) in [0, P2], so three of them fit a byte, and the subtraction is byte-wise on the packed words
//     (three IADD3 per word).  ncu (r1n): with the horizontal rows read here the pass spent a third of its stall samples
//     waiting for them although they were requested four pixels ahead (mixed read/write DRAM traffic at 4 TB/s).
//   * final pass: reads that byte volume and the two horizontal volumes (3 B per voxel, read-only traffic), rebuilds
//     1024 + 3*C + sum of all eight directions as a biased fp16 pattern (one HFMA2 for 3*C - 2048, the byte rows added as
//     integers onto the pattern: it stays below 0x6800) and runs winner-take-all on the patterns, which order like the sums;
//     the bias comes off the values that leave the kernel (minC here, the subpixel neighbours in vs_finalize_kernel).
// Every other combination (single pass, one horizontal volume, stage dumps, larger P2) takes the run-time checks and the
// u16 sum volume.
template <int NREG, bool FINAL, bool FAST>
__device__ __forceinline__ void vs_fetch(const VsThread<NREG>& th, uint32_t pix, VsGlobals<NREG>& g)
{
    constexpr int D = 64 * NREG;
    const size_t vox = (size_t)pix * D;
    if (FAST) {
        if (FINAL) {
            ld_row<NREG>(th.addA_l + vox, 0, g.a);
            ld_row<NREG>(th.addB_l + vox, 0, g.b);
            ld_row<NREG>(reinterpret_cast<const uint8_t*>(th.Sin_l) + vox, 0, g.s8);     // byte volume: Sin_l carries a BYTE lane offset here
        }
        return;
    }
    if (th.addA_l != nullptr) ld_row<NREG>(th.addA_l + vox, 0, g.a);
    if (th.addB_l != nullptr) ld_row<NREG>(th.addB_l + vox, 0, g.b);
    if (FINAL && th.Sin_l) {
        const uint32_t* sp = reinterpret_cast<const uint32_t*>(th.Sin_l + vox);
        if (NREG == 4) { uint4 v = *reinterpret_cast<const uint4*>(sp); g.s[0] = v.x; g.s[1] = v.y; g.s[2] = v.z; g.s[3] = v.w; }
        else if (NREG == 2) { uint2 v = *reinterpret_cast<const uint2*>(sp); g.s[0] = v.x; g.s[1] = v.y; }
        else g.s[0] = *sp;
    }
}

// winner-take-all over the accumulated sums of one pixel (lane = 2*NREG consecutive labels) + the record vs_finalize_kernel reads
template <int NREG, bool FAST>
__device__ __forceinline__ void vs_wta(const VsThread<NREG>& th, uint32_t (&acc)[NREG], uint32_t pix)
{
    const int lane = th.lane;
    // winner-take-all: first minimum (strict <, calc_cost_sgm.cpp:267) via (sum << 16 | label)
    uint32_t key = 0xFFFFFFFFu;
    uint16_t* ws = th.ws;
#pragma unroll
    for (int i = 0; i < NREG; ++i) {
        // (sum << 16 | position inside the lane) for the two labels of this register: one PRMT each, the position an
        // immediate; the lane's first label (a multiple of 2*NREG) is OR-ed in once after the lane-level minimum
        key = min(key, __byte_perm(acc[i], 2 * i, 0x1054));
        key = min(key, __byte_perm(acc[i], 2 * i + 1, 0x3254));
    }
    key |= (uint32_t)(lane * 2 * NREG);
    if (NREG == 4) reinterpret_cast<uint4*>(ws)[lane] = make_uint4(acc[0], acc[1], acc[2], acc[3]);
    else if (NREG == 2) reinterpret_cast<uint2*>(ws)[lane] = make_uint2(acc[0], acc[1]);
    else reinterpret_cast<uint32_t*>(ws)[lane] = acc[0];
    key = __reduce_min_sync(0xffffffffu, key);
    __syncwarp();
    if (lane == 0) {
        const uint32_t idx = key & 0xFFFFu;
        th.minC[pix] = (key >> 16) - (FAST ? (H2_BIAS2 & 0xFFFFu) : 0u);
        uint16_t* r = th.rec + (size_t)pix * 4;
        // unconditional neighbour reads: for idx == 0 / idx == D-1 they land in the adjacent shared-memory words (inside the
        // allocation) and vs_finalize_kernel never looks at those fields (label 0 and 1 are not refined, label D-1 takes the
        // next pixel's Sp[0]) — two predicated branches less in a section the whole warp waits for
        const uint16_t c_1 = ws[(int)idx - 1], c1 = ws[idx + 1];
        *reinterpret_cast<uint2*>(r) = make_uint2(idx | ((uint32_t)c_1 << 16), (uint32_t)c1 | (acc[0] << 16));      // lane 0's acc[0] low half = Sp[0]
    }
    __syncwarp();
}

// Tail of a FAST final-pass pixel: both horizontal rows (their byte-wise sum cannot carry: 2*(cmax+P2) <= 255) and the first pass's
// byte row are added as integers onto the biased pattern, then winner-take-all.  Separate from vs_pixel so that the row loop can
// run it one pixel late (FSGM_VS_DEFER, an A/B build): the rows were requested PD pixels ahead, but behind the row barrier every
// warp's ring is empty and the first pixel waits for DRAM (ncu r2s: 17.5 % of the pass's stall samples on the first use of the rows);
// with the tail of pixel i behind the direction steps of pixel i+1 the first use comes two pixel-times after the request — and the
// pass gets slower (see the macro).
template <int NREG>
__device__ __forceinline__ void vs_tail_fast(const VsThread<NREG>& th, uint32_t (&acc)[NREG], const VsGlobals<NREG>& g, uint32_t pix)
{
    constexpr int NW = (NREG + 1) / 2;
    uint32_t ab[NW], t[NREG], u[NREG];
#pragma unroll
    for (int i = 0; i < NW; ++i) ab[i] = g.a[i] + g.b[i];
    unpack_cost<NREG>(ab, t);
    unpack_cost<NREG>(g.s8, u);
#pragma unroll
    for (int i = 0; i < NREG; ++i) acc[i] += t[i] + u[i];
    vs_wta<NREG, true>(th, acc, pix);
}

// One pixel of one row: all NDIR directions, sum, output.  EDGE = the pixel may restart a path, take one from a
// neighbour CTA's hand-over, or hand one over (first row, first / last column of the strip); interior pixels compile
// to a straight line of LDS -> step -> STS per direction.
template <int NREG, int NDIR, bool FINAL, bool EDGE, bool FAST, bool DEFER = false>
__device__ __forceinline__ void vs_pixel(const VsThread<NREG>& th, uint32_t crow_s, int xl, int yy, int par, int off, uint32_t pix,
                                         const VsGlobals<NREG>& g, bool arrive, uint32_t* acc_out = nullptr)
{
    constexpr int D = 64 * NREG, NW = (NREG + 1) / 2, NB = 2 * NREG;
    const int lane = th.lane, Wk = th.Wk;
    const size_t vox = (size_t)pix * D;
    // cost, path state and minima are biased fp16x2 numbers (sgm_step.cuh): min on the ALU pipe, add on the FMA pipe
    uint32_t cw[NW], c[NREG], cP2[NREG], acc[NREG];
    lds_row<NREG>(crow_s + xl * D, cw);
    unpack_cost_h2<NREG>(cw, c);
#pragma unroll
    for (int i = 0; i < NREG; ++i) cP2[i] = h2_add(c[i], th.P2P2);
    // The three directions are independent.  All previous-state rows are loaded first and all new rows stored last:
    // with the loads and stores of one direction between those of another the compiler must assume they alias and
    // serialises the three dependency chains (ncu r1h: fixed-latency "wait" stalls were the largest stall class).
    // k = 0 is the vertical direction, k = 1 the diagonal with dx = +1, k = 2 the one with dx = -1 (in both passes: which of the
    // two is "L2" and which "L4" does not matter, only their sum leaves the kernel)
    uint32_t st[NDIR], sm[NDIR];
    uint32_t lw[NDIR][NW], Mv[NDIR];
    bool restart[NDIR];
#pragma unroll
    for (int k = 0; k < NDIR; ++k) {
        constexpr int dxs[3] = {0, 1, -1};
        const int dx = dxs[k];
        int slot = xl;
        if (dx > 0) { slot = xl - off; if (slot < 0) slot += Wk; }
        if (dx < 0) { slot = xl + off; if (slot >= Wk) slot -= Wk; }
        st[k] = th.state_s + (k * th.Wk_max + slot) * D;
        sm[k] = th.stmin_s + (k * th.Wk_max + slot) * 4;
        restart[k] = false;
        uint32_t src = st[k];
        bool boxed = false;
        if (EDGE) {
            const int x = th.xb + xl;
            restart[k] = (yy == 0) || (dx > 0 && x == 0) || (dx < 0 && x == th.W - 1);
            const bool from_left = dx > 0 && xl == 0, from_right = dx < 0 && xl == Wk - 1;
            if (!restart[k] && (from_left || from_right)) {
                src = th.inbox_s + (uint32_t)((((yy - 1) & 1) * 2 + (from_right ? 1 : 0)) * (D + 16));
                Mv[k] = lds_u32(src + D);
                src += lane * NB;
                boxed = true;
            }
        }
        if (!boxed) Mv[k] = lds_u32(sm[k]);
        lds_row<NREG>(src, lw[k]);
    }
    uint32_t pw[NDIR][NW], Mn[NDIR];
    static_assert(NDIR == 1 || NDIR == 3, "accumulator bias below assumes one or three directions");
    // !FAST: -1024 * (NDIR - 1): the sum of the directions alone is 1024 + sum and is turned into an integer below.
    // FAST FINAL: the accumulator starts at NDIR*C + 1024 - NDIR*1024 (one HFMA2 on the biased cost: (1024 + C)*NDIR - 2*NDIR*1024
    // + 1024, exact: a single rounding of an integer of magnitude <= 2048), so that after the NDIR biased rows it holds
    // 1024 + NDIR*C + sum_k L_k with every partial sum an integer of magnitude <= 2048.
    constexpr uint32_t ACC0 = 0xE800E800u;
    constexpr uint32_t H2_NDIR = NDIR == 3 ? 0x42004200u : 0x3C003C00u;       // half2(3) / half2(1)
    constexpr uint32_t H2_K0 = NDIR == 3 ? 0xED00ED00u : 0xE400E400u;         // half2(-5120) / half2(-1024)
    if (FAST && FINAL) {
#pragma unroll
        for (int i = 0; i < NREG; ++i) asm("fma.rn.f16x2 %0, %1, %2, %3;" : "=r"(acc[i]) : "r"(c[i]), "r"(H2_NDIR), "r"(H2_K0));
    }
#pragma unroll
    for (int k = 0; k < NDIR; ++k) {
        uint32_t L[NREG];
        Mn[k] = H2_BIAS2;                              // biased 0 (both halves): the minimum a restarted path hands to its next step
        if (EDGE && restart[k]) {
#pragma unroll
            for (int i = 0; i < NREG; ++i) L[i] = c[i];
        } else {
            uint32_t Lpre[NREG];
            unpack_cost_h2<NREG>(lw[k], Lpre);
            // minima travel duplicated in both halves (state, inbox), ready for the fp16x2 add inside the step
            Mn[k] = sgm_step_h2<NREG>(cP2, Lpre, Mv[k], th.P1P1, th.P2P2, th.sel_lo, th.sel_hi, L);
        }
        pack_cost<NREG>(L, pw[k]);
        // sum of the directions, still on the FMA pipe: -1024*(NDIR-1) + sum(1024 + L_k) = 1024 + sum(L_k) < 2048, every
        // partial sum is an integer of magnitude <= 2048
#pragma unroll
        for (int i = 0; i < NREG; ++i) {
            if (FAST) { if (FINAL) acc[i] = h2_add(acc[i], L[i]); }
            else acc[i] = k == 0 ? (NDIR == 1 ? L[i] : h2_add(L[i], ACC0)) : h2_add(acc[i], L[i]);
        }
    }
    // biased half -> integer u16x2
    if (!FAST) {
#pragma unroll
        for (int i = 0; i < NREG; ++i) acc[i] -= H2_BIAS2;
    }
#pragma unroll
    for (int k = 0; k < NDIR; ++k) {
        constexpr int dxs[3] = {0, 1, -1};
        const int dx = dxs[k];
        sts_row<NREG>(st[k], pw[k]);
        sts_u32(sm[k], Mn[k]);                         // every lane writes the same value
        if (EDGE) {
            // hand the path over when it leaves the strip (it continues in the neighbour's column next row)
            if (dx > 0 && xl == Wk - 1 && th.inbox_right) {
                uint8_t* dst = th.inbox_right + ((size_t)par * 2 + 0) * (D + 16);
                st_row<NREG>(dst + lane * NB, 0, pw[k]);
                if (lane == 0) *reinterpret_cast<uint32_t*>(dst + D) = Mn[k];
            }
            if (dx < 0 && xl == 0 && th.inbox_left) {
                uint8_t* dst = th.inbox_left + ((size_t)par * 2 + 1) * (D + 16);
                st_row<NREG>(dst + lane * NB, 0, pw[k]);
                if (lane == 0) *reinterpret_cast<uint32_t*>(dst + D) = Mn[k];
            }
        }
    }
    // the row's hand-overs are written: publish them before this pixel's global stores are issued (an arrive.release
    // after those would wait for them to reach L2)
    if (EDGE && arrive) cluster_arrive();
    const bool hasA = !FAST && th.addA_l != nullptr, hasB = !FAST && th.addB_l != nullptr;
    if (FAST) {
        if (!FINAL) {
            // one byte per voxel: sum_k (L_k - C), byte-wise on the packed words (each byte of the result is in [0, NDIR*P2];
            // borrows and carries of the partial sums cancel in the 32-bit arithmetic)
            uint32_t sb[NW];
#pragma unroll
            for (int i = 0; i < NW; ++i) {
                sb[i] = pw[0][i] - cw[i];
#pragma unroll
                for (int k = 1; k < NDIR; ++k) sb[i] += pw[k][i] - cw[i];
            }
            st_row<NREG>(reinterpret_cast<uint8_t*>(th.Sout_l) + vox, 0, sb);      // Sout_l carries a BYTE lane offset here
            return;
        }
        if (DEFER) {
            // the caller runs vs_tail_fast for this pixel after the next pixel's direction steps
#pragma unroll
            for (int i = 0; i < NREG; ++i) acc_out[i] = acc[i];
            return;
        }
        vs_tail_fast<NREG>(th, acc, g, pix);
        return;
    } else if (hasA && hasB && th.preadd) {
        // both horizontal rows present and their byte-wise sum cannot carry (2*(cmax+P2) <= 255): add first, unpack once
        uint32_t ab[NW], t[NREG];
#pragma unroll
        for (int i = 0; i < NW; ++i) ab[i] = g.a[i] + g.b[i];
        unpack_cost<NREG>(ab, t);
#pragma unroll
        for (int i = 0; i < NREG; ++i) acc[i] += t[i];
    } else {
        if (hasA) { uint32_t t[NREG]; unpack_cost<NREG>(g.a, t);
#pragma unroll
            for (int i = 0; i < NREG; ++i) acc[i] += t[i]; }
        if (hasB) { uint32_t t[NREG]; unpack_cost<NREG>(g.b, t);
#pragma unroll
            for (int i = 0; i < NREG; ++i) acc[i] += t[i]; }
    }
    if (!FINAL) {
        uint32_t* sp = reinterpret_cast<uint32_t*>(th.Sout_l + vox);
        if (NREG == 4) *reinterpret_cast<uint4*>(sp) = make_uint4(acc[0], acc[1], acc[2], acc[3]);
        else if (NREG == 2) *reinterpret_cast<uint2*>(sp) = make_uint2(acc[0], acc[1]);
        else *sp = acc[0];
    } else {
        if (!FAST && th.Sin_l) {
#pragma unroll
            for (int i = 0; i < NREG; ++i) acc[i] += g.s[i];
        }
        if (!FAST && th.Sout_l) {
            uint32_t* sp = reinterpret_cast<uint32_t*>(th.Sout_l + vox);
#pragma unroll
            for (int i = 0; i < NREG; ++i) sp[i] = acc[i];
        }
        vs_wta<NREG, FAST>(th, acc, pix);
    }
}

// NDIR: 1 (vertical only, the reference's 4-path setting) or 3.  FINAL: add Sin and do WTA instead of writing Sout.
// (Capping the registers at 72 so that a front-end CTA of the next wave could share the SM was measured: the cap costs the
// cluster kernel 8 % and the co-resident kernels give nothing back — 1549 -> 1475 pairs/s.)
template <int NREG, int NDIR, bool FINAL, bool FAST>
__global__ void __launch_bounds__(VS_WARPS * 32, 1)
vsweep_kernel(const VsParams prm)
{
    constexpr int D = 64 * NREG, NW = (NREG + 1) / 2;
    cg::cluster_group cluster = cg::this_cluster();
    const int CS = (int)cluster.num_blocks(), rank = (int)cluster.block_rank();
    const int pair = blockIdx.x / CS;
    const int W = prm.W, H = prm.H, Wk_max = prm.Wk;
    const int xb = rank * Wk_max, Wk = min(Wk_max, W - xb);
    const size_t N = (size_t)W * H;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;

    extern __shared__ __align__(128) unsigned char vs_smem[];
    uint8_t* cbuf = vs_smem;                                               // [2][Wk_max*D]
    uint8_t* state = cbuf + 2 * (size_t)Wk_max * D;                        // [NDIR][Wk_max][D]
    uint32_t* stmin = reinterpret_cast<uint32_t*>(state + (size_t)NDIR * Wk_max * D);   // [NDIR][Wk_max]
    uint8_t* inbox = reinterpret_cast<uint8_t*>(stmin) + (((size_t)NDIR * Wk_max * 4 + 15) & ~(size_t)15);   // [2][2][D + 16]  (row parity; 0: from left, 1: from right)
    uint16_t* wsc = reinterpret_cast<uint16_t*>(inbox + 4 * (D + 16));     // FINAL: [VS_WARPS][D] per-warp scratch
    uint64_t* bars = reinterpret_cast<uint64_t*>(reinterpret_cast<uint8_t*>(wsc) + (FINAL ? VS_WARPS * D * 2 : 0));

    const uint8_t* Cb = prm.C + pair * N * D;
    const uint32_t row_bytes = (uint32_t)Wk * D;
    auto row_y = [&](int yy) { return prm.up ? H - 1 - yy : yy; };

    if (threadIdx.x == 0) {
        mbar_init(&bars[0], 1);
        mbar_init(&bars[1], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    cluster.sync();
    if (threadIdx.x == 0) {
        for (int yy = 0; yy < min(2, H); ++yy) {
            mbar_expect_tx(&bars[yy], row_bytes);
            tma_load_1d(cbuf + (size_t)yy * Wk_max * D, Cb + ((size_t)row_y(yy) * W + xb) * D, row_bytes, &bars[yy]);
        }
    }

    // neighbours' inboxes (distributed shared memory)
    uint8_t* inbox_right = (rank + 1 < CS) ? cluster.map_shared_rank(inbox, rank + 1) : nullptr;   // we write slot "from left" there
    uint8_t* inbox_left = (rank > 0) ? cluster.map_shared_rank(inbox, rank - 1) : nullptr;        // we write slot "from right" there

    const uint32_t P1P1 = h2_const(prm.P1), P2P2 = h2_const(prm.P2);          // fp16x2 constants (sgm_step.cuh)
    // per-thread bases with the lane's byte offset folded in
    constexpr int NB = 2 * NREG;
    VsThread<NREG> th;
    const uint32_t cbuf_s = smem_u32(cbuf) + lane * NB;
    th.state_s = smem_u32(state) + lane * NB; th.stmin_s = smem_u32(stmin); th.inbox_s = smem_u32(inbox);
    th.inbox_right = inbox_right; th.inbox_left = inbox_left;
    th.addA_l = prm.addA ? prm.addA + pair * N * D + lane * NB : nullptr;
    th.addB_l = prm.addB ? prm.addB + pair * N * D + lane * NB : nullptr;
    if (FAST) {      // byte volumes behind the u16 pointers: offsets in bytes
        th.Sin_l = prm.Sin ? reinterpret_cast<const uint16_t*>(reinterpret_cast<const uint8_t*>(prm.Sin) + pair * N * D + lane * NB) : nullptr;
        th.Sout_l = prm.Sout ? reinterpret_cast<uint16_t*>(reinterpret_cast<uint8_t*>(prm.Sout) + pair * N * D + lane * NB) : nullptr;
    } else {
        th.Sin_l = prm.Sin ? prm.Sin + pair * N * D + lane * NB : nullptr;
        th.Sout_l = prm.Sout ? prm.Sout + pair * N * D + lane * NB : nullptr;
    }
    th.minC = prm.minC ? prm.minC + pair * N : nullptr;
    th.rec = prm.rec ? prm.rec + pair * N * 4 : nullptr;
    th.ws = wsc + warp * D;
    th.Wk = Wk; th.Wk_max = Wk_max; th.W = W; th.xb = xb; th.lane = lane;
    th.P1P1 = P1P1; th.P2P2 = P2P2; th.sel_lo = h2_edge_sel_lo(lane); th.sel_hi = h2_edge_sel_hi(lane);
    th.preadd = 2 * (24 + prm.P2) <= 255;          // cost values are <= 24 on this path (no-wrap domain precondition)

    // Pixel order inside a row is free (every path slot is touched by exactly one pixel per row), so the strip is split in
    // two halves walked from the outside in: warps 0..HW-1 take xl = 0, HW, 2*HW, ... of the lower half, warps HW..2*HW-1 take
    // xl = Wk-1, Wk-1-HW, ... of the upper half.  The two pixels that exchange paths with the neighbour CTAs (xl = 0 and
    // xl = Wk-1) are then the FIRST pixels of warps 0 and HW, and the cluster barrier is split around them: every warp
    // arrives right after its first pixel of the row and waits at the start of the next row, so the hand-over of row yy is
    // ordered (arrive.release / wait.acquire) while the rest of the row overlaps the neighbours' progress.  What remains
    // per row inside the CTA is one bar.sync (state slots move between warps from row to row; the cost buffer is reused).
    constexpr int HW = VS_WARPS / 2;
    const int wsub = warp % HW;
    const bool upper = warp >= HW;
    const int n_lo = (Wk + 1) / 2, cnt = upper ? Wk - n_lo : n_lo;
    const int xfirst = upper ? Wk - 1 : 0, xstep = upper ? -1 : 1;
    // Warp w walks the pixels i = w, w + HW, w + 2*HW, ... of its half.  The warp that owns the edge pixel (w = 0) also pays for
    // the hand-over and the releasing cluster arrive, and every other warp waits for it at the row-end barrier; when the half
    // does not divide evenly (KITTI: 69 pixels over 10 warps) its last pixel is therefore given to the first warp that would
    // otherwise have one pixel less.  `lim` is the exclusive bound of the warp's natural sequence, `imax` clamps the index of
    // the receiving warp's extra step onto the moved pixel.
    int lim = cnt, imax = 0x7FFFFFFF;
#if FSGM_VS_RELIEF
    {
        const int full = cnt / HW, rem = cnt % HW;
        if (full >= 1 && rem > 0) {
            if (wsub == 0) lim = full * HW;                              // drops i = full*HW
            if (wsub == rem) { lim = cnt + HW; imax = full * HW; }       // one more step, clamped onto that pixel
        }
    }
#endif
    auto xl_of = [&](int i) { return xfirst + xstep * min(i, imax); };

    int off = 0;                                   // yy mod Wk, kept incrementally
    constexpr int PD = FINAL ? FSGM_VS_PDF : 4;
    VsGlobals<NREG> gq[PD];
    for (int yy = 0; yy < H; ++yy) {
        const int y = row_y(yy), par = yy & 1;
        if (yy > 0) cluster_wait();                // hand-overs of row yy-1 are visible; neighbours are done reading inbox[par]
        mbar_wait(&bars[par], (uint32_t)((yy >> 1) & 1));
        const uint32_t crow_l = cbuf_s + (uint32_t)(par * Wk_max * D);
        const uint32_t rowpix = (uint32_t)y * (uint32_t)W + (uint32_t)xb;
        // global rows (horizontal volumes / the other pass's sum) are fetched PD pixels ahead: with only 20 warps per SM an
        // L2 miss (~2 us) is not hidden by other warps (ncu r1g: 27 % of the stall samples of the down pass sat on the first
        // use of the horizontal-volume row).  The ring is indexed statically (inner loop unrolled by PD): copying a register
        // that is still waiting for its load would stall on the copy, which is exactly what a rotating ring does.
        // All warps start a row together behind the barrier and wait for their first pixel's rows (ncu r1p: 15 % of the final
        // pass's stall samples), yet every way of requesting those rows EARLIER was measured slower: the next row's first pixels
        // before the row-end barrier (21.9 -> 23.3 ms per 60 pairs), the whole ring carried across the barrier (each slot
        // refilled with its next-row pixel after its last use: 1837 -> 1814 pairs/s, the same with ld.global.cg), the first
        // pixels' rows staged by bulk copies next to the cost row two rows ahead (1821 -> 1776); an L2 prefetch changes nothing.
        // The row-synchronous bursts are what DRAM serves best here.
#pragma unroll
        for (int u = 0; u < PD; ++u)
            if (wsub + u * HW < lim) vs_fetch<NREG, FINAL, FAST>(th, rowpix + xl_of(wsub + u * HW), gq[u]);
        // warps 0 and HW publish the row's hand-overs from inside their first (edge) pixel; every other warp has nothing to
        // publish and arrives right away
        if (wsub != 0 || lim == 0) cluster_arrive_relaxed();
        if (yy == 0) {
            // first row: every path starts here (edge body for every pixel); one row, so no software pipelining
            for (int i = wsub; i < lim; i += HW) {
                const int xl = xl_of(i);
                if (i != wsub) vs_fetch<NREG, FINAL, FAST>(th, rowpix + xl, gq[0]);
                vs_pixel<NREG, NDIR, FINAL, true, FAST>(th, crow_l, xl, yy, par, off, rowpix + xl, gq[0], i == 0);
            }
        } else {
            // the only edge pixels of a later row are xl = 0 and xl = Wk-1: pixel i = 0 of warps 0 and HW, i.e. ring slot 0 of
            // their first round
            constexpr bool DEFER = FAST && FINAL && FSGM_VS_DEFER;
            uint32_t accp[NREG];                       // DEFER: sums of the previous pixel, whose tail runs after this pixel's steps
            uint32_t pixp = 0;
            for (int i0 = wsub; i0 < lim; i0 += PD * HW) {
#pragma unroll
                for (int u = 0; u < PD; ++u) {
                    const int i = i0 + u * HW;
                    if (i < lim) {
                        const int xl = xl_of(i);
                        if (DEFER) {
                            uint32_t accn[NREG];
                            if (u == 0 && i == 0) vs_pixel<NREG, NDIR, FINAL, true, FAST, true>(th, crow_l, xl, yy, par, off, rowpix + xl, gq[u], true, accn);
                            else vs_pixel<NREG, NDIR, FINAL, false, FAST, true>(th, crow_l, xl, yy, par, off, rowpix + xl, gq[u], false, accn);
                            const int up = (u + PD - 1) % PD;                  // ring slot of the previous pixel (a constant once unrolled)
                            if (i != wsub) {
                                vs_tail_fast<NREG>(th, accp, gq[up], pixp);
                                if (i - HW + PD * HW < lim) vs_fetch<NREG, FINAL, FAST>(th, rowpix + xl_of(i - HW + PD * HW), gq[up]);
                            }
                            if (i + HW >= lim) vs_tail_fast<NREG>(th, accn, gq[u], rowpix + xl);     // the warp's last pixel of the row: nothing to hide behind
#pragma unroll
                            for (int r = 0; r < NREG; ++r) accp[r] = accn[r];
                            pixp = rowpix + xl;
                        } else {
                            if (u == 0 && i == 0) vs_pixel<NREG, NDIR, FINAL, true, FAST>(th, crow_l, xl, yy, par, off, rowpix + xl, gq[u], true);
                            else vs_pixel<NREG, NDIR, FINAL, false, FAST>(th, crow_l, xl, yy, par, off, rowpix + xl, gq[u], false);
                            if (i + PD * HW < lim) vs_fetch<NREG, FINAL, FAST>(th, rowpix + xl_of(i + PD * HW), gq[u]);
                        }
                    }
                }
            }
        }
        if (++off == Wk) off = 0;
        // everyone in this CTA is done with the row: cost buffer `par` is free, state slots may change hands
        __syncthreads();
        if (threadIdx.x == 0 && yy + 2 < H) {
            mbar_expect_tx(&bars[par], row_bytes);
            tma_load_1d(cbuf + (size_t)par * Wk_max * D, Cb + ((size_t)row_y(yy + 2) * W + xb) * D, row_bytes, &bars[par]);
        }
    }
    cluster_wait();                                // nobody writes into this CTA's shared memory after this point
}

// subpixel + vz -> disparity from the WTA records (calc_cost_sgm.cpp:278-308, :414-426); the value the reference reads
// for argmin == D-1 is the NEXT pixel's Sp[0] (0 past the last pixel)
__device__ __forceinline__ uint32_t vs_x86_d2u(double v)
{
    if (!(v > -9223372036854775809.0 && v < 9223372036854775808.0)) return 0u;
    return (uint32_t)(unsigned long long)__double2ll_rz(v);
}

__global__ void vs_finalize_kernel(const uint16_t* __restrict__ rec, const uint32_t* __restrict__ minC, const double* __restrict__ O,
                                   size_t N, int D, int subpixel, int vz_to_disp, double vMax, uint32_t bias, uint32_t* __restrict__ bestD)
{
    const size_t p = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= N) return;
    const size_t gp = blockIdx.y * N + p;
    const uint2 r = *reinterpret_cast<const uint2*>(rec + gp * 4);
    const uint32_t idx = r.x & 0xFFFFu;
    uint32_t q = idx;
    if (subpixel) {
        if (idx > 1) {
            // `bias`: the FAST cluster passes record the neighbours as biased fp16 bit patterns (0x6400 | value)
            const double c_1 = (double)((r.x >> 16) - bias), c = (double)minC[gp];
            double c1 = (double)((r.y & 0xFFFFu) - bias);
            if (idx + 1 >= (uint32_t)D) c1 = (p + 1 < N) ? (double)(rec[(gp + 1) * 4 + 3] - bias) : 0.0;
            const double num = __dsub_rn(c1, c_1);
            const double den = (c1 < c_1) ? __dsub_rn(c, c_1) : __dsub_rn(c, c1);
            q = vs_x86_d2u(__dmul_rn(__dadd_rn((double)idx, __ddiv_rn(__ddiv_rn(num, den), 2.0)), 256.0));
        } else q = idx * 256u;
    }
    if (vz_to_disp) {
        const double d = __ddiv_rn((double)q, 256.0);
        const double rr = __dmul_rn(__ddiv_rn(d, (double)(D + 1)), vMax);
        const double vz = __ddiv_rn(rr, __dsub_rn(1.0, rr));
        q = vs_x86_d2u(__dmul_rn(__dmul_rn(O[gp], vz), 256.0));
    }
    bestD[gp] = q;
}

// ---- host side ------------------------------------------------------------------------------------------
static size_t vs_smem_bytes(int D, int Wk, int ndir, bool final_)
{
    return 2 * (size_t)Wk * D + (size_t)ndir * Wk * D + (((size_t)ndir * Wk * 4 + 15) & ~(size_t)15) + 4 * (size_t)(D + 16) +
           (final_ ? (size_t)VS_WARPS * D * 2 : 0) + 64;
}

size_t vsweep_smem_bytes(int D, int Wk, int ndir) { return vs_smem_bytes(D, Wk, ndir, true); }
int vsweep_threads() { return VS_WARPS * 32; }

static bool vs_cluster_fits(int W, int D, int ndir, int cs, int max_smem)
{
    const int Wk = (W + cs - 1) / cs;
    if (Wk < 2 || (cs - 1) * Wk >= W) return false;                        // every CTA needs at least one column ...
    if (cs > 1 && W - (cs - 1) * Wk < 2) return false;                     // ... and the last one two
    return vs_smem_bytes(D, Wk, ndir, true) <= (size_t)max_smem;
}

// smallest cluster size (1..16) for which the state fits, or 0
int vsweep_cluster_size(int W, int D, int ndir, int max_smem)
{
    if (D != 64 && D != 128 && D != 256) return 0;
    for (int cs = 1; cs <= 16; ++cs)
        if (vs_cluster_fits(W, D, ndir, cs, max_smem)) return cs;
    return 0;
}
bool vsweep_cluster_ok(int W, int D, int ndir, int cs, int max_smem)
{
    return (D == 64 || D == 128 || D == 256) && cs >= 1 && cs <= 16 && vs_cluster_fits(W, D, ndir, cs, max_smem);
}

// The cluster size that finishes a full wave of pairs soonest: a pair's pass takes (rows x pixel rounds per row), a round
// being one pixel per warp (each half of the strip is walked by VS_WARPS/2 warps); resident clusters run concurrently.
// Sizes above the smallest that fits are worth it when they keep more SMs busy (KITTI: 9 x 15 clusters = 135 SMs with 7
// rounds per row against 8 x 15 = 120 SMs with 8 rounds).  *clusters receives the resident-cluster count of the choice.
int vsweep_best_cluster(int W, int D, int ndir, int max_smem, int* clusters)
{
    const int cs0 = vsweep_cluster_size(W, D, ndir, max_smem);
    *clusters = 1;
    if (!cs0) return 0;
    int best = 0; double best_score = 0;
    for (int cs = cs0; cs <= std::min(16, cs0 + 3); ++cs) {
        if (!vs_cluster_fits(W, D, ndir, cs, max_smem)) continue;
        const int Wk = (W + cs - 1) / cs;
        const int k = vsweep_max_clusters(cs, vs_smem_bytes(D, Wk, ndir, true), VS_WARPS * 32);
        if (k < 1) continue;
        const int rounds = ((Wk + 1) / 2 + VS_WARPS / 2 - 1) / (VS_WARPS / 2);
        const double score = (double)rounds / k;                           // time per pair, arbitrary units
        if (!best || score < best_score * 0.97) { best = cs; best_score = score; *clusters = k; }
    }
    return best;
}

template <int NREG, int NDIR, bool FINAL, bool FAST>
static int vs_launch_t(fsgm_ctx* c, int n, int cs, size_t smem, const VsParams& p)
{
    auto kern = vsweep_kernel<NREG, NDIR, FINAL, FAST>;
    const unsigned bit = 1u << (3 + ((NREG == 1 ? 0 : NREG == 2 ? 1 : 2) * 4 + (NDIR == 3 ? 2 : 0) + (FINAL ? 1 : 0)) * 2 + (FAST ? 1 : 0));
    if (!(c->attr_mask & bit)) {
        FSGM_CUDA(c, cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
        FSGM_CUDA(c, cudaFuncSetAttribute(kern, cudaFuncAttributeNonPortableClusterSizeAllowed, 1));
        c->attr_mask |= bit;
    }
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(cs * n); cfg.blockDim = dim3(VS_WARPS * 32); cfg.dynamicSmemBytes = smem; cfg.stream = c->stream;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = cs; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
    cfg.attrs = at; cfg.numAttrs = 1;
    FSGM_CUDA(c, cudaLaunchKernelEx(&cfg, kern, p));
    c->launches++;
    return FSGM_OK;
}

// one pass (down or up) over n pairs; see VsParams
// fast = the FAST operand configuration (vsweep_fast_ok() must hold): the first pass takes no addA / addB / Sin and writes a
// BYTE volume through Sout; the final pass takes addA, addB and that byte volume as Sin, no Sout; its WTA records are biased
// (launch_vs_finalize takes the same flag).
bool vsweep_fast_ok(int ndir, int P2) { return ndir * P2 <= 255 && 2 * (24 + P2) <= 255 && 8 * (24 + P2) < 1024; }

int launch_vsweep(fsgm_ctx* c, int n, int cs, int ndir, bool final_, const uint8_t* C, const uint8_t* addA, const uint8_t* addB,
                  const uint16_t* Sin, uint16_t* Sout, uint32_t* minC, uint16_t* rec, int W, int H, int D, int P1, int P2, int up,
                  bool fast)
{
    StageScope ss(c, ST_VSWEEP);
    VsParams p{};
    p.C = C; p.addA = addA; p.addB = addB; p.Sin = Sin; p.Sout = Sout; p.minC = minC; p.rec = rec;
    p.W = W; p.H = H; p.Wk = (W + cs - 1) / cs; p.P1 = P1; p.P2 = P2; p.up = up;
    const size_t smem = vs_smem_bytes(D, p.Wk, ndir, final_);
    const int nreg = D / 64;
    // compile-time operand configuration of the standard two-pass run (see vs_fetch)
    if (fast) {
        const bool ok = vsweep_fast_ok(ndir, P2) && (final_ ? (addA && addB && Sin && !Sout) : (!addA && !addB && !Sin && Sout));
        if (!ok) return fail(c, FSGM_ERR_DOMAIN, "vsweep: operands do not match the FAST configuration");
    }
    p.fast = fast;
#define VS_GO(NR, ND, FN) do { if (fast) return vs_launch_t<NR, ND, FN, true>(c, n, cs, smem, p); return vs_launch_t<NR, ND, FN, false>(c, n, cs, smem, p); } while (0)
    if (ndir == 3) {
        if (final_) { if (nreg == 4) VS_GO(4, 3, true); if (nreg == 2) VS_GO(2, 3, true); VS_GO(1, 3, true); }
        else        { if (nreg == 4) VS_GO(4, 3, false); if (nreg == 2) VS_GO(2, 3, false); VS_GO(1, 3, false); }
    } else {
        if (final_) { if (nreg == 4) VS_GO(4, 1, true); if (nreg == 2) VS_GO(2, 1, true); VS_GO(1, 1, true); }
        else        { if (nreg == 4) VS_GO(4, 1, false); if (nreg == 2) VS_GO(2, 1, false); VS_GO(1, 1, false); }
    }
#undef VS_GO
}

// occupancy probe used by the tuning notes in DESIGN.md: how many clusters of `cs` CTAs with `smem` bytes can be resident
int vsweep_max_clusters(int cs, size_t smem, int threads)
{
    auto kern = vsweep_kernel<4, 3, true, true>;
    cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    cudaFuncSetAttribute(kern, cudaFuncAttributeNonPortableClusterSizeAllowed, 1);
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(cs * 64); cfg.blockDim = dim3(threads); cfg.dynamicSmemBytes = smem;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = cs; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
    cfg.attrs = at; cfg.numAttrs = 1;
    int n = -1;
    if (cudaOccupancyMaxActiveClusters(&n, kern, &cfg) != cudaSuccess) { cudaGetLastError(); return -1; }
    return n;
}

int launch_vs_finalize(fsgm_ctx* c, int n, const uint16_t* rec, const uint32_t* minC, const double* O, int W, int H, int D,
                       int subpixel, int vz_to_disp, double vMax, int biased, uint32_t* bestD)
{
    StageScope ss(c, ST_WTA);
    const size_t N = (size_t)W * H;
    dim3 grid((unsigned)((N + 255) / 256), n);
    vs_finalize_kernel<<<grid, 256, 0, c->stream>>>(rec, minC, O, N, D, subpixel, vz_to_disp, vMax, biased ? (H2_BIAS2 & 0xFFFFu) : 0u, bestD);
    FSGM_LAUNCHED(c);
    return FSGM_OK;
}

}  // namespace fsgm
