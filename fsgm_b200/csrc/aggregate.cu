// Path aggregation — replaces the sweeps of sgm() (reference calc_cost_sgm.cpp:114-257, step :33-66).
//
//   L_r(p,d) = C(p,d) + min( L_r(p-r,d), L_r(p-r,d-1)+P1, L_r(p-r,d+1)+P1, min_k L_r(p-r,k)+P2 ) - min_k L_r(p-r,k)
//
// with the reference's path-start rule: where p-r is outside the image, L = C and the stored minimum
// is 0, not min(C) (:152-180).  The reference advances four ring-buffered paths per pixel in raster
// order; the paths never read each other, so here every direction is an independent set of scanlines.
//
// Mapping: one warp per scanline, labels spread across lanes (2*NREG consecutive labels per lane, kept
// as u16x2 in NREG registers).  Neighbour labels come from two warp shuffles, min_k from a packed min
// tree + one REDUX.MIN.  The next pixels' cost rows are prefetched PF steps ahead into a register ring
// so the serial chain never waits on HBM.  Horizontal directions walk a row (contiguous 256-B steps);
// the six others walk "wrapped" columns: x advances by dx every row and the path restarts when it wraps,
// which gives every warp exactly H steps and keeps neighbouring warps on neighbouring 256-B rows.
//
// HBM layout: C and every L_r are u8 [pair][y][x][d] (label-contiguous): 1 B read + 1 B written per voxel
// and direction.  u16 exists only in registers.
//
// Arithmetic: the reference computes in unsigned char with mod-256 truncation (common.h:4-8).  When
// P1,P2 >= 0, cmax+P1+P2 <= 255 and 2*cmax+P2 <= 255 no truncation can fire and the u16 lanes are exact;
// otherwise the WRAP=true instantiation of sweep_kernel reproduces every truncation explicitly.
//
// Kernels in this file (launch_sweeps picks one):
//   hsweep_tma_kernel  horizontal directions only, no-wrap domain, D in {64,128,256}: cost rows staged by 1-D bulk TMA
//   sweep_fast_kernel  any set of directions, no-wrap domain: register prefetch ring, branch-free path restarts
//   sweep_kernel       WRAP=true: explicit mod-256 emulation (parameters or cost values outside the no-wrap domain)
// The three non-horizontal directions of a pass have a faster, row-synchronous implementation in vsweep.cu.
#include "fsgm_internal.h"
#include "sgm_step.cuh"
#include <type_traits>

namespace fsgm {

constexpr int SWEEP_WARPS = 8;       // warps per CTA
constexpr int PF = 4;                // prefetch depth (steps)
constexpr uint32_t BIG2 = 0x3F003F00u;

enum LoadMode { LM_FULL = 0, LM_VECPAD = 1, LM_BYTES = 2 };

struct SweepParams {
    const uint8_t* C;
    const uint8_t* I1;
    uint8_t* L[8];
    int dir[8];
    int line_start[9];    // prefix sum of scanline counts over the enabled directions
    int n_dirs;
    int W, H, D;
    int P1, P2, adaptive_thr;
    // SCATTER (direction split over GPUs, dist.cu): a pixel's L row goes straight into the memory of the rank that owns the pixel's
    // STRIPE — stripes of a few image rows are dealt to the ranks round-robin, so that the down / up sweeps of all ranks, which
    // move through the image rows together, spread their stores over every owner at any time instead of all hitting the one
    // rank that owns the current part of the image (NVLink ingress of one GPU).  peer[j] is rank j's receive buffer (peer-mapped
    // over NVLink, or local for j == this rank), laid out [direction slot][local stripe][chunk_pixels][D] with
    // chunk_pixels = stripe_pixels + 1: the extra pixel holds a copy of the first pixel of the NEXT stripe (the reference's read of
    // the next pixel's label 0, calc_cost_sgm.cpp:293-296).  Direction k of this launch writes slot slot[k].
    uint8_t* peer[16];
    int slot[8];
    unsigned stripe_pixels, chunk_pixels, local_pixels;
    int world;
};

template <int NREG> struct Words { uint32_t w[(NREG + 1) / 2]; };

template <int NREG, int MODE>
__device__ __forceinline__ Words<NREG> load_row(const uint8_t* __restrict__ pix, int lane, int D)
{
    Words<NREG> r;
    constexpr int NB = 2 * NREG;
    if (MODE == LM_BYTES) {
#pragma unroll
        for (int k = 0; k < (NREG + 1) / 2; ++k) {
            uint32_t v = 0;
#pragma unroll
            for (int j = 0; j < 4 && k * 4 + j < NB; ++j) {
                int d = lane * NB + k * 4 + j;
                if (d < D) v |= (uint32_t)__ldg(pix + d) << (8 * j);
            }
            r.w[k] = v;
        }
    } else {
        const bool alive = (MODE == LM_FULL) || (lane * NB < D);
        if (NREG == 1) { r.w[0] = alive ? __ldg(reinterpret_cast<const uint16_t*>(pix) + lane) : 0; }
        else if (NREG == 2) { r.w[0] = alive ? __ldg(reinterpret_cast<const uint32_t*>(pix) + lane) : 0; }
        else if (NREG == 4) {
            uint2 v = alive ? __ldg(reinterpret_cast<const uint2*>(pix) + lane) : make_uint2(0, 0);
            r.w[0] = v.x; r.w[1] = v.y;
        } else {
            uint4 v = alive ? __ldg(reinterpret_cast<const uint4*>(pix) + lane) : make_uint4(0, 0, 0, 0);
            r.w[0] = v.x; r.w[1] = v.y; r.w[2] = v.z; r.w[3] = v.w;
        }
    }
    return r;
}

template <int NREG, int MODE>
__device__ __forceinline__ void store_row(uint8_t* __restrict__ pix, int lane, int D, const uint32_t (&L)[NREG])
{
    constexpr int NB = 2 * NREG;
    uint32_t w[(NREG + 1) / 2];
#pragma unroll
    for (int k = 0; k < (NREG + 1) / 2; ++k)
        w[k] = (2 * k + 1 < NREG) ? __byte_perm(L[2 * k], L[2 * k + 1], 0x6420) : __byte_perm(L[2 * k], 0, 0x6420);
    if (MODE == LM_BYTES) {
#pragma unroll
        for (int j = 0; j < NB; ++j) {
            int d = lane * NB + j;
            if (d < D) pix[d] = (uint8_t)(w[j >> 2] >> (8 * (j & 3)));
        }
    } else {
        const bool alive = (MODE == LM_FULL) || (lane * NB < D);
        if (!alive) return;
        if (NREG == 1) reinterpret_cast<uint16_t*>(pix)[lane] = (uint16_t)w[0];
        else if (NREG == 2) reinterpret_cast<uint32_t*>(pix)[lane] = w[0];
        else if (NREG == 4) reinterpret_cast<uint2*>(pix)[lane] = make_uint2(w[0], w[1]);
        else reinterpret_cast<uint4*>(pix)[lane] = make_uint4(w[0], w[1], w[2], w[3]);
    }
}


// ------------------------------------------------------------------------------------------------------------
// No-wrap fast kernel.  Same mapping as sweep_kernel below, restructured so the per-step overhead is small:
//   * the scanline is tracked as a 32-bit pixel index; addresses are one IMAD.WIDE each (pixel * D + base);
//   * the pixel indices of the prefetched rows ride along in the register ring, so the store needs no second cursor;
//   * a path restart is branch-free: the previous state is replaced by zeros with M = 0, for which the step formula
//     yields L = C, and the new minimum is forced to 0 afterwards (the reference's path-start rule, :152-180).
// ------------------------------------------------------------------------------------------------------------
template <int NREG, int MODE, bool ADAPT, bool SCATTER = false>
__global__ void __launch_bounds__(SWEEP_WARPS * 32)
sweep_fast_kernel(const SweepParams prm)
{
    const int lane = threadIdx.x & 31;
    const int gw = blockIdx.x * SWEEP_WARPS + (threadIdx.x >> 5);
    if (gw >= prm.line_start[prm.n_dirs]) return;
    int k = 0;
    while (gw >= prm.line_start[k + 1]) ++k;
    const int line = gw - prm.line_start[k];
    const int r = prm.dir[k];
    const int dx = dir_dx(r), dy = dir_dy(r);
    const int W = prm.W, H = prm.H, D = prm.D;
    const size_t N = (size_t)W * H;
    constexpr int NB = 2 * NREG;
    const uint8_t* __restrict__ Cb = prm.C + blockIdx.y * N * D;
    uint8_t* __restrict__ Lb = prm.L[k] + blockIdx.y * N * D;
    const uint8_t* __restrict__ Ib = ADAPT ? prm.I1 + blockIdx.y * N : nullptr;

    // scanline as pixel indices: p(t+1) = p(t) + dp, plus `fix` whenever x wraps (every W steps after the first `wrap_in`)
    int p, len, wrap_in;
    const int dp = dy * W + dx, fix = -dx * W;
    if (dy == 0) { p = line * W + (dx > 0 ? 0 : W - 1); len = W; wrap_in = 0x7FFFFFFF; }
    else {
        p = (dy > 0 ? 0 : H - 1) * W + line; len = H;
        wrap_in = dx == 0 ? 0x7FFFFFFF : (dx > 0 ? W - line : line + 1);
    }
    int pf = p, pf_wrap = wrap_in;                                   // prefetch cursor
    int restart_in = wrap_in;                                        // steps until the compute cursor sits on a restart pixel

    uint32_t cdead[NREG];
#pragma unroll
    for (int i = 0; i < NREG; ++i) {
        int d0 = lane * NB + 2 * i;
        cdead[i] = (MODE == LM_FULL) ? 0u : (((d0 >= D ? 0xFFFFu : 0u) | (d0 + 1 >= D ? 0xFFFF0000u : 0u)) & STEP_BIG2);
    }
    const uint32_t lo_mask = (lane == 0) ? (STEP_BIG2 & 0x0000FFFFu) : 0u;
    const uint32_t hi_mask = (lane == 31) ? (STEP_BIG2 & 0xFFFF0000u) : 0u;
    const uint32_t P1P1 = (uint32_t)prm.P1 * 0x10001u;
    uint32_t P2P2 = (uint32_t)prm.P2 * 0x10001u;

    auto advance = [&](int& pix, int& cnt) {
        pix += dp;
        if (--cnt == 0) { pix += fix; cnt = W; }
    };

    Words<NREG> ring[PF];
    int pring[PF];
#pragma unroll
    for (int j = 0; j < PF; ++j) {
        pring[j] = pf;
        if (j < len) ring[j] = load_row<NREG, MODE>(Cb + (size_t)(uint32_t)pf * (uint32_t)D, lane, D);
        advance(pf, pf_wrap);
    }

    uint32_t Lr[NREG];
#pragma unroll
    for (int i = 0; i < NREG; ++i) Lr[i] = 0;
    uint32_t M = 0;
    int prev_pix = 0;
    bool restart = true;
    // SCATTER: the stripe [own_lo, own_hi) the cursor is in and the base its rows are written against; the divisions run only
    // when a scanline enters another stripe (horizontal: never)
    unsigned own_lo = 1, own_hi = 0;
    uint8_t* own_base = nullptr;

    for (int t0 = 0; t0 < len; t0 += PF) {
#pragma unroll
        for (int j = 0; j < PF; ++j) {
            const int t = t0 + j;
            if (t >= len) break;
            const Words<NREG> cw = ring[j];
            const int pix = pring[j];
            if (t + PF < len) {
                pring[j] = pf;
                ring[j] = load_row<NREG, MODE>(Cb + (size_t)(uint32_t)pf * (uint32_t)D, lane, D);
            }
            advance(pf, pf_wrap);

            uint32_t c[NREG], cP2[NREG];
            unpack_cost<NREG>(cw.w, c);
            if (ADAPT) {
                int P2 = prm.P2;
                if (!restart && abs((int)Ib[pix] - (int)Ib[prev_pix]) > prm.adaptive_thr) P2 = P2 / 8;
                P2P2 = (uint32_t)P2 * 0x10001u;
            }
#pragma unroll
            for (int i = 0; i < NREG; ++i) {
                if (MODE != LM_FULL) c[i] |= cdead[i];
                cP2[i] = c[i] + P2P2;
            }
            if (restart) {
#pragma unroll
                for (int i = 0; i < NREG; ++i) Lr[i] = 0;
                M = 0;
            }
            uint32_t Ln[NREG];
            const uint32_t m = sgm_step_u16<NREG>(c, cP2, Lr, M, P1P1, lo_mask, hi_mask, Ln);
            M = restart ? 0u : m;
#pragma unroll
            for (int i = 0; i < NREG; ++i) Lr[i] = Ln[i];
            if (SCATTER) {
                const unsigned up = (unsigned)pix;
                if (up < own_lo || up >= own_hi) {
                    const unsigned j = up / prm.stripe_pixels;
                    const unsigned owner = j % (unsigned)prm.world, ls = j / (unsigned)prm.world;
                    own_lo = j * prm.stripe_pixels; own_hi = own_lo + prm.stripe_pixels;
                    own_base = prm.peer[owner] + ((size_t)prm.slot[k] * prm.local_pixels + (size_t)ls * prm.chunk_pixels - own_lo) * (size_t)D;
                }
                store_row<NREG, MODE>(own_base + (size_t)up * (uint32_t)D, lane, D, Lr);
                if (up == own_lo && up != 0) {                      // first pixel of a stripe: a copy behind the previous stripe
                    const unsigned j = up / prm.stripe_pixels - 1;
                    const unsigned owner = j % (unsigned)prm.world, ls = j / (unsigned)prm.world;
                    uint8_t* ex = prm.peer[owner] + ((size_t)prm.slot[k] * prm.local_pixels + (size_t)ls * prm.chunk_pixels + prm.stripe_pixels) * (size_t)D;
                    store_row<NREG, MODE>(ex, lane, D, Lr);
                }
            } else
            store_row<NREG, MODE>(Lb + (size_t)(uint32_t)pix * (uint32_t)D, lane, D, Lr);
            prev_pix = pix;
            restart = (--restart_in == 0);
            if (restart) restart_in = W;
        }
    }
}


// ------------------------------------------------------------------------------------------------------------
// Horizontal sweeps with TMA staging.  Along an image row the cost rows of consecutive steps are contiguous in memory,
// so one elected lane fetches HCH steps (HCH*D bytes) per 1-D bulk copy (cp.async.bulk + mbarrier complete_tx) into a
// per-warp two-chunk shared-memory ring; the serial chain then reads its cost row with one LDS and never waits on HBM
// (the register-prefetch kernel spent most of its stall samples on the first use of the prefetched row: long scoreboard).
// Used for the two horizontal directions next to the row-synchronous cluster kernels (D in {64,128,256}, no-wrap domain).
// ------------------------------------------------------------------------------------------------------------
constexpr int HCH = 8;                       // steps per chunk

__device__ __forceinline__ uint32_t hs_smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

template <int NREG>
__global__ void __launch_bounds__(SWEEP_WARPS * 32)
hsweep_tma_kernel(const SweepParams prm)
{
    constexpr int D = 64 * NREG, NB = 2 * NREG;
    __shared__ __align__(128) uint8_t ring[SWEEP_WARPS][2][HCH * D];
    __shared__ __align__(8) uint64_t bars[SWEEP_WARPS][2];
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    const int gw = blockIdx.x * SWEEP_WARPS + wib;
    const int W = prm.W, H = prm.H;
    const size_t N = (size_t)W * H;
    const bool active = gw < prm.line_start[prm.n_dirs];
    // Both horizontal directions of a launch read the same cost rows: the two sweeps of one image row sit in neighbouring
    // warps, so the second reading of a chunk follows the first within one row sweep instead of a whole volume later
    // (measured: 4.53 -> 4.49 ms per 60 pairs; the reuse distance, ~90 MB of traffic, is still at the edge of the 126 MB L2).
    int k = 0, y = 0;
    if (active) {
        if (prm.n_dirs == 2) { k = gw & 1; y = gw >> 1; }
        else { while (gw >= prm.line_start[k + 1]) ++k; y = gw - prm.line_start[k]; }
    }
    const int dx = dir_dx(prm.dir[k]);                        // +1 or -1 (dy == 0 for every direction of this launch)
    const uint8_t* __restrict__ Crow = prm.C + (blockIdx.y * N + (size_t)y * W) * D;
    uint8_t* __restrict__ Lrow = prm.L[k] + (blockIdx.y * N + (size_t)y * W) * D + lane * NB;
    if (!active) return;

    const int nch = (W + HCH - 1) / HCH;
    // chunk c covers steps [c*HCH, min(W, (c+1)*HCH)); in memory that is a contiguous run of pixels in either direction
    auto issue = [&](int c) {
        const int t0 = c * HCH, cnt = min(HCH, W - t0);
        const int xlo = dx > 0 ? t0 : W - t0 - cnt;
        const uint32_t bytes = (uint32_t)cnt * D;
        uint64_t* bar = &bars[wib][c & 1];
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(hs_smem_u32(bar)), "r"(bytes) : "memory");
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                     ::"r"(hs_smem_u32(ring[wib][c & 1])), "l"(Crow + (size_t)xlo * D), "r"(bytes), "r"(hs_smem_u32(bar)) : "memory");
    };
    if (lane == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(hs_smem_u32(&bars[wib][0])));
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(hs_smem_u32(&bars[wib][1])));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        issue(0);
        if (nch > 1) issue(1);
    }
    __syncwarp();

    // biased fp16x2 arithmetic (sgm_step.cuh): min on the ALU pipe, additions on the FMA pipe
    const uint32_t sel_lo = h2_edge_sel_lo(lane), sel_hi = h2_edge_sel_hi(lane);
    const uint32_t P1h = h2_const(prm.P1), P2h = h2_const(prm.P2);
    constexpr uint32_t ZERO_B = H2_BIAS2;                                // biased 0 in both halves (minima travel duplicated)
    uint32_t Lr[NREG];
#pragma unroll
    for (int i = 0; i < NREG; ++i) Lr[i] = H2_BIAS2;
    uint32_t M = ZERO_B;
    bool first = true;
    for (int c = 0; c < nch; ++c) {
        const int t0 = c * HCH, cnt = min(HCH, W - t0);
        {   // wait for the chunk (phase parity = number of earlier uses of this buffer, mod 2)
            const uint32_t parity = (uint32_t)((c >> 1) & 1), bar = hs_smem_u32(&bars[wib][c & 1]);
            asm volatile("{\n\t.reg .pred p;\n\tHW_%=:\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t@p bra HD_%=;\n\tbra HW_%=;\n\tHD_%=:\n\t}"
                         ::"r"(bar), "r"(parity) : "memory");
        }
        const uint8_t* buf = ring[wib][c & 1] + lane * NB;
        const int xlo = dx > 0 ? t0 : W - t0 - cnt;
        // the chunk is walked with running pointers whose stride sign is a compile-time constant (one loop per direction), so
        // the unrolled steps address the ring and the output row with immediate offsets
        auto walk = [&](auto fwd_tag) {
            constexpr bool FWD = decltype(fwd_tag)::value;
            constexpr int STEP = FWD ? D : -D;
            const uint8_t* bp = buf + (FWD ? 0 : (cnt - 1) * D);
            uint8_t* dp = Lrow + (size_t)(xlo + (FWD ? 0 : cnt - 1)) * D;
#pragma unroll 4
            for (int s = 0; s < cnt; ++s, bp += STEP, dp += STEP) {
                uint32_t cw[(NREG + 1) / 2];
                if (NREG == 1) cw[0] = *reinterpret_cast<const uint16_t*>(bp);
                else if (NREG == 2) cw[0] = *reinterpret_cast<const uint32_t*>(bp);
                else { const uint2 v = *reinterpret_cast<const uint2*>(bp); cw[0] = v.x; cw[1] = v.y; }
                uint32_t cc[NREG], cP2[NREG], Ln[NREG];
                unpack_cost_h2<NREG>(cw, cc);
#pragma unroll
                for (int i = 0; i < NREG; ++i) cP2[i] = h2_add(cc[i], P2h);
                // at the path start the (biased) zero state with M = 0 makes the step return L = C; the minimum is then forced to 0
                const uint32_t m = sgm_step_h2<NREG>(cP2, Lr, M, P1h, P2h, sel_lo, sel_hi, Ln);
                M = first ? ZERO_B : m;
                first = false;
#pragma unroll
                for (int i = 0; i < NREG; ++i) Lr[i] = Ln[i];
                uint32_t pw[(NREG + 1) / 2];
                pack_cost<NREG>(Lr, pw);
                if (NREG == 1) *reinterpret_cast<uint16_t*>(dp) = (uint16_t)pw[0];
                else if (NREG == 2) *reinterpret_cast<uint32_t*>(dp) = pw[0];
                else *reinterpret_cast<uint2*>(dp) = make_uint2(pw[0], pw[1]);
            }
        };
        if (dx > 0) walk(std::true_type{}); else walk(std::false_type{});
        __syncwarp();                                                    // every lane is done reading this buffer
        if (lane == 0 && c + 2 < nch) issue(c + 2);
    }
}

template <int NREG, int MODE, bool WRAP, bool ADAPT>
__global__ void __launch_bounds__(SWEEP_WARPS * 32)
sweep_kernel(const SweepParams prm)
{
    const int lane = threadIdx.x & 31;
    const int gw = blockIdx.x * SWEEP_WARPS + (threadIdx.x >> 5);
    if (gw >= prm.line_start[prm.n_dirs]) return;
    int k = 0;
    while (gw >= prm.line_start[k + 1]) ++k;
    const int line = gw - prm.line_start[k];
    const int r = prm.dir[k];
    const int dx = dir_dx(r), dy = dir_dy(r);
    const int W = prm.W, H = prm.H, D = prm.D;
    const size_t N = (size_t)W * H;
    const uint8_t* __restrict__ Cb = prm.C + blockIdx.y * N * D;
    uint8_t* __restrict__ Lb = prm.L[k] + blockIdx.y * N * D;
    const uint8_t* __restrict__ Ib = ADAPT ? prm.I1 + blockIdx.y * N : nullptr;
    constexpr int NB = 2 * NREG;

    // scanline geometry
    int x, y, len;
    if (dy == 0) { y = line; x = dx > 0 ? 0 : W - 1; len = W; }
    else         { x = line; y = dy > 0 ? 0 : H - 1; len = H; }
    int xf = x, yf = y;                       // prefetch cursor

    // per-lane masks for labels >= D and for the warp's outer neighbours
    uint32_t cdead[NREG];
#pragma unroll
    for (int i = 0; i < NREG; ++i) {
        int d0 = lane * NB + 2 * i;
        cdead[i] = (MODE == LM_FULL) ? 0u : ((d0 >= D ? 0xFFFFu : 0u) | (d0 + 1 >= D ? 0xFFFF0000u : 0u));
    }
    const uint32_t qdead_first = (lane == 0) ? 0x0000FFFFu : 0u;                 // label -1
    const uint32_t qdead_last = (lane == 31) ? 0xFFFF0000u : 0u;                 // label 64*NREG

    const uint32_t P1P1 = WRAP ? ((prm.P1 & 0xFF) * 0x10001u) : (uint32_t)prm.P1 * 0x10001u;

    Words<NREG> ring[PF];
#pragma unroll
    for (int j = 0; j < PF; ++j) {
        if (j < len) ring[j] = load_row<NREG, MODE>(Cb + ((size_t)yf * W + xf) * D, lane, D);
        if (dy == 0) xf += dx;
        else { yf += dy; xf += dx; xf = xf < 0 ? W - 1 : (xf >= W ? 0 : xf); }
    }

    uint32_t Lr[NREG];
    uint32_t M = 0;
    int prev_pix = 0;

    for (int t0 = 0; t0 < len; t0 += PF) {
#pragma unroll
        for (int j = 0; j < PF; ++j) {
            const int t = t0 + j;
            if (t >= len) break;
            const Words<NREG> cw = ring[j];
            if (t + PF < len) ring[j] = load_row<NREG, MODE>(Cb + ((size_t)yf * W + xf) * D, lane, D);
            if (dy == 0) xf += dx;
            else { yf += dy; xf += dx; xf = xf < 0 ? W - 1 : (xf >= W ? 0 : xf); }

            // unpack the cost row to u16x2
            uint32_t c[NREG];
#pragma unroll
            for (int i = 0; i < NREG; ++i) {
                c[i] = __byte_perm(cw.w[i >> 1], 0, (i & 1) ? 0x4342 : 0x4140);
                if (MODE != LM_FULL && !WRAP) c[i] |= cdead[i] & BIG2;
            }
            const int pix = y * W + x;
            // a path (re)starts at the first step and, for the wrapped columns, whenever x just wrapped
            const bool start = (t == 0) || (dy != 0 && dx != 0 && x == (dx > 0 ? 0 : W - 1));
            if (start) {
#pragma unroll
                for (int i = 0; i < NREG; ++i) Lr[i] = c[i];
                M = 0;
            } else {
                int P2 = prm.P2;
                if (ADAPT) {
                    int a = Ib[pix], b = Ib[prev_pix];
                    if (abs(a - b) > prm.adaptive_thr) P2 = P2 / 8;
                }
                uint32_t q[NREG + 1];
                uint32_t up = __shfl_up_sync(0xffffffffu, Lr[NREG - 1], 1);
                uint32_t dn = __shfl_down_sync(0xffffffffu, Lr[0], 1);
                q[0] = __byte_perm(up, Lr[0], 0x5432);
#pragma unroll
                for (int i = 1; i < NREG; ++i) q[i] = __byte_perm(Lr[i - 1], Lr[i], 0x5432);
                q[NREG] = __byte_perm(Lr[NREG - 1], dn, 0x5432);
                uint32_t newL[NREG];
                if (!WRAP) {
                    q[0] |= qdead_first & BIG2;
                    q[NREG] |= qdead_last & BIG2;
                    const uint32_t MM = M * 0x10001u;
                    const uint32_t far2 = MM + (uint32_t)P2 * 0x10001u;
#pragma unroll
                    for (int i = 0; i < NREG; ++i) {
                        uint32_t nb = __vminu2(q[i], q[i + 1]);
                        uint32_t b = __viaddmin_u16x2(nb, P1P1, Lr[i]);        // min(nb + P1, L(p-r,d))
                        b = __vminu2(b, far2);
                        newL[i] = c[i] + b - MM;
                    }
                } else {
                    // explicit mod-256 emulation of every unsigned-char store / std::min<PathCost> in :40-61
                    const uint32_t K = 0x00FF00FFu;
                    const uint32_t MM = M * 0x10001u;
                    const uint32_t far2 = (MM + ((uint32_t)(P2 & 0xFF)) * 0x10001u) & K;
                    uint32_t qm[NREG + 1];
#pragma unroll
                    for (int i = 0; i <= NREG; ++i) {
                        // dead neighbour (label < 0 or >= D): 0xFF is neutral for an unsigned-char min
                        uint32_t dead = 0;
                        int dl = lane * NB + 2 * i - 1;       // label held in the low half of q[i]
                        if (dl < 0 || dl >= D) dead |= 0x000000FFu;
                        if (dl + 1 >= D) dead |= 0x00FF0000u;
                        qm[i] = ((q[i] + P1P1) & K) | dead;
                    }
#pragma unroll
                    for (int i = 0; i < NREG; ++i) {
                        uint32_t b = __vminu2(__vminu2(qm[i], qm[i + 1]), __vminu2(Lr[i], far2));
                        newL[i] = (c[i] + b + (0x01000100u - MM)) & K;
                    }
                }
#pragma unroll
                for (int i = 0; i < NREG; ++i) Lr[i] = newL[i];
                // new minimum over the live labels
                uint32_t m = WRAP ? (Lr[0] | (cdead[0] & 0x00FF00FFu)) : Lr[0];
#pragma unroll
                for (int i = 1; i < NREG; ++i) m = __vminu2(m, WRAP ? (Lr[i] | (cdead[i] & 0x00FF00FFu)) : Lr[i]);
                m = min(m & 0xFFFFu, m >> 16);
                M = __reduce_min_sync(0xffffffffu, m);
            }
            store_row<NREG, MODE>(Lb + (size_t)pix * D, lane, D, Lr);
            prev_pix = pix;
            if (dy == 0) x += dx;
            else { y += dy; x += dx; x = x < 0 ? W - 1 : (x >= W ? 0 : x); }
        }
    }
}

template <int NREG, int MODE>
static void sweep_dispatch2(const SweepParams& p, bool wrap, bool adapt, dim3 grid, cudaStream_t s)
{
    if (!wrap && !adapt) sweep_fast_kernel<NREG, MODE, false><<<grid, SWEEP_WARPS * 32, 0, s>>>(p);
    else if (!wrap && adapt) sweep_fast_kernel<NREG, MODE, true><<<grid, SWEEP_WARPS * 32, 0, s>>>(p);
    else if (wrap && !adapt) sweep_kernel<NREG, MODE, true, false><<<grid, SWEEP_WARPS * 32, 0, s>>>(p);
    else sweep_kernel<NREG, MODE, true, true><<<grid, SWEEP_WARPS * 32, 0, s>>>(p);
}

template <int NREG>
static void sweep_dispatch(const SweepParams& p, int mode, bool wrap, bool adapt, dim3 grid, cudaStream_t s)
{
    if (mode == LM_FULL) sweep_dispatch2<NREG, LM_FULL>(p, wrap, adapt, grid, s);
    else if (mode == LM_VECPAD) sweep_dispatch2<NREG, LM_VECPAD>(p, wrap, adapt, grid, s);
    else sweep_dispatch2<NREG, LM_BYTES>(p, wrap, adapt, grid, s);
}

bool sweep_needs_wrap(int P1, int P2, int cmax)
{
    return !(P1 >= 0 && P2 >= 0 && cmax + P1 + P2 <= 255 && 2 * cmax + P2 <= 255);
}

int launch_sweeps(fsgm_ctx* c, int n, const uint8_t* C, const uint8_t* I1, int W, int H, int D,
                  int P1, int P2, int adaptive_thr, int cmax, const int* dirs, int n_dirs, uint8_t* const* Lvols)
{
    if (D < 1 || D > 512) return fail(c, FSGM_ERR_DOMAIN, "label count must be in 1..512");
    if (n_dirs < 1 || n_dirs > 8) return fail(c, FSGM_ERR_ARG, "n_dirs");
    StageScope ss(c, ST_SWEEP);
    SweepParams p{};
    p.C = C; p.I1 = I1; p.n_dirs = n_dirs; p.W = W; p.H = H; p.D = D; p.P1 = P1; p.P2 = P2; p.adaptive_thr = adaptive_thr;
    p.line_start[0] = 0;
    for (int k = 0; k < n_dirs; ++k) {
        p.dir[k] = dirs[k];
        p.L[k] = Lvols[k];
        p.line_start[k + 1] = p.line_start[k] + (dir_dy(dirs[k]) == 0 ? H : W);
    }
    const int nreg = D <= 64 ? 1 : D <= 128 ? 2 : D <= 256 ? 4 : 8;
    const int nb = 2 * nreg;
    const int mode = (D == 32 * nb) ? LM_FULL : (D % nb == 0 ? LM_VECPAD : LM_BYTES);
    const bool wrap = sweep_needs_wrap(P1, P2, cmax);
    const bool adapt = adaptive_thr > 0;
    dim3 grid((p.line_start[n_dirs] + SWEEP_WARPS - 1) / SWEEP_WARPS, n);
    bool all_horizontal = true;
    for (int k = 0; k < n_dirs; ++k) all_horizontal &= dir_dy(dirs[k]) == 0;
    if (all_horizontal && mode == LM_FULL && !wrap && !adapt && nreg <= 4) {
        if (nreg == 4) hsweep_tma_kernel<4><<<grid, SWEEP_WARPS * 32, 0, c->stream>>>(p);
        else if (nreg == 2) hsweep_tma_kernel<2><<<grid, SWEEP_WARPS * 32, 0, c->stream>>>(p);
        else hsweep_tma_kernel<1><<<grid, SWEEP_WARPS * 32, 0, c->stream>>>(p);
        FSGM_LAUNCHED(c);
        return FSGM_OK;
    }
    switch (nreg) {
        case 1: sweep_dispatch<1>(p, mode, wrap, adapt, grid, c->stream); break;
        case 2: sweep_dispatch<2>(p, mode, wrap, adapt, grid, c->stream); break;
        case 4: sweep_dispatch<4>(p, mode, wrap, adapt, grid, c->stream); break;
        default: sweep_dispatch<8>(p, mode, wrap, adapt, grid, c->stream); break;
    }
    FSGM_LAUNCHED(c);
    return FSGM_OK;
}

// Direction-split form (dist.cu): the listed directions of ONE pair, every pixel's L row written into the receive buffer of the
// rank that owns the pixel's slab (see SweepParams).  No-wrap domain, no adaptive P2 (the caller falls back to local volumes +
// an NCCL exchange otherwise).
int launch_sweeps_scatter(fsgm_ctx* c, const uint8_t* C, int W, int H, int D, int P1, int P2, const int* dirs, const int* slots,
                          int n_dirs, uint8_t* const* peer, int world, size_t stripe_pixels, size_t local_pixels)
{
    if (D < 1 || D > 512) return fail(c, FSGM_ERR_DOMAIN, "label count must be in 1..512");
    if (n_dirs < 1 || n_dirs > 8 || world < 1 || world > 16) return fail(c, FSGM_ERR_ARG, "n_dirs / world");
    if (sweep_needs_wrap(P1, P2, 24)) return fail(c, FSGM_ERR_DOMAIN, "scatter sweeps need the no-wrap parameter domain");
    StageScope ss(c, ST_SWEEP);
    SweepParams p{};
    p.C = C; p.n_dirs = n_dirs; p.W = W; p.H = H; p.D = D; p.P1 = P1; p.P2 = P2; p.adaptive_thr = 0;
    p.stripe_pixels = (unsigned)stripe_pixels; p.chunk_pixels = (unsigned)stripe_pixels + 1; p.local_pixels = (unsigned)local_pixels; p.world = world;
    for (int j = 0; j < world; ++j) p.peer[j] = peer[j];
    p.line_start[0] = 0;
    for (int k = 0; k < n_dirs; ++k) {
        p.dir[k] = dirs[k]; p.slot[k] = slots[k]; p.L[k] = nullptr;
        p.line_start[k + 1] = p.line_start[k] + (dir_dy(dirs[k]) == 0 ? H : W);
    }
    const int nreg = D <= 64 ? 1 : D <= 128 ? 2 : D <= 256 ? 4 : 8;
    const int nb = 2 * nreg;
    const int mode = (D == 32 * nb) ? LM_FULL : (D % nb == 0 ? LM_VECPAD : LM_BYTES);
    dim3 grid((p.line_start[n_dirs] + SWEEP_WARPS - 1) / SWEEP_WARPS, 1);
#define SC_GO(NR) do { \
        if (mode == LM_FULL) sweep_fast_kernel<NR, LM_FULL, false, true><<<grid, SWEEP_WARPS * 32, 0, c->stream>>>(p); \
        else if (mode == LM_VECPAD) sweep_fast_kernel<NR, LM_VECPAD, false, true><<<grid, SWEEP_WARPS * 32, 0, c->stream>>>(p); \
        else sweep_fast_kernel<NR, LM_BYTES, false, true><<<grid, SWEEP_WARPS * 32, 0, c->stream>>>(p); } while (0)
    switch (nreg) { case 1: SC_GO(1); break; case 2: SC_GO(2); break; case 4: SC_GO(4); break; default: SC_GO(8); break; }
#undef SC_GO
    FSGM_LAUNCHED(c);
    return FSGM_OK;
}

}  // namespace fsgm
