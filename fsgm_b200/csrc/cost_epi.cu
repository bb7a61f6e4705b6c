// Epipolar matching-cost volume — replaces calc_cost() (reference calc_cost_sgm.cpp:319-412).
//
//   raw[p][d] = popc(cen1[p] ^ cen2[q(p,d)])            q = clamp(round(Pd0(p) - 1 + (O(p)*vz(d))*u(p)))
//   C[p][d]   = (u8)(1.0*sum_{5x5}(raw[.][d])/25 + 0.5)  replicate border over the COST volume (:387-404)
//
// Geometry is evaluated in fp64 with explicit round-to-nearest mul/add (no FMA contraction — the
// reference is compiled without it and a fused multiply-add changes round() outcomes), C `round`
// (half away from zero), then the x86 double->int conversion the reference gets from gcc
// (cvttsd2si: out-of-range / NaN -> INT_MIN, which the following clamp turns into 0).
// vz(d) is tabulated on the host with the reference's exact expression (:360-361).
//
// Layout: label-contiguous u8 [pair][y][x][d]; a warp writes 128 B (uchar4 per lane) per store.
#include "fsgm_internal.h"
#include <type_traits>

namespace fsgm {

__device__ __forceinline__ int ref_round_to_index(double v, int hi)
{
    // C round(): half away from zero.  v - trunc(v) is exact, so the tie test is exact too.
    double t = trunc(v);
    double f = __dsub_rn(v, t);
    if (fabs(f) >= 0.5) t = __dadd_rn(t, copysign(1.0, v));
    // x86 cvttsd2si semantics followed by clamp(.,0,hi) (calc_cost_sgm.cpp:371-375)
    if (!(t < 2147483648.0)) return 0;                  // too large or NaN -> INT_MIN -> 0
    int i = __double2int_rz(t);                         // saturates to INT_MIN below -2^31 -> 0 after clamp
    return min(max(i, 0), hi);
}

// One warp per pixel, each lane 4 consecutive labels per 128-label chunk.
__global__ void __launch_bounds__(256)
epi_raw_cost_kernel(const uint32_t* __restrict__ cen1, const uint32_t* __restrict__ cen2,
                    const double* __restrict__ Pd0, const double* __restrict__ dirn, const double* __restrict__ O,
                    const double* __restrict__ vz, int W, int H, int D, uint8_t* __restrict__ raw)
{
    const size_t N = (size_t)W * H;
    const int lane = threadIdx.x & 31;
    const size_t warp = (size_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const int pair = blockIdx.y;
    if (warp >= N) return;
    const size_t p = warp, gp = pair * N + p;
    const double bx = __dsub_rn(Pd0[pair * 2 * N + p], 1.0), by = __dsub_rn(Pd0[pair * 2 * N + N + p], 1.0);
    const double ux = dirn[pair * 2 * N + p], uy = dirn[pair * 2 * N + N + p], off = O[gp];
    const uint32_t c1 = cen1[gp];
    const uint32_t* c2 = cen2 + pair * N;
    uint8_t* out = raw + gp * D;
    const bool vec = (D & 3) == 0;
    for (int d0 = lane * 4; d0 < D; d0 += 128) {
        uint32_t packed = 0;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            int d = d0 + j;
            uint32_t h = 0;
            if (d < D) {
                double t = __dmul_rn(off, vz[d]);
                int x2 = ref_round_to_index(__dadd_rn(bx, __dmul_rn(t, ux)), W - 1);
                int y2 = ref_round_to_index(__dadd_rn(by, __dmul_rn(t, uy)), H - 1);
                h = __popc(c1 ^ __ldg(c2 + (size_t)y2 * W + x2));
            }
            packed |= h << (8 * j);
        }
        if (vec) *reinterpret_cast<uint32_t*>(out + d0) = packed;
        else
            for (int j = 0; j < 4 && d0 + j < D; ++j) out[d0 + j] = (uint8_t)(packed >> (8 * j));
    }
}

// 5x5 box over (x,y) per label, replicate border.  Thread = (x, 4 labels), marching down a row segment
// with a register ring of the five horizontal 5-sums (each <= 5*24 so four of them fit one 32-bit word).
// (u8)(1.0*s/25 + 0.5) == (2*s + 25) / 50 in integers: 2*s+25 is odd, never a multiple of 50, and the
// fp64 quotient is within 1e-13 of a value whose distance to the next integer is >= 1/50.
__device__ __forceinline__ uint32_t box_norm4(uint32_t lo, uint32_t hi)   // lo: bytes 0,2 as u16x2; hi: bytes 1,3
{
    uint32_t a = ((2 * (lo & 0xFFFF) + 25) * 1311u) >> 16;      // /50 for values < 2^12: 1311/65536 = 1/49.99
    uint32_t b = ((2 * (lo >> 16) + 25) * 1311u) >> 16;
    uint32_t c = ((2 * (hi & 0xFFFF) + 25) * 1311u) >> 16;
    uint32_t d = ((2 * (hi >> 16) + 25) * 1311u) >> 16;
    return a | (c << 8) | (b << 16) | (d << 24);
}

constexpr int BOX_ROWS = 16;     // rows produced per thread (plus 4 warm-up rows)

__global__ void __launch_bounds__(256)
epi_box5_kernel(const uint8_t* __restrict__ raw, uint8_t* __restrict__ C, int W, int H, int D)
{
    // thread -> (x, label quad).  D4 = D/4 quads; blockDim.x threads cover `xs` columns x D4 quads.
    const int D4 = D >> 2;
    const int xs = blockDim.x / D4;
    const int q = threadIdx.x % D4, xl = threadIdx.x / D4;
    if (xl >= xs) return;
    const int x = blockIdx.x * xs + xl;
    if (x >= W) return;
    const size_t N = (size_t)W * H;
    const uint32_t* src = reinterpret_cast<const uint32_t*>(raw + blockIdx.z * N * D);
    uint32_t* dst = reinterpret_cast<uint32_t*>(C + blockIdx.z * N * D);
    const int y0 = blockIdx.y * BOX_ROWS;
    int xc[5];
#pragma unroll
    for (int k = 0; k < 5; ++k) xc[k] = min(max(x + k - 2, 0), W - 1);
    auto hsum = [&](int y) -> uint32_t {            // horizontal 5-sum of four labels, bytes stay < 256
        int yy = min(max(y, 0), H - 1);
        const uint32_t* row = src + (size_t)yy * W * D4 + q;
        uint32_t s = 0;
#pragma unroll
        for (int k = 0; k < 5; ++k) s += __ldg(row + (size_t)xc[k] * D4);
        return s;
    };
    uint32_t h0 = hsum(y0 - 2), h1 = hsum(y0 - 1), h2 = hsum(y0), h3 = hsum(y0 + 1);
    for (int y = y0; y < min(y0 + BOX_ROWS, H); ++y) {
        uint32_t h4 = hsum(y + 2);
        uint32_t lo = (h0 & 0x00FF00FFu) + (h1 & 0x00FF00FFu) + (h2 & 0x00FF00FFu) + (h3 & 0x00FF00FFu) + (h4 & 0x00FF00FFu);
        uint32_t hi = ((h0 >> 8) & 0x00FF00FFu) + ((h1 >> 8) & 0x00FF00FFu) + ((h2 >> 8) & 0x00FF00FFu) +
                      ((h3 >> 8) & 0x00FF00FFu) + ((h4 >> 8) & 0x00FF00FFu);
        dst[((size_t)y * W + x) * D4 + q] = box_norm4(lo, hi);
        h0 = h1; h1 = h2; h2 = h3; h3 = h4;
    }
}

// generic fallback for D not a multiple of 4: one thread per voxel
__global__ void epi_box5_scalar_kernel(const uint8_t* __restrict__ raw, uint8_t* __restrict__ C, int W, int H, int D)
{
    const size_t N = (size_t)W * H, V = N * D;
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= V) return;
    const uint8_t* src = raw + blockIdx.z * V;
    int d = (int)(i % D);
    size_t p = i / D;
    int x = (int)(p % W), y = (int)(p / W);
    uint32_t s = 0;
    for (int dy = -2; dy <= 2; ++dy)
        for (int dx = -2; dx <= 2; ++dx) {
            int yy = min(max(y + dy, 0), H - 1), xx = min(max(x + dx, 0), W - 1);
            s += src[((size_t)yy * W + xx) * D + d];
        }
    C[blockIdx.z * V + i] = (uint8_t)((2 * s + 25) / 50);
}

// vz(d) = r/(1-r), r = 1.0*d/(D+1)*vMax — the reference's expression and evaluation order
// (calc_cost_sgm.cpp:339,360-361); IEEE double div/mul on the device are bit-identical to SSE2.
__global__ void vz_table_kernel(int D, double vMax, double* __restrict__ vz)
{
    int d = blockIdx.x * blockDim.x + threadIdx.x;
    if (d >= D) return;
    double r = __dmul_rn(__ddiv_rn((double)d, (double)(D + 1)), vMax);
    vz[d] = __ddiv_rn(r, __dsub_rn(1.0, r));
}

int launch_vz_table(fsgm_ctx* c, int D, double vMax, double* d_vz)
{
    vz_table_kernel<<<(D + 127) / 128, 128, 0, c->stream>>>(D, vMax, d_vz);
    FSGM_LAUNCHED(c);
    return FSGM_OK;
}

int launch_epi_cost(fsgm_ctx* c, int n, const double* d_vz, const uint32_t* cen1, const uint32_t* cen2, int W, int H, int D, double vMax,
                    const double* Pd0, const double* dirn, const double* O, uint8_t* raw, uint8_t* C)
{
    StageScope ss(c, ST_EPI_COST);
    (void)vMax;
    const size_t N = (size_t)W * H;
    {
        dim3 grid((unsigned)((N + 7) / 8), n);
        epi_raw_cost_kernel<<<grid, 256, 0, c->stream>>>(cen1, cen2, Pd0, dirn, O, d_vz, W, H, D, raw);
        FSGM_LAUNCHED(c);
    }
    if ((D & 3) == 0 && D / 4 <= 256) {
        int D4 = D / 4, xs = 256 / D4;
        dim3 grid((W + xs - 1) / xs, (H + BOX_ROWS - 1) / BOX_ROWS, n);
        epi_box5_kernel<<<grid, 256, 0, c->stream>>>(raw, C, W, H, D);
    } else {
        size_t V = N * D;
        dim3 grid((unsigned)((V + 255) / 256), 1, n);
        epi_box5_scalar_kernel<<<grid, 256, 0, c->stream>>>(raw, C, W, H, D);
    }
    FSGM_LAUNCHED(c);
    return FSGM_OK;
}



// =====================================================================================================
// Fused cost kernel: raw Hamming cost and the 5x5 box filter in one pass, the raw volume never touches HBM.
//
// A CTA owns a strip of TX = fc_tx(D/4) output columns and marches down FC_TY rows.  Per image row it (1) evaluates the raw
// cost of the TX+4 columns (2-px halo, replicate border = clamped coordinates) for all D labels into shared
// memory, (2) forms the horizontal 5-sums, keeps them in a 5-row shared-memory ring and, once five rows are in,
// emits the vertical sum normalised to (2*s+25)/50 as u8, label-contiguous.  Redundant work: (TX+4)/TX * (TY+4)/TY.
//
// Rounding: the reference needs clamp((int)round(v), 0, hi) with C round() (half away from zero) and the x86
// out-of-range conversion (cvttsd2si -> INT_MIN -> clamped to 0).  Working on w = 2v (exact: the per-pixel constants
// are doubled once, and scaling by two commutes with every IEEE rounding in bx + (off*vz)*ux):
//   v >= 0:  round(v) = floor(v + 0.5) = (floor(w) + 1) >> 1     (integer identity, no double rounding)
//   v <  0:  round(v) <= 0, clamps to 0
//   NaN, or round(v) >= 2^31: INT_MIN on x86, clamps to 0
// Fast path: k = clamp(floor(w) + 1, 0, 2*hi + 1) and round-clamp(v) = k >> 1, with floor(w) + 1 taken as the low mantissa word
// of w + (1.5 * 2^52 + 1) rounded DOWN (two's complement for negative w, exact for |w| < 2^31 - 2) and the clamp as one
// VIMNMX.RELU: the conversion runs on the fp64 pipe instead of the quarter-rate XU pipe, which POPC alone keeps busy (ncu r1s:
// two F2I + POPC per voxel = 8 of every ~36 issue cycles each).  A strip row with a pixel whose rays can leave |w| < 2^30 — or
// hold a NaN (non-finite input, inf*0) — takes the checked conversion (unsigned floor conversion, NaN test) as a whole; the
// bound |2b| + |2 off| * max|vz| * |u| is evaluated per pixel and row by the warps themselves.
//
// Mapping.  Raw-cost phase: LANE = PIXEL of the strip row (strip + halo = 32 pixels at D = 256), WARP = QPW consecutive label
// quads.  Neighbouring pixels look at neighbouring census words for the same label, so one gather touches 4-5 sectors (with lane =
// label quad of ONE pixel a warp's addresses were strung along the epipolar line: 17 sectors, 5.6 L1 wavefronts per request, ncu
// r1q).  The four gathers of a quad are consumed one quad later (software pipeline: the compiler alone kept one quad in flight and
// waited for it, 2015 -> 2062 pairs/s).  Box phase: the SAME warp filters its own QPW quads — lane = (quad, group of XP columns) —
// so the raw row is warp-private ([quad][pixel] words, pitch PQ chosen so that both phases are bank-conflict free) and the
// row needs ONE block barrier, at its end, where the next row's geometry (cp.async, a row ahead, no staging registers) must have
// landed; with a CTA-wide raw row a second barrier sat between the phases and held 12 % of the stall samples (ncu r2s).
// Box filter: horizontal 5-sums slide over the lane's XP columns (bytes <= 120), the vertical 5-sum is a running u16x2 pair
// (one PRMT per half of the entering and of the leaving row, one IADD3), and the /25 rounding is ONE fp16 fma per pair: the
// integer s < 1024 in a 16-bit half IS the fp16 subnormal s * 2^-24, and fp16(2^20 / 25) * (s * 2^-24) + 64 rounds (to a multiple
// of 2^-4, the spacing at 64) to 64 + round(s / 25) / 16, pattern 0x5400 | q — checked for every s <= 1023: the relative error of
// the constant times s stays below the 1/50 that separates s / 25 from a rounding boundary.
// History of the row loop (instructions per row and thread / pairs per s at 60 pairs per step): round 1 1091 / 1912; index
// clamp on the doubled value, word-index gathers, IMAD packing, fp16 normalisation, cp.async geometry 765 / 1981; magic-number
// conversion 2013; pipelined gathers 2062.
constexpr int FC_THREADS = 256, FC_WARPS = FC_THREADS / 32;          // rows per CTA are chosen at launch (32..128)
__host__ __device__ constexpr int fc_tx(int D4) { return D4 == 64 ? 28 : D4 == 32 ? 24 : 16; }      // strip width: TX + 4 <= 32 lanes
__host__ __device__ constexpr int fc_pq(int D4) { return D4 == 64 ? 36 : D4 == 32 ? 40 : 48; }      // raw-tile pitch per quad (words)

__device__ __forceinline__ uint32_t ref_round_clamp_w(double w, uint32_t hi)          // checked path: all of the table above but NaN
{
    const uint32_t f = __double2uint_rd(w);
    return min((f + 1u) >> 1, hi);
}
__device__ __forceinline__ uint32_t ref_round_clamp_k(double w, uint32_t k2)          // fast path, doubled index
{
    const int L = __double2loint(__dadd_rd(w, 6755399441055745.0));
    return (uint32_t)__vimin_s32_relu(L, (int)k2);
}
__device__ __forceinline__ uint32_t mad_u32(uint32_t a, uint32_t b, uint32_t c)
{
    uint32_t r;
    asm("mad.lo.u32 %0, %1, %2, %3;" : "=r"(r) : "r"(a), "r"(b), "r"(c));
    return r;
}
__device__ __forceinline__ uint32_t h2fma(uint32_t a, uint32_t b, uint32_t c) { uint32_t r; asm("fma.rn.f16x2 %0, %1, %2, %3;" : "=r"(r) : "r"(a), "r"(b), "r"(c)); return r; }

// measured (60 pairs per step): 4 CTAs/SM 2014 pairs/s, 3 CTAs/SM (84 registers) 1987, 2 CTAs/SM 1888
#ifndef FSGM_FC_MINB
#define FSGM_FC_MINB 4
#endif
// vz(d) travels as a kernel parameter (constant bank, warp-uniform LDC reads): the table is evaluated on the host with the
// reference's expression (the reference IS host code: same IEEE operations) and 2 KB of shared memory stay free — four CTAs per SM
// then fit the 196 KB carve-out and the gathers keep 60 KB of L1 (one more KB per CTA and the carve-out jumps to 228 KB).
template <int D4> struct alignas(16) FcVzTable { double v[4 * D4]; };
template <int D4>      // D = 4*D4 labels, D4 in {16, 32, 64}
__global__ void __launch_bounds__(FC_THREADS, FSGM_FC_MINB)
epi_cost_fused_kernel(const uint32_t* __restrict__ cen1, const uint32_t* __restrict__ cen2,
                      const double* __restrict__ Pd0, const double* __restrict__ dirn, const double* __restrict__ O,
                      const __grid_constant__ FcVzTable<D4> vzt, int W, int H, int FC_TY, uint8_t* __restrict__ C)
{
    constexpr int FC_TX = fc_tx(D4), NPIX = FC_TX + 4, PQ = fc_pq(D4);
    constexpr int QPW = D4 / FC_WARPS;               // label quads per warp (both phases)
    constexpr int NCG = 32 / QPW, XP = FC_TX / NCG;  // box phase: lane = (quad, column group), XP consecutive output columns per lane
    static_assert(NPIX <= 32 && NPIX <= PQ && QPW >= 1 && NCG * QPW == 32 && XP * NCG == FC_TX, "strip geometry");
    extern __shared__ __align__(16) unsigned char smem_raw[];
    double* geo = reinterpret_cast<double*>(smem_raw);                                      // [2][NPIX][5] raw planes of a strip row
    uint32_t* rawt = reinterpret_cast<uint32_t*>(geo + 2 * NPIX * 5);                       // [warp][QPW][PQ]
    uint32_t* hring = rawt + FC_WARPS * QPW * PQ;                                           // [warp][5][FC_TX][QPW]
    uint32_t* gcen = hring + 5 * FC_TX * D4;                                                // [2][NPIX], then one word for max|vz|

    const size_t N = (size_t)W * H;
    const int pair = blockIdx.z, x0 = blockIdx.x * FC_TX, y0 = blockIdx.y * FC_TY;
    const int tid = threadIdx.x;
    const double* PdX = Pd0 + (size_t)pair * 2 * N; const double* PdY = PdX + N;
    const double* DrX = dirn + (size_t)pair * 2 * N; const double* DrY = DrX + N;
    const double* Op = O + pair * N;
    const int lane = tid & 31, warp = __shfl_sync(0xFFFFFFFFu, tid >> 5, 0);      // warp-uniform for the compiler: label-table addresses live in uniform registers
    const int kq = lane % QPW, cg = lane / QPW;                                   // box phase: this lane's quad (of the warp's) and column group
    uint32_t* Cout = reinterpret_cast<uint32_t*>(C + pair * N * (size_t)(4 * D4)) + warp * QPW + kq;
    uint32_t* rawt_w = rawt + warp * (QPW * PQ);
    uint32_t* hring_w = hring + warp * (5 * FC_TX * QPW);

    // upper bound of |vz| over the table (high word + 1; inf / NaN give a non-finite bound and every row takes the checked path)
    uint32_t* vzhi_s = gcen + 2 * NPIX;
    if (tid == 0) *vzhi_s = 0;
    for (int i = tid; i < 5 * FC_TX * D4; i += FC_THREADS) hring[i] = 0;      // rows before the window count as zero
    __syncthreads();
    uint32_t vzhi = 0;
    for (int i = tid; i < 4 * D4; i += FC_THREADS) vzhi = max(vzhi, (uint32_t)__double2hiint(fabs(vzt.v[i])));
    vzhi = __reduce_max_sync(0xFFFFFFFFu, vzhi);
    if (lane == 0) atomicMax(vzhi_s, vzhi);

    const int yend = min(y0 + FC_TY, H);
    const uint32_t wmax = (uint32_t)(W - 1), hmax = (uint32_t)(H - 1);
    const uint32_t wk2 = 2u * wmax + 1u, hk2 = 2u * hmax + 1u, Wu = (uint32_t)W;
    const uint32_t* c2w = cen2 + pair * N;
    asm volatile("" : "+l"(c2w));                    // one opaque 64-bit base: a gather's address is IMAD.WIDE.U32(index, 4, base), not a rebuilt sum
    uint32_t vlo[XP], vhi[XP];                                               // running vertical 5-sums (u16x2) per column
#pragma unroll
    for (int k = 0; k < XP; ++k) { vlo[k] = 0; vhi[k] = 0; }

    // geometry rows: global -> shared by cp.async, one 8-byte (census: 4-byte) element per thread, a row ahead
    constexpr int GEO_T = 6 * NPIX;                                           // threads that copy: plane = tid / NPIX, pixel = tid % NPIX
    static_assert(GEO_T <= FC_THREADS, "one geometry element per thread");
    const int gpl = tid / NPIX, gpx = tid - gpl * NPIX;
    const int gxc = min(max(x0 - 2 + gpx, 0), W - 1);
    const char* gsrc = gpl == 0 ? (const char*)PdX : gpl == 1 ? (const char*)PdY : gpl == 2 ? (const char*)DrX : gpl == 3 ? (const char*)DrY
                     : gpl == 4 ? (const char*)Op : (const char*)(cen1 + pair * N);
    const uint32_t gdst0 = gpl < 5 ? (uint32_t)__cvta_generic_to_shared(geo + gpx * 5 + gpl) : (uint32_t)__cvta_generic_to_shared(gcen + gpx);
    auto fetch_geo = [&](int r, int buf) {
        if (tid < GEO_T) {
            const int yc = min(max(r, 0), H - 1);
            const uint32_t p = (uint32_t)yc * Wu + (uint32_t)gxc;
            if (gpl < 5) asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" :: "r"(gdst0 + buf * (NPIX * 5 * 8)), "l"(gsrc + (size_t)p * 8) : "memory");
            else         asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" :: "r"(gdst0 + buf * (NPIX * 4)), "l"(gsrc + (size_t)p * 4) : "memory");
        }
    };
    fetch_geo(y0 - 2, 0);
    asm volatile("cp.async.wait_all;" ::: "memory");
    __syncthreads();
    const double vzmax = __hiloint2double((int)(*vzhi_s + 1u), 0);

    for (int r = y0 - 2; r < yend + 2; ++r) {
        const int buf = (r - (y0 - 2)) & 1;
        const double* geo_r = geo + buf * NPIX * 5;
        if (r + 1 < yend + 2) fetch_geo(r + 1, buf ^ 1);
        // this lane's pixel: doubled constants, w = 2*(b + (off*vz)*u) = 2b + ((2 off)*vz)*u exactly
        const int gl = min(lane, NPIX - 1);
        const double a0 = geo_r[gl * 5], a1 = geo_r[gl * 5 + 1], ux = geo_r[gl * 5 + 2], uy = geo_r[gl * 5 + 3], a4 = geo_r[gl * 5 + 4];
        const double bx = __dmul_rn(__dsub_rn(a0, 1.0), 2.0), by = __dmul_rn(__dsub_rn(a1, 1.0), 2.0), off = __dmul_rn(a4, 2.0);
        const uint32_t c1 = gcen[buf * NPIX + gl];
        // every |w| of the pixel stays below 2^30 (a NaN anywhere fails the comparison): the magic-number conversion is exact
        const double reach = __dmul_rn(fabs(off), vzmax);
        const bool fin = __dadd_rn(fabs(bx), __dmul_rn(reach, fabs(ux))) < 1073741824.0 && __dadd_rn(fabs(by), __dmul_rn(reach, fabs(uy))) < 1073741824.0;
        const int slow = __any_sync(0xFFFFFFFFu, !fin);
        // (1) raw cost of NPIX pixels x the warp's QPW label quads
        if (lane < NPIX) {
            uint32_t* rr_out = rawt_w + lane;
            if (!slow) {
                uint32_t g[4];
#pragma unroll
                for (int k = 0; k < QPW + 1; ++k) {
                    const int qq = warp * QPW + k;
                    uint32_t idx[4] = {0, 0, 0, 0};
                    if (k < QPW) {
#pragma unroll
                        for (int j = 0; j < 4; ++j) {
                            const double t = __dmul_rn(off, vzt.v[4 * qq + j]);
                            const uint32_t xk = ref_round_clamp_k(__dadd_rn(bx, __dmul_rn(t, ux)), wk2);
                            const uint32_t yk = ref_round_clamp_k(__dadd_rn(by, __dmul_rn(t, uy)), hk2);
                            idx[j] = (yk >> 1) * Wu + (xk >> 1);                   // 32-bit word index: IMAD + IMAD.WIDE
                        }
                    }
                    if (k >= 1) {
                        const uint32_t h0 = __popc(c1 ^ g[0]), h1 = __popc(c1 ^ g[1]), h2 = __popc(c1 ^ g[2]), h3 = __popc(c1 ^ g[3]);
                        rr_out[(k - 1) * PQ] = mad_u32(mad_u32(h3, 256u, h2), 65536u, mad_u32(h1, 256u, h0));      // three IMADs (FMA pipe)
                    }
                    if (k < QPW) {
#pragma unroll
                        for (int j = 0; j < 4; ++j) g[j] = __ldg(c2w + idx[j]);
                    }
                }
            } else {
                for (int k = 0; k < QPW; ++k) {
                    const int qq = warp * QPW + k;
                    uint32_t packed = 0;
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        const double t = __dmul_rn(off, vzt.v[4 * qq + j]);
                        const double wx = __dadd_rn(bx, __dmul_rn(t, ux)), wy = __dadd_rn(by, __dmul_rn(t, uy));
                        const uint32_t x2 = (wx != wx) ? 0u : ref_round_clamp_w(wx, wmax);
                        const uint32_t y2 = (wy != wy) ? 0u : ref_round_clamp_w(wy, hmax);
                        packed |= (uint32_t)__popc(c1 ^ __ldg(c2w + (y2 * Wu + x2))) << (8 * j);
                    }
                    rr_out[k * PQ] = packed;
                }
            }
        }
        __syncwarp();
        // (2) horizontal 5-sums by a sliding window over this lane's XP consecutive columns; vertical 5-sums as
        //     running sums (add the new row, drop the row that leaves the window: it sits in the ring slot being rewritten)
        const int slot = (r + 10) % 5;
        const uint32_t* rr = rawt_w + kq * PQ + cg * XP;
        uint32_t* hr = hring_w + (slot * FC_TX + cg * XP) * QPW + kq;
        uint32_t h = rr[0] + rr[1] + rr[2] + rr[3] + rr[4];                                  // bytes <= 120: no carries
        const int yo = r - 2;
        const bool emit = yo >= y0;
        uint32_t* crow = Cout + ((size_t)yo * W + x0 + cg * XP) * D4;
        auto box_cols = [&](auto checked) {
#pragma unroll
            for (int k = 0; k < XP; ++k) {
                if (k) h = h - rr[k - 1] + rr[k + 4];
                const uint32_t old = hr[k * QPW];
                hr[k * QPW] = h;
                vlo[k] = vlo[k] + __byte_perm(h, 0u, 0x4240) - __byte_perm(old, 0u, 0x4240);
                vhi[k] = vhi[k] + __byte_perm(h, 0u, 0x4341) - __byte_perm(old, 0u, 0x4341);
                const uint32_t out = __byte_perm(h2fma(vlo[k], 0x791F791Fu, 0x54005400u), h2fma(vhi[k], 0x791F791Fu, 0x54005400u), 0x6240);
                if (emit && (!decltype(checked)::value || x0 + cg * XP + k < W)) crow[k * D4] = out;
            }
        };
        if (x0 + FC_TX <= W) box_cols(std::false_type{}); else box_cols(std::true_type{});      // CTA-uniform: only the last strip tests columns
        asm volatile("cp.async.wait_all;" ::: "memory");
        __syncthreads();                              // next row's geometry has landed (and this warp is done with its raw tile)
    }
}

static size_t fused_cost_smem(int D4)
{
    const size_t tx = fc_tx(D4), npix = tx + 4, qpw = D4 / FC_WARPS;
    return 2 * npix * 5 * 8 + (size_t)FC_WARPS * qpw * fc_pq(D4) * 4 + 5 * tx * D4 * 4 + 2 * npix * 4 + 64;
}

// Rows per CTA: tall strips waste less on the 4 warm-up rows, but a small batch needs enough CTAs to fill the GPU.  (A list-schedule
// estimate that trims the partly filled last round of CTAs — 2025 CTAs on 592 resident slots are 3.4 rounds — picked 54 rows at
// KITTI size; measured in the two-stream wave pipeline, where other kernels fill that tail: 128 rows 2014 pairs/s, 96: 2013, 75: 2010,
// 63: 2005, 54: 2001, 47: 1996.  fsgm_tune key 8 overrides.)
static int fused_cost_rows(fsgm_ctx* c, int W, int H, int n, int tx)
{
    int ty = 128;
    while (ty > 32 && (size_t)((W + tx - 1) / tx) * ((H + ty - 1) / ty) * n < (size_t)c->sm_count * 8) ty >>= 1;
    return ty;
}

// returns FSGM_OK and sets *done = true if the fused kernel handles this label count
int launch_epi_cost_fused(fsgm_ctx* c, int n, double vMax, const uint32_t* cen1, const uint32_t* cen2, int W, int H, int D,
                          const double* Pd0, const double* dirn, const double* O, uint8_t* C, bool* done)
{
    *done = false;
    if (D != 64 && D != 128 && D != 256) return FSGM_OK;
    if ((size_t)W * H >= (size_t)1 << 31) return FSGM_OK;
    StageScope ss(c, ST_EPI_COST);
    const int D4 = D / 4;
    const size_t smem = fused_cost_smem(D4);
    // rows per CTA: tall strips waste less on the 4 warm-up rows, but a small batch needs enough CTAs to fill the GPU
    const int tx = fc_tx(D4);
    int ty = c->fc_rows > 0 ? c->fc_rows : fused_cost_rows(c, W, H, n, tx);
    dim3 grid((W + tx - 1) / tx, (H + ty - 1) / ty, n);
#define FSGM_FC(D4V)                                                                                              \
    do {                                                                                                          \
        const unsigned bit = 1u << (D4V == 16 ? 0 : D4V == 32 ? 1 : 2);                                           \
        if (!(c->attr_mask & bit)) {                                                                              \
            FSGM_CUDA(c, cudaFuncSetAttribute(epi_cost_fused_kernel<D4V>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); \
            c->attr_mask |= bit;                                                                                  \
        }                                                                                                         \
        FcVzTable<D4V> vzt;                                                                                       \
        for (int d = 0; d < D; ++d) {        /* calc_cost_sgm.cpp:339,360-361, the reference's evaluation order */   \
            const double r = 1.0 * d / (D + 1) * vMax;                                                                \
            vzt.v[d] = r / (1 - r);                                                                                   \
        }                                                                                                             \
        epi_cost_fused_kernel<D4V><<<grid, FC_THREADS, smem, c->stream>>>(cen1, cen2, Pd0, dirn, O, vzt, W, H, ty, C); \
    } while (0)
    if (D4 == 16) FSGM_FC(16); else if (D4 == 32) FSGM_FC(32); else FSGM_FC(64);
#undef FSGM_FC
    FSGM_LAUNCHED(c);
    *done = true;
    return FSGM_OK;
}

}  // namespace fsgm
