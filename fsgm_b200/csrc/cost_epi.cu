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
//   v <  0:  round(v) <= 0, clamps to 0                            (cvt.rmi.u32 saturates negatives to 0 -> 0)
//   NaN   :  INT_MIN on x86, clamps to 0                           (cvt of NaN gives 0 -> 0)
//   round(v) >= 2^31, i.e. w >= 2^32 - 1: INT_MIN on x86, clamps to 0   (cvt.rmi.u32 saturates to 0xFFFFFFFF, +1 wraps to 0)
// so one unsigned conversion, an add, a shift and an unsigned min reproduce all of it — except NaN, for which the
// hardware conversion does not return 0.  NaN can only appear when an input is non-finite or |offset| is so large
// that offset*vz overflows (inf*0); such pixels are flagged while staging and take a checked path.
// Raw-cost phase mapping: LANE = PIXEL of the strip row (strip + halo = 32 pixels at D = 256), warp = a set of label quads.
// Neighbouring pixels look at neighbouring census words for the same label, so one gather instruction touches 4-5 sectors; with
// the former mapping (lane = label quad of ONE pixel) a warp's 32 addresses were strung along the epipolar line, 17 sectors
// and 5.6 L1 wavefronts per request, and the L1 data pipe (67 % busy, ncu r1q) limited the kernel as much as instruction issue.
// The pixel's geometry stays in registers for the whole row; the labels' vz values are warp-uniform shared-memory reads.
constexpr int FC_THREADS = 256;                          // rows per CTA are chosen at launch (32..128)
__host__ __device__ constexpr int fc_tx(int D4) { return D4 == 64 ? 28 : D4 == 32 ? 24 : 16; }      // strip width: TX + 4 <= 32 lanes, TX % (256 / D4) == 0

__device__ __forceinline__ uint32_t ref_round_clamp_w(double w, uint32_t hi)
{
    const uint32_t f = __double2uint_rd(w);
    return min((f + 1u) >> 1, hi);
}

// The fast loop keeps the clamped value DOUBLED, k = min(floor(w) + 1, 2*hi + 1) (round(v) = k >> 1), so that add and clamp are
// one instruction (VIADDMNMX.U32: the add wraps, 0xFFFFFFFF + 1 -> 0 as above) and the shift is all that is left.
#ifndef FSGM_FC_VARIANT
#define FSGM_FC_VARIANT 2            // A/B (tools/build_variant.sh, 60 pairs per step): 0 = round-1 loop 1912 pairs/s, 1 = VIADDMNMX form 1981, 2 = magic-number form 2013
#endif
#ifndef FSGM_FC_ASYNC
#define FSGM_FC_ASYNC 1
#endif
#ifndef FSGM_FC_PIPE
#define FSGM_FC_PIPE 1            // gathers consumed one quad later: 2015 -> 2062 pairs/s (lag 2: 2061, lag 3: 2050)
#endif
__device__ __forceinline__ uint32_t mad_u32(uint32_t a, uint32_t b, uint32_t c)
{
    uint32_t r;
    asm("mad.lo.u32 %0, %1, %2, %3;" : "=r"(r) : "r"(a), "r"(b), "r"(c));
    return r;
}
__device__ __forceinline__ uint32_t h2fma(uint32_t a, uint32_t b, uint32_t c) { uint32_t r; asm("fma.rn.f16x2 %0, %1, %2, %3;" : "=r"(r) : "r"(a), "r"(b), "r"(c)); return r; }
__device__ __forceinline__ uint32_t ref_round_clamp_k(double w, uint32_t k2)
{
#if FSGM_FC_VARIANT == 2
    // floor(w) + 1 as the low mantissa word of w + (1.5 * 2^52 + 1) rounded down (two's complement for negative w; exact for
    // |w| < 2^31 - 2, which the staging step checks per pixel), clamped to [0, k2] by one VIMNMX.RELU: the conversion moves from
    // the quarter-rate XU pipe to the fp64 pipe.
    const int L = __double2loint(__dadd_rd(w, 6755399441055745.0));
    return (uint32_t)__vimin_s32_relu(L, (int)k2);
#else
    return __viaddmin_u32(__double2uint_rd(w), 1u, k2);
#endif
}

// (u8)(1.0*s/25 + 0.5) == (2s + 25)/50 == (s*2622 + 32775) >> 16 for s <= 600; the quotient (<= 24) is byte 2 of the product
__device__ __forceinline__ uint32_t box_norm4_fast(uint32_t lo, uint32_t hi)   // lo: labels 0,2 as u16x2; hi: labels 1,3
{
    const uint32_t p0 = (lo & 0xFFFFu) * 2622u + 32775u, p2 = (lo >> 16) * 2622u + 32775u;
    const uint32_t p1 = (hi & 0xFFFFu) * 2622u + 32775u, p3 = (hi >> 16) * 2622u + 32775u;
    return __byte_perm(__byte_perm(p0, p1, 0x0062), __byte_perm(p2, p3, 0x0062), 0x5410);
}

// measured (60 pairs, cost stage, former lane = label-quad mapping): 4 CTAs/SM + full unroll 11.17 ms, 4 + unroll 3 11.75, 3 CTAs/SM 11.29,
// no unroll at 108 registers 13.9
#ifndef FSGM_FC_MINB
#define FSGM_FC_MINB 4
#endif
template <int D4>      // D = 4*D4 labels, D4 in {16, 32, 64}
__global__ void __launch_bounds__(FC_THREADS, FSGM_FC_MINB)
epi_cost_fused_kernel(const uint32_t* __restrict__ cen1, const uint32_t* __restrict__ cen2,
                      const double* __restrict__ Pd0, const double* __restrict__ dirn, const double* __restrict__ O,
                      const double* __restrict__ vz, int W, int H, int FC_TY, uint8_t* __restrict__ C)
{
    constexpr int FC_TX = fc_tx(D4), NPIX = FC_TX + 4, IPT = FC_THREADS / D4, XP = FC_TX / IPT;   // XP consecutive output columns per thread
    constexpr int D4S = D4 + 1;                      // raw-row pitch in words: lanes write different pixels of one quad (bank = pixel + quad)
    constexpr int QPW = D4 / (FC_THREADS / 32);      // label quads per warp in the raw-cost phase
    static_assert(NPIX <= 32 && FC_TX % IPT == 0 && QPW >= 1, "strip geometry");
    extern __shared__ __align__(16) unsigned char smem_raw[];
    double* vzs = reinterpret_cast<double*>(smem_raw);                                      // [4*D4]
    double* geo = vzs + 4 * D4;                                                             // [2][NPIX][5]
    uint32_t* raw_row = reinterpret_cast<uint32_t*>(geo + 2 * NPIX * 5);                    // [NPIX][D4S]
    uint32_t* hring = raw_row + NPIX * D4S;                                                 // [5][FC_TX][D4]
    uint32_t* gcen = hring + 5 * FC_TX * D4;                                                // [2][NPIX]

    const size_t N = (size_t)W * H;
    const int pair = blockIdx.z, x0 = blockIdx.x * FC_TX, y0 = blockIdx.y * FC_TY;
    const int tid = threadIdx.x, q = tid % D4, i0 = tid / D4;
    const char* c2b = reinterpret_cast<const char*>(cen2 + pair * N);          // gathers use a 32-bit BYTE offset from this base
    const uint32_t W4 = (uint32_t)W * 4u;
    const double* PdX = Pd0 + (size_t)pair * 2 * N; const double* PdY = PdX + N;
    const double* DrX = dirn + (size_t)pair * 2 * N; const double* DrY = DrX + N;
    const double* Op = O + pair * N;
    uint32_t* Cout = reinterpret_cast<uint32_t*>(C + pair * N * (size_t)(4 * D4)) + q;
    const int lane = tid & 31, warp = __shfl_sync(0xFFFFFFFFu, tid >> 5, 0);      // warp-uniform for the compiler: label-table addresses live in uniform registers
#if FSGM_FC_VARIANT == 2
    // upper bound of |vz| over the table (high word + 1; inf / NaN give a non-finite bound and every pixel takes the checked path)
    uint32_t* vzhi_s = gcen + 2 * NPIX;
    if (tid == 0) *vzhi_s = 0;
    __syncthreads();
    uint32_t vzhi = 0;
    for (int i = tid; i < 4 * D4; i += FC_THREADS) { const double v = vz[i]; vzs[i] = v; vzhi = max(vzhi, (uint32_t)__double2hiint(fabs(v))); }
    vzhi = __reduce_max_sync(0xFFFFFFFFu, vzhi);
    if ((tid & 31) == 0) atomicMax(vzhi_s, vzhi);
    __syncthreads();
    const double vzmax = __hiloint2double((int)(*vzhi_s + 1u), 0);
#else
    for (int i = tid; i < 4 * D4; i += FC_THREADS) vzs[i] = vz[i];
#endif
    const int yend = min(y0 + FC_TY, H);
    const uint32_t wmax = (uint32_t)(W - 1), hmax = (uint32_t)(H - 1);
    const uint32_t wk2 = 2u * wmax + 1u, hk2 = 2u * hmax + 1u, Wu = (uint32_t)W;
    const uint32_t* c2w = cen2 + pair * N;
    (void)wk2; (void)hk2; (void)Wu; (void)c2w;

    for (int i = tid; i < 5 * FC_TX * D4; i += FC_THREADS) hring[i] = 0;      // rows before the window count as zero
    uint32_t vlo[XP], vhi[XP];                                               // running vertical 5-sums (u16x2) per column
#pragma unroll
    for (int k = 0; k < XP; ++k) { vlo[k] = 0; vhi[k] = 0; }

    // Geometry of a row is staged in a double buffer: the global loads for row r+1 are issued before the raw-cost phase of row r
    // and written to shared memory after its box phase, so their latency hides under the arithmetic and a row needs two
    // barriers instead of three.  `slow` = some pixel of the strip row can produce NaN (checked path, CTA-uniform branch).
    // (One barrier per row — raw row double-buffered, geometry staged two rows ahead, vz read with __ldg — was measured slower,
    // 10.15 -> 10.59 ms per 60 pairs: at four CTAs per SM the extra 7 KB per CTA leave the gathers half the L1.)
#if !FSGM_FC_ASYNC
    auto load_geo = [&](int r, double (&a)[5], uint32_t& cen) {
        const int yc = min(max(r, 0), H - 1), xc = min(max(x0 - 2 + tid, 0), W - 1);
        const size_t p = (size_t)yc * W + xc;
        a[0] = PdX[p]; a[1] = PdY[p]; a[2] = DrX[p]; a[3] = DrY[p]; a[4] = Op[p];
        cen = cen1[pair * N + p];
    };
    auto store_geo = [&](int buf, const double (&a)[5], uint32_t cen) -> int {
        double* g = geo + (buf * NPIX + tid) * 5;
        // doubled constants: w = 2*(b + (off*vz)*u) = 2b + ((2 off)*vz)*u exactly
        g[0] = __dmul_rn(__dsub_rn(a[0], 1.0), 2.0); g[1] = __dmul_rn(__dsub_rn(a[1], 1.0), 2.0);
        g[2] = a[2]; g[3] = a[3]; g[4] = __dmul_rn(a[4], 2.0);
        gcen[buf * NPIX + tid] = cen;
#if FSGM_FC_VARIANT == 2
        // every |w| of this pixel stays below 2^30 (NaN anywhere fails the comparison): the magic-number conversion is exact
        const double reach = __dmul_rn(fabs(g[4]), vzmax);
        const bool fin = __dadd_rn(fabs(g[0]), __dmul_rn(reach, fabs(a[2]))) < 1073741824.0 &&
                         __dadd_rn(fabs(g[1]), __dmul_rn(reach, fabs(a[3]))) < 1073741824.0;
#else
        const bool fin = isfinite(a[0]) && isfinite(a[1]) && isfinite(a[2]) && isfinite(a[3]) && fabs(a[4]) <= 1e300;
#endif
        return fin ? 0 : 1;
    };
#endif
#if FSGM_FC_ASYNC
    // Geometry rows travel global -> shared by cp.async, one 8-byte (census: 4-byte) element per thread, a row ahead: no staging
    // registers live across the raw-cost phase (at 64 registers per thread the eleven of them cost rematerialised addresses all
    // over the row loop).  Each warp turns the raw planes of its lane's pixel into the doubled constants itself and votes on the
    // checked path, so the row's second barrier carries no reduction.
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
#else
    double ga[5] = {0, 0, 0, 0, 0};
    uint32_t gc = 0;
    int slow_mine = 0;
    if (tid < NPIX) { load_geo(y0 - 2, ga, gc); slow_mine = store_geo(0, ga, gc); }
    int slow = __syncthreads_or(slow_mine);
#endif

    for (int r = y0 - 2; r < yend + 2; ++r) {
        const int buf = (r - (y0 - 2)) & 1;
        const bool more = r + 1 < yend + 2;
        const double* geo_r = geo + buf * NPIX * 5;
        const uint32_t* gcen_r = gcen + buf * NPIX;
#if FSGM_FC_ASYNC
        if (more) fetch_geo(r + 1, buf ^ 1);
        const int gl = min(lane, NPIX - 1);
        const double a0 = geo_r[gl * 5], a1 = geo_r[gl * 5 + 1], ux = geo_r[gl * 5 + 2], uy = geo_r[gl * 5 + 3], a4 = geo_r[gl * 5 + 4];
        // doubled constants: w = 2*(b + (off*vz)*u) = 2b + ((2 off)*vz)*u exactly
        const double bx = __dmul_rn(__dsub_rn(a0, 1.0), 2.0), by = __dmul_rn(__dsub_rn(a1, 1.0), 2.0), off = __dmul_rn(a4, 2.0);
#if FSGM_FC_VARIANT == 2
        const double reach = __dmul_rn(fabs(off), vzmax);
        const bool fin = __dadd_rn(fabs(bx), __dmul_rn(reach, fabs(ux))) < 1073741824.0 && __dadd_rn(fabs(by), __dmul_rn(reach, fabs(uy))) < 1073741824.0;
#else
        const bool fin = isfinite(a0) && isfinite(a1) && isfinite(ux) && isfinite(uy) && fabs(a4) <= 1e300;
#endif
        const int slow = __any_sync(0xFFFFFFFFu, !fin);
#else
        if (tid < NPIX && more) load_geo(r + 1, ga, gc);
#endif
        // (1) raw cost of NPIX pixels x D labels: this lane's pixel, this warp's label quads
        if (lane < NPIX) {
#if !FSGM_FC_ASYNC
            const double bx = geo_r[lane * 5], by = geo_r[lane * 5 + 1], ux = geo_r[lane * 5 + 2], uy = geo_r[lane * 5 + 3], off = geo_r[lane * 5 + 4];
#endif
            const uint32_t c1 = gcen_r[lane];
            uint32_t* rr_out = raw_row + lane * D4S;
            if (!slow) {
#if FSGM_FC_PIPE
                // software pipeline: the four gathers of a quad are consumed FSGM_FC_PIPE quads later, after the coordinates of the
                // following quads have been computed (the compiler alone keeps one quad in flight and waits for it)
                constexpr int LAG = FSGM_FC_PIPE < QPW ? FSGM_FC_PIPE : QPW;
                uint32_t g[LAG][4];
#pragma unroll
                for (int k = 0; k < QPW + LAG; ++k) {
                    const int qq = warp * QPW + k;
                    uint32_t idx[4] = {0, 0, 0, 0};
                    if (k < QPW) {
#pragma unroll
                        for (int j = 0; j < 4; ++j) {
                            const double t = __dmul_rn(off, vzs[4 * qq + j]);
                            const uint32_t xk = ref_round_clamp_k(__dadd_rn(bx, __dmul_rn(t, ux)), wk2);
                            const uint32_t yk = ref_round_clamp_k(__dadd_rn(by, __dmul_rn(t, uy)), hk2);
                            idx[j] = (yk >> 1) * Wu + (xk >> 1);
                        }
                    }
                    if (k >= LAG) {
                        const uint32_t* gg = g[k % LAG];
                        const uint32_t h0 = __popc(c1 ^ gg[0]), h1 = __popc(c1 ^ gg[1]), h2 = __popc(c1 ^ gg[2]), h3 = __popc(c1 ^ gg[3]);
                        rr_out[qq - LAG] = mad_u32(mad_u32(h3, 256u, h2), 65536u, mad_u32(h1, 256u, h0));
                    }
                    if (k < QPW) {
#pragma unroll
                        for (int j = 0; j < 4; ++j) g[k % LAG][j] = __ldg(c2w + idx[j]);
                    }
                }
#else
#pragma unroll
                for (int k = 0; k < QPW; ++k) {
                    const int qq = warp * QPW + k;
                    uint32_t packed = 0;
                    uint32_t hq[4];
                    (void)hq;
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        const double t = __dmul_rn(off, vzs[4 * qq + j]);
#if FSGM_FC_VARIANT == 0
                        const uint32_t x2 = ref_round_clamp_w(__dadd_rn(bx, __dmul_rn(t, ux)), wmax);
                        const uint32_t y2 = ref_round_clamp_w(__dadd_rn(by, __dmul_rn(t, uy)), hmax);
                        packed |= (uint32_t)__popc(c1 ^ __ldg(reinterpret_cast<const uint32_t*>(c2b + (y2 * W4 + x2 * 4u)))) << (8 * j);
#else
                        const uint32_t xk = ref_round_clamp_k(__dadd_rn(bx, __dmul_rn(t, ux)), wk2);
                        const uint32_t yk = ref_round_clamp_k(__dadd_rn(by, __dmul_rn(t, uy)), hk2);
                        hq[j] = (uint32_t)__popc(c1 ^ __ldg(c2w + ((yk >> 1) * Wu + (xk >> 1))));    // 32-bit word index: IMAD + IMAD.WIDE
#endif
                    }
#if FSGM_FC_VARIANT != 0
                    packed = mad_u32(mad_u32(hq[3], 256u, hq[2]), 65536u, mad_u32(hq[1], 256u, hq[0]));         // three IMADs (FMA pipe)
#endif
                    rr_out[qq] = packed;
                }
#endif
            } else {
                for (int k = 0; k < QPW; ++k) {
                    const int qq = warp * QPW + k;
                    uint32_t packed = 0;
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        const double t = __dmul_rn(off, vzs[4 * qq + j]);
                        const double wx = __dadd_rn(bx, __dmul_rn(t, ux)), wy = __dadd_rn(by, __dmul_rn(t, uy));
                        const uint32_t x2 = (wx != wx) ? 0u : ref_round_clamp_w(wx, wmax);
                        const uint32_t y2 = (wy != wy) ? 0u : ref_round_clamp_w(wy, hmax);
                        packed |= (uint32_t)__popc(c1 ^ __ldg(reinterpret_cast<const uint32_t*>(c2b + (y2 * W4 + x2 * 4u)))) << (8 * j);
                    }
                    rr_out[qq] = packed;
                }
            }
        }
        __syncthreads();
        // (2) horizontal 5-sums by a sliding window over this thread's XP consecutive columns; vertical 5-sums as
        //     running sums (add the new row, drop the row that leaves the window: it sits in the ring slot being rewritten)
        const int slot = (r + 10) % 5;
        const uint32_t* rr = raw_row + (i0 * XP) * D4S + q;
        uint32_t* hr = hring + (slot * FC_TX + i0 * XP) * D4 + q;
        uint32_t h = rr[0] + rr[D4S] + rr[2 * D4S] + rr[3 * D4S] + rr[4 * D4S];           // bytes <= 120: no carries
        const int yo = r - 2;
        const bool emit = yo >= y0;
        uint32_t* crow = Cout + ((size_t)yo * W + x0 + i0 * XP) * D4;
#if FSGM_FC_VARIANT == 0
#pragma unroll
        for (int k = 0; k < XP; ++k) {
            if (k) h = h - rr[(k - 1) * D4S] + rr[(k + 4) * D4S];
            const uint32_t old = hr[k * D4];
            hr[k * D4] = h;
            vlo[k] += (h & 0x00FF00FFu) - (old & 0x00FF00FFu);
            vhi[k] += ((h >> 8) & 0x00FF00FFu) - ((old >> 8) & 0x00FF00FFu);
            if (emit && x0 + i0 * XP + k < W) crow[k * D4] = box_norm4_fast(vlo[k], vhi[k]);
        }
#else
        // The running sums stay u16 pairs (one PRMT per half of h and of the leaving row, one IADD3 per pair).  Normalisation is ONE
        // fma per pair: the integer s < 1024 in a 16-bit half IS the fp16 subnormal s * 2^-24, and
        // fp16(2^20 / 25) * (s * 2^-24) + 64 rounds (to a multiple of 2^-4, the spacing at 64) to 64 + round(s / 25) / 16, pattern
        // 0x5400 | q — checked for every s <= 1023: the relative error of the constant times s stays below the 1/50 that separates
        // s / 25 from a rounding boundary.  FMA pipe instead of mask / shift / multiply chains on the ALU pipe.
        auto box_cols = [&](auto checked) {
#pragma unroll
            for (int k = 0; k < XP; ++k) {
                if (k) h = h - rr[(k - 1) * D4S] + rr[(k + 4) * D4S];
                const uint32_t old = hr[k * D4];
                hr[k * D4] = h;
                vlo[k] = vlo[k] + __byte_perm(h, 0u, 0x4240) - __byte_perm(old, 0u, 0x4240);
                vhi[k] = vhi[k] + __byte_perm(h, 0u, 0x4341) - __byte_perm(old, 0u, 0x4341);
                const uint32_t out = __byte_perm(h2fma(vlo[k], 0x791F791Fu, 0x54005400u), h2fma(vhi[k], 0x791F791Fu, 0x54005400u), 0x6240);
                if (emit && (!decltype(checked)::value || x0 + i0 * XP + k < W)) crow[k * D4] = out;
            }
        };
        if (x0 + FC_TX <= W) box_cols(std::false_type{}); else box_cols(std::true_type{});      // CTA-uniform: only the last strip tests columns
#endif
#if FSGM_FC_ASYNC
        asm volatile("cp.async.wait_all;" ::: "memory");
        __syncthreads();                              // next row's geometry has landed; everyone is done with raw_row
#else
        slow_mine = (tid < NPIX && more) ? store_geo(buf ^ 1, ga, gc) : 0;
        slow = __syncthreads_or(slow_mine);           // also: everyone is done with raw_row
#endif
    }
}

static size_t fused_cost_smem(int D4)
{
    const size_t tx = fc_tx(D4), npix = tx + 4;
    return (size_t)4 * D4 * 8 + 2 * npix * 5 * 8 + npix * (D4 + 1) * 4 + 5 * tx * D4 * 4 + 2 * npix * 4 + 64;
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
int launch_epi_cost_fused(fsgm_ctx* c, int n, const double* d_vz, const uint32_t* cen1, const uint32_t* cen2, int W, int H, int D,
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
        epi_cost_fused_kernel<D4V><<<grid, FC_THREADS, smem, c->stream>>>(cen1, cen2, Pd0, dirn, O, d_vz, W, H, ty, C); \
    } while (0)
    if (D4 == 16) FSGM_FC(16); else if (D4 == 32) FSGM_FC(32); else FSGM_FC(64);
#undef FSGM_FC
    FSGM_LAUNCHED(c);
    *done = true;
    return FSGM_OK;
}

}  // namespace fsgm
