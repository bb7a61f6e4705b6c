// Neighbour-guided variant — replaces calc_cost_based_on_hint(), sgm_step() and sgm2d() of the reference's
// calc_cost_sgm_ng.cpp (:122-186, :46-98, :188-419).
//
// Per pixel, 108 candidate flow vectors = 4 directions x (top-2 of that direction's ring slot + 1 random) x 3x3
// offsets; cost on the fly (5x5 census Hamming, clamped coordinates); one forward pass of 4-direction SGM over the
// explicit (mv, cost) entries with an O(108^2) compatibility search per direction; adaptive P2 (threshold 50);
// WTA -> flow.
//
// The data dependence is a single serial chain over all pixels in raster order: the hints of pixel p come from ring
// slots last written at p-2 (L1's two-slot ring is swapped every pixel and never reset per row, :247-248,:365-367 —
// so the first two pixels of a row are seeded by the last two of the previous row) and at (x, y-2) (row rings,
// :371-385), and the random hints consume libc rand() in raster order (:148-149).  There is no cross-pixel
// parallelism inside a pair, so one CTA walks one pair and pairs are spread over the SMs (batch parallel); inside
// a pixel the 2700 census taps and the 4 x 108 x 108 compatibility tests are spread over the CTA's threads.
// Integer-issue / latency bound; HBM traffic is negligible (the ring rows stay in L2).
#include "fsgm_internal.h"

namespace fsgm {

constexpr int NGD = 108;              // DIRECTION_NUM * (N + M) * MV_PER_HINT (:194)
constexpr int NG_THREADS = 512;
constexpr int NG_TP = 65;               // words per grid table of the pipelined kernel (64 + 1 pad)

struct Top2 { int mvx[2], mvy[2], cost[2]; };       // the two extra slots L[D], L[D+1] of the reference (:196)

struct NgWork {                       // per-pair global scratch, zero-initialised (:204-208)
    int*     mvrow;                   // [2][W][108][2]  candidate mvs of rows y (parity) and y-1
    int16_t* Lrow;                    // [3][2][W][108]  path costs of L2, L3, L4 for the two live rows
    Top2*    toprow;                  // [3][2][W]       top-2 slots of L2, L3, L4 (the reference's row rings)
};

struct NgParams {
    const uint8_t* I1; const uint32_t* cen1; const uint32_t* cen2;
    int* mvrow; int16_t* Lrow; Top2* toprow;            // bases; pair stride computed from W
    const int* rand_stream;                             // optional caller-supplied rand() stream, 8 per pixel
    const uint32_t* rng_state;                          // else: 31-word glibc TYPE_3 state per pair (after srand)
    int W, H, P1, P2;
    uint32_t* Sp32; int* Cent; uint32_t* minC; double* flow;
};

__device__ __forceinline__ int clampi(int v, int lo, int hi) { return min(max(v, lo), hi); }

__global__ void __launch_bounds__(NG_THREADS)
ng_kernel(const NgParams prm)
{
    const int W = prm.W, H = prm.H;
    const size_t N = (size_t)W * H;
    const int pair = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const uint8_t* I1 = prm.I1 + pair * N;
    const uint32_t* cen1 = prm.cen1 + pair * N;
    const uint32_t* cen2 = prm.cen2 + pair * N;
    int* mvrow = prm.mvrow + (size_t)pair * 2 * W * NGD * 2;
    int16_t* Lrow = prm.Lrow + (size_t)pair * 3 * 2 * W * NGD;
    Top2* toprow = prm.toprow + (size_t)pair * 3 * 2 * W;

    __shared__ int cmx[2][NGD], cmy[2][NGD];            // candidate mvs: [pixel parity] (cur / previous pixel)
    __shared__ int ccost[NGD];
    __shared__ int Lc[4][NGD];                          // this pixel's path costs, directions L1,L2,L3,L4
    // predecessor entries (mvx, mvy, path cost, -) of the four directions: [0] L1 = previous pixel, [1] L2 = (x-1,y-1),
    // [2] L3 = (x,y-1), [3] L4 = (x+1,y-1); one LDS.128 per compatibility test
    __shared__ int4 pent[4][NGD];
    __shared__ Top2 top1[2];                            // L1's two ring slots
    __shared__ Top2 tcur[4];                            // top-2 being built for this pixel
    __shared__ int preMin[4];
    __shared__ uint32_t c1win[25];
    __shared__ int rnd[8];
    __shared__ uint32_t rstate[31];
    __shared__ int rf, rr;

    if (tid < 2) { for (int i = 0; i < 2; ++i) { top1[tid].mvx[i] = 0; top1[tid].mvy[i] = 0; top1[tid].cost[i] = 0; } }
    if (tid < NGD) { pent[0][tid] = make_int4(0, 0, 0, 0); cmx[0][tid] = cmx[1][tid] = cmy[0][tid] = cmy[1][tid] = 0; }
    if (tid < 31 && !prm.rand_stream) rstate[tid] = prm.rng_state[pair * 31 + tid];
    if (tid == 0) { rf = 3; rr = 0; }                   // glibc TYPE_3: front = state + SEP_3, rear = state
    __syncthreads();

    for (int y = 0; y < H; ++y) {
        const int curRow = (y + 1) & 1, preRow = y & 1;            // the reference's row rings start Pre=0, Cur=1
        for (int x = 0; x < W; ++x) {
            const size_t p = (size_t)y * W + x;
            const int cp = (int)(p & 1);                            // parity buffer of this pixel's candidates
            const int cur1 = (int)((p + 1) & 1);                    // L1 ring: Cur slot of this pixel
            const bool startX = (x == 0), startY = (y == 0), startR = (x == W - 1);

            // ---- phase 0: random hints, census window, predecessor rows ------------------------------
            if (tid == 0) {
                if (prm.rand_stream) { for (int i = 0; i < 8; ++i) rnd[i] = prm.rand_stream[(pair * N + p) * 8 + i]; }
                else {
                    int f = rf, r = rr;
                    for (int i = 0; i < 8; ++i) {
                        uint32_t v = rstate[f] + rstate[r];
                        rstate[f] = v;
                        rnd[i] = (int)((v >> 1) & 0x7FFFFFFFu);
                        f = (f + 1 == 31) ? 0 : f + 1; r = (r + 1 == 31) ? 0 : r + 1;
                    }
                    rf = f; rr = r;
                }
            }
            if (tid >= 32 && tid < 57) {
                const int k = tid - 32;
                c1win[k] = cen1[(size_t)clampi(y + k / 5 - 2, 0, H - 1) * W + clampi(x + k % 5 - 2, 0, W - 1)];
            }
            if (!startY) {
                // predecessor slots: q=0 -> (x-1,y-1) for L2, q=1 -> (x,y-1) for L3, q=2 -> (x+1,y-1) for L4
                for (int i = tid; i < 3 * NGD; i += NG_THREADS) {
                    const int q = i / NGD, d = i - q * NGD, xs = x + q - 1;
                    if (xs >= 0 && xs < W) {
                        const size_t cell = (size_t)preRow * W + xs;
                        const int2 mv = *reinterpret_cast<const int2*>(mvrow + (cell * NGD + d) * 2);
                        pent[1 + q][d] = make_int4(mv.x, mv.y, (int)Lrow[((size_t)q * 2 * W + cell) * NGD + d], 0);
                    }
                }
            }
            __syncthreads();

            // ---- phase A: 108 candidates and their costs (2 threads per candidate) --------------------
            {
                // every thread runs the shuffle (warp 6 is only partly populated with candidates)
                const bool act = tid < 2 * NGD;
                const int c = act ? (tid >> 1) : 0, half = tid & 1;
                const int l = c / 27, i = (c % 27) / 9, off = c % 9, oy = off / 3 - 1, ox = off % 3 - 1;
                int hx, hy;
                if (i < 2) {
                    // hints are read from the CURRENT ring slot before it is overwritten (:276-277): stale data
                    const Top2& t = (l == 0) ? top1[cur1] : toprow[((size_t)(l - 1) * 2 + curRow) * W + x];
                    hx = t.mvx[i]; hy = t.mvy[i];
                } else {
                    hx = rnd[2 * l] % 256 - 128; hy = rnd[2 * l + 1] % 128 - 64;
                }
                uint32_t s = 0;
                const int k0 = half ? 13 : 0, k1 = act ? (half ? 25 : 13) : 0;
                for (int k = k0; k < k1; ++k) {
                    const int y1 = clampi(y + k / 5 - 2, 0, H - 1), x1 = clampi(x + k % 5 - 2, 0, W - 1);
                    const int y2 = clampi((int)((uint32_t)(oy + y1) + (uint32_t)hy), 0, H - 1);
                    const int x2 = clampi((int)((uint32_t)(ox + x1) + (uint32_t)hx), 0, W - 1);
                    s += __popc(c1win[k] ^ __ldg(cen2 + (size_t)W * y2 + x2));
                }
                s += __shfl_xor_sync(0xffffffffu, s, 1);
                if (act && !half) {
                    ccost[c] = (int)((2 * s + 25) / 50);             // (int)(1.0*s/25 + 0.5) (:177)
                    cmx[cp][c] = (int)((uint32_t)hx + (uint32_t)ox);
                    cmy[cp][c] = (int)((uint32_t)hy + (uint32_t)oy);
                }
            }
            // previous minima (the top-1 slot of each predecessor, as unsigned char, :53)
            if (tid == 224) preMin[0] = top1[cur1 ^ 1].cost[0] & 0xFF;
            if (tid >= 225 && tid < 228 && !startY) {
                const int q = tid - 225, xs = x + q - 1;
                if (xs >= 0 && xs < W) preMin[1 + q] = toprow[((size_t)q * 2 + preRow) * W + xs].cost[0] & 0xFF;
            }
            __syncthreads();

            // ---- phase B: the four path steps; work item = (direction, candidate) ----------------------
            const int pixCur = I1[p];
            for (int item = tid; item < 4 * NGD; item += NG_THREADS) {
                const int dir = item / NGD, d = item - dir * NGD;    // 0 L1, 1 L2, 2 L3, 3 L4
                const bool start = dir == 0 ? startX : dir == 1 ? (startX || startY) : dir == 2 ? startY : (startY || startR);
                int out;
                if (start) out = ccost[d];
                else {
                    const int4* q = pent[dir];
                    const int pixPre = dir == 0 ? I1[p - 1] : I1[p - W + (dir - 2)];
                    const int P2 = abs(pixCur - pixPre) > 50 ? prm.P2 / 8 : prm.P2;          // :101-105
                    const uint32_t pm = (uint32_t)preMin[dir];
                    const uint32_t far_ = (pm + (uint32_t)P2) & 0xFFu;
                    uint32_t same = far_, near_ = far_;
                    const int mx = cmx[cp][d], my = cmy[cp][d];
#pragma unroll 4
                    for (int d2 = 0; d2 < NGD; ++d2) {
                        // branch-free: data-dependent branches here cost more than the work they skip
                        const int4 e = q[d2];
                        const uint32_t c2 = (uint32_t)e.z;
                        const bool eq = (e.x == mx) & (e.y == my);
                        const bool nr = ((uint32_t)(e.x - mx + 2) <= 4u) & ((uint32_t)(e.y - my + 2) <= 4u);
                        same = eq ? (c2 & 0xFFu) : same;                                      // last match wins (:71-72)
                        const uint32_t cand = (nr & !eq) ? ((c2 + (uint32_t)prm.P1) & 0xFFu) : 0xFFu;
                        near_ = min(near_, cand);
                    }
                    out = ccost[d] + (int)min(min(far_, same), near_) - (int)pm;                // int, not truncated (:80)
                }
                Lc[dir][d] = out;
            }
            __syncthreads();

            // ---- phase C: top-2 per direction (warp `dir`), Sp + WTA (warp 4) ---------------------------
            if (warp < 4) {
                const int dir = warp;
                const bool start = dir == 0 ? startX : dir == 1 ? (startX || startY) : dir == 2 ? startY : (startY || startR);
                // the slot being overwritten keeps whatever it held (stale mvs are part of the reference's behaviour)
                Top2 old = (dir == 0) ? top1[cur1] : toprow[((size_t)(dir - 1) * 2 + curRow) * W + x];
                Top2 nw = old;
                if (start) nw.cost[0] = 0;                                                     // :281,:285,:291,...
                else {
                    // sequential strict-< insertion into two slots preset to 255 == the two smallest (cost, index) among
                    // the entries below 255; missing ones are padded with (255, stale mv of slot 0) / untouched slot 1
                    uint32_t k1 = 0xFFFFFFFFu, k2 = 0xFFFFFFFFu;
                    for (int d = lane; d < NGD; d += 32) {
                        const int cst = Lc[dir][d];
                        if (cst < 255) {
                            const uint32_t key = ((uint32_t)(cst + 1024) << 8) | (uint32_t)d;
                            if (key < k1) { k2 = k1; k1 = key; } else if (key < k2) k2 = key;
                        }
                    }
                    const uint32_t g1 = __reduce_min_sync(0xffffffffu, k1);
                    const uint32_t g2 = __reduce_min_sync(0xffffffffu, k1 == g1 ? k2 : k1);
                    nw.cost[0] = 255; nw.cost[1] = 255;
                    if (g1 != 0xFFFFFFFFu) {
                        const int d1 = g1 & 0xFF;
                        nw.mvx[1] = old.mvx[0]; nw.mvy[1] = old.mvy[0];                        // shifted-down preset slot
                        nw.mvx[0] = cmx[cp][d1]; nw.mvy[0] = cmy[cp][d1]; nw.cost[0] = Lc[dir][d1];
                        if (g2 != 0xFFFFFFFFu) {
                            const int d2 = g2 & 0xFF;
                            nw.mvx[1] = cmx[cp][d2]; nw.mvy[1] = cmy[cp][d2]; nw.cost[1] = Lc[dir][d2];
                        }
                    }
                }
                if (lane == 0) tcur[dir] = nw;
            } else if (warp == 4) {
                unsigned long long key = ~0ull;
                for (int d = lane; d < NGD; d += 32) {
                    const uint32_t a = (uint32_t)(Lc[0][d] + Lc[2][d]) + (uint32_t)(Lc[1][d] + Lc[3][d]);   // :357-362
                    if (prm.Sp32) prm.Sp32[(pair * N + p) * NGD + d] = a;
                    const unsigned long long kk = ((unsigned long long)a << 32) | (uint32_t)d;
                    key = kk < key ? kk : key;
                }
                for (int o = 16; o; o >>= 1) {
                    unsigned long long other = __shfl_xor_sync(0xffffffffu, key, o);
                    key = other < key ? other : key;
                }
                if (lane == 0) {
                    const int d = (int)(key & 0xFFFFFFFFu);
                    prm.minC[pair * N + p] = (uint32_t)(key >> 32);
                    prm.flow[(size_t)pair * 2 * N + p] = (double)cmx[cp][d];
                    prm.flow[(size_t)pair * 2 * N + N + p] = (double)cmy[cp][d];
                }
            } else if (prm.Cent) {
                for (int d = tid - 160; d < NGD; d += NG_THREADS - 160) {
                    int* e = prm.Cent + ((pair * N + p) * NGD + d) * 3;
                    e[0] = cmx[cp][d]; e[1] = cmy[cp][d]; e[2] = ccost[d];
                }
            }
            __syncthreads();

            // ---- phase D: commit this pixel's state to the rings -----------------------------------------
            {
                const size_t cell = (size_t)curRow * W + x;
                for (int i = tid; i < NGD; i += NG_THREADS) {
                    *reinterpret_cast<int2*>(mvrow + (cell * NGD + i) * 2) = make_int2(cmx[cp][i], cmy[cp][i]);
                    pent[0][i] = make_int4(cmx[cp][i], cmy[cp][i], Lc[0][i], 0);
                }
                for (int i = tid; i < 3 * NGD; i += NG_THREADS) {
                    const int q = i / NGD, d = i - q * NGD;
                    Lrow[((size_t)q * 2 * W + cell) * NGD + d] = (int16_t)Lc[q + 1][d];
                }
                if (tid == 0) top1[cur1] = tcur[0];
                if (tid >= 1 && tid < 4) toprow[((size_t)(tid - 1) * 2 + curRow) * W + x] = tcur[tid];
            }
            __syncthreads();
        }
    }
}


// ------------------------------------------------------------------------------------------------------------
// Pipelined variant (W >= 4).  Same arithmetic, two block barriers per pixel instead of five:
//   phase X: warps 0-13 run the 4 x 108 compatibility searches of pixel p (through per-grid tables, below); warp 14 builds the
//            18 candidates of pixel p+1 that hang on the L1 ring slot written by pixel p-1; the global rows pixel p+1 needs and
//            the row slots / census window of pixel p+3 are fetched into registers;
//   phase Y: top-2 per direction, Sp + WTA and the ring commits of pixel p; warps 12-14 build the 90 candidates of pixel p+2
//            whose hints are a row old (row ring slots, random hints); the prefetched rows go to shared memory.
// Candidates live in three buffers (pixel q: q % 3), the stale row slots in four (q % 4).
// ------------------------------------------------------------------------------------------------------------
// OCC = CTAs (pairs) resident per SM.  The walk is a chain of dependent phases separated by block barriers, so one CTA leaves most
// issue slots empty; a second pair on the same SM fills them.  OCC = 2 fits in 64 registers per thread; OCC = 3 caps the kernel
// at 40 registers and spills (slower than 2, kept for A/B through fsgm_tune key 3).
template <int OCC>
__global__ void __launch_bounds__(NG_THREADS, OCC)
ng_pipe_kernel(const NgParams prm)
{
    const int W = prm.W, H = prm.H;
    const size_t N = (size_t)W * H;
    const int pair = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const uint8_t* I1 = prm.I1 + pair * N;
    const uint32_t* cen1 = prm.cen1 + pair * N;
    const uint32_t* cen2 = prm.cen2 + pair * N;
    int* mvrow = prm.mvrow + (size_t)pair * 2 * W * NGD * 2;
    int16_t* Lrow = prm.Lrow + (size_t)pair * 3 * 2 * W * NGD;
    Top2* toprow = prm.toprow + (size_t)pair * 3 * 2 * W;

    __shared__ int cmx[3][NGD], cmy[3][NGD], ccost[3][NGD];   // candidates of pixel q: buffer q % 3
    __shared__ int Lc[4][NGD];
    __shared__ int2 pxy[4][NGD];                              // predecessor entries of the pixel being stepped: flow vectors
    __shared__ int pz[4][NGD];                                //                                                 path costs
    // per predecessor and 3x3 candidate grid (12 per predecessor): the answer of the compatibility search for every displacement
    // of a candidate against the grid corner, T[dir][grid][uy * 8 + ux] with (ux, uy) = candidate - corner + 2 clamped to 7
    // (row / column 7 = "no cell within +-2": 0xFF, written once); low half = smallest P1 term among the compatible cells, bits
    // 16-23 = cost of the cell of equal flow, bit 24 = such a cell exists
    __shared__ uint32_t T[4][12][NG_TP];                      // pitch 65 words: the builders of neighbouring grids hit different banks
    __shared__ int2 corner[4][12];                            // grid corner - 2
    __shared__ Top2 top1[2];
    __shared__ Top2 hint[4][3];                               // stale ring slots of L2,L3,L4 for pixel q (slot q % 4): its hints, and the
                                                              // content its own top-2 commit starts from
    __shared__ int preMin[4];
    __shared__ uint32_t c1win[3][25];                         // 5 x 5 census window of pixel q: q % 3
    __shared__ int rnd[3][8];
    __shared__ uint32_t rstate[31];
    __shared__ int rf, rr;

    auto gen_rnd = [&](size_t q, int q3) {                    // one thread: the 8 rand() values of pixel q
        int* out = rnd[q3];
        if (prm.rand_stream) { for (int i = 0; i < 8; ++i) out[i] = prm.rand_stream[(pair * N + q) * 8 + i]; return; }
        int f = rf, r = rr;
        for (int i = 0; i < 8; ++i) {
            const uint32_t v = rstate[f] + rstate[r];
            rstate[f] = v;
            out[i] = (int)((v >> 1) & 0x7FFFFFFFu);
            f = (f + 1 == 31) ? 0 : f + 1; r = (r + 1 == 31) ? 0 : r + 1;
        }
        rf = f; rr = r;
    };
    auto load_c1 = [&](int x, int y, int k) {
        return cen1[(size_t)clampi(y + k / 5 - 2, 0, H - 1) * W + clampi(x + k % 5 - 2, 0, W - 1)];
    };
    // candidate c of pixel q = (x, y): hint, 25 Hamming taps, entry (mv, cost) into buffer q3 = q % 3; q4 = q % 4
    auto make_candidate = [&](int c, size_t q, int q3, int q4, int x, int y) {
        const int cur1q = (int)((q + 1) & 1);
        const bool pin = x >= 2 && x + 2 < W && y >= 2 && y + 2 < H;          // the pixel's own 5 x 5 window is not clamped
        const int l = c / 27, i = (c % 27) / 9, off = c % 9, oy = off / 3 - 1, ox = off % 3 - 1;
        int hx, hy;
        if (i < 2) {
            const Top2& tt = (l == 0) ? top1[cur1q] : hint[q4][l - 1];
            hx = tt.mvx[i]; hy = tt.mvy[i];
        } else { hx = rnd[q3][2 * l] % 256 - 128; hy = rnd[q3][2 * l + 1] % 128 - 64; }
        uint32_t s = 0;
        const int bx = (int)((uint32_t)(x - 2 + ox) + (uint32_t)hx), by = (int)((uint32_t)(y - 2 + oy) + (uint32_t)hy);
        if (pin && bx >= 0 && bx < W - 4 && by >= 0 && by < H - 4) {          // nor is the displaced one: 25 loads at fixed offsets
            const uint32_t* r = cen2 + (size_t)by * W + bx;
#pragma unroll
            for (int ky = 0; ky < 5; ++ky, r += W)
#pragma unroll
                for (int kx = 0; kx < 5; ++kx) s += __popc(c1win[q3][ky * 5 + kx] ^ __ldg(r + kx));
        } else {                                                              // clamped twice (:161-166): five columns, five rows
            int xo[5];
#pragma unroll
            for (int kx = 0; kx < 5; ++kx) xo[kx] = clampi((int)((uint32_t)(ox + clampi(x + kx - 2, 0, W - 1)) + (uint32_t)hx), 0, W - 1);
#pragma unroll
            for (int ky = 0; ky < 5; ++ky) {
                const int y2 = clampi((int)((uint32_t)(oy + clampi(y + ky - 2, 0, H - 1)) + (uint32_t)hy), 0, H - 1);
                const uint32_t* r = cen2 + (size_t)W * y2;
#pragma unroll
                for (int kx = 0; kx < 5; ++kx) s += __popc(c1win[q3][ky * 5 + kx] ^ __ldg(r + xo[kx]));
            }
        }
        ccost[q3][c] = (int)((2 * s + 25) / 50);
        cmx[q3][c] = (int)((uint32_t)hx + (uint32_t)ox);
        cmy[q3][c] = (int)((uint32_t)hy + (uint32_t)oy);
    };
    constexpr int NFRESH = 18;                                 // candidates 0..17: the two L1 hints (l = 0, i < 2) x 9 offsets

    // ---- prologue: everything pixels 0, 1 and 2 need ---------------------------------------------------------
    if (tid < 2) { for (int i = 0; i < 2; ++i) { top1[tid].mvx[i] = 0; top1[tid].mvy[i] = 0; top1[tid].cost[i] = 0; } }
    if (tid < 12) { Top2 z; for (int i = 0; i < 2; ++i) { z.mvx[i] = z.mvy[i] = z.cost[i] = 0; } hint[tid / 3][tid % 3] = z; }   // ring rows start zeroed
    if (tid < NGD) { pxy[0][tid] = make_int2(0, 0); pz[0][tid] = 0; }
    if (tid < 31 && !prm.rand_stream) rstate[tid] = prm.rng_state[pair * 31 + tid];
    if (tid == 0) { rf = 3; rr = 0; }
    if (tid >= 64 && tid < 64 + 75) {
        const int q = (tid - 64) / 25, k = (tid - 64) % 25;
        if ((size_t)q < N) c1win[q][k] = load_c1(q % W, q / W, k);
    }
    for (int i = tid; i < 4 * 12 * NG_TP; i += NG_THREADS) (&T[0][0][0])[i] = 0xFFu;
    __syncthreads();
    if (tid == 0) { for (int q = 0; q < 3 && (size_t)q < N; ++q) gen_rnd(q, q); }
    __syncthreads();
    if (tid < NGD) make_candidate(tid, 0, 0, 0, 0, 0);
    else if (tid >= 128 && tid < 128 + NGD - NFRESH && N > 1) make_candidate(tid - 128 + NFRESH, 1, 1, 1, 1 % W, 1 / W);
    __syncthreads();

    int x = 0, y = 0, p3 = 0, p4 = 0;                         // p3 = p % 3, p4 = p % 4
    const uint32_t Nn = (uint32_t)N;                          // N < 2^31 (checked at launch): 32-bit pixel counters
    const int pfq = tid / NGD, pfd = tid - pfq * NGD;         // prefetch role of threads 0..323: (row ring q, entry d)
    const uint32_t pf_L_base = (uint32_t)pfq * 2u * (uint32_t)W * NGD + (uint32_t)pfd;
    for (uint32_t p = 0; p < Nn; ++p) {
        const int curRow = (y + 1) & 1;
        const int cur1 = (int)((p + 1) & 1);
        const bool startX = (x == 0), startY = (y == 0), startR = (x == W - 1);
        const uint32_t p1 = p + 1, p2 = p + 2, p3n = p + 3;
        const int x1n = x + 1 == W ? 0 : x + 1, y1n = x + 1 == W ? y + 1 : y;
        const int x2n = x1n + 1 == W ? 0 : x1n + 1, y2n = x1n + 1 == W ? y1n + 1 : y1n;
        const int x3n = x2n + 1 == W ? 0 : x2n + 1, y3n = x2n + 1 == W ? y2n + 1 : y2n;
        const int cp = p3;                                    // candidate buffer of this pixel

        // ---- phase X ----------------------------------------------------------------------------------------------
        // register prefetch (consumed in phase Y): predecessor rows + previous minima of p+1, stale row slots + census of p+3
        int2 pf_mv = make_int2(0, 0); int pf_c = 0; bool pf_ok = false;
        Top2 pf_hint; uint32_t pf_c1 = 0; int pf_min = 0;
        if (p1 < Nn && y1n > 0) {
            const int preRow1 = y1n & 1;
            if (tid < 3 * NGD) {
                const int xs = x1n + pfq - 1;
                if ((unsigned)xs < (unsigned)W) {
                    const uint32_t cell = (uint32_t)(preRow1 * W + xs);
                    pf_mv = *reinterpret_cast<const int2*>(mvrow + (cell * NGD + (uint32_t)pfd) * 2u);
                    pf_c = (int)Lrow[cell * NGD + pf_L_base];
                    pf_ok = true;
                }
            } else if (tid >= 330 && tid < 333) {
                const int q = tid - 330, xs = x1n + q - 1;
                if (xs >= 0 && xs < W) pf_min = toprow[((size_t)q * 2 + preRow1) * W + xs].cost[0] & 0xFF;
            }
        }
        if (p3n < Nn) {
            if (tid >= 324 && tid < 327) pf_hint = toprow[((size_t)(tid - 324) * 2 + ((y3n + 1) & 1)) * W + x3n];
            if (tid >= 352 && tid < 377) pf_c1 = load_c1(x3n, y3n, tid - 352);
        }
        if (warp < 14) {
            // A predecessor's 108 entries are twelve 3 x 3 grids of consecutive flow vectors (hint + offset, offy outer).  For a
            // candidate (mx, my) and a grid with corner (X0, Y0) the cell of EQUAL flow is (ex, ey) = (mx - X0, my - Y0); the cells
            // within +-2 (the P1 term, :73-74) are [ex-2, ex+2] x [ey-2, ey+2] clipped to the grid — the whole grid minus the equal
            // cell when that cell lies inside, a rectangle that touches a grid edge otherwise.  The answer depends only on
            // (ux, uy) = (ex + 2, ey + 2) in 0..6 (anything else: no compatible cell), so it is tabulated per grid by 336 threads
            // (one per direction, grid and uy) and the search is ONE table look-up per grid: 12 per candidate and direction
            // instead of 108 entry tests.  Values are the reference's own per-entry terms ((cost + P1) & 255, cost & 255), so no
            // parameter domain is excluded.
            if (tid < 336) {
                const int uy = tid / 48, r = tid - uy * 48, dir = r / 12, g = r - dir * 12;
                const int* e = pz[dir] + g * 9;
                uint32_t z[9], a[9];
#pragma unroll
                for (int c = 0; c < 9; ++c) { z[c] = (uint32_t)e[c]; a[c] = (z[c] + (uint32_t)prm.P1) & 0xFFu; }
                if (uy == 0) { const int2 c0 = pxy[dir][g * 9]; corner[dir][g] = make_int2(c0.x - 2, c0.y - 2); }
                uint32_t* t = &T[dir][g][uy * 8];
                // rows [uy-4, uy] clipped to [0, 2]
                const bool r0 = uy <= 4, r1 = uy >= 1 && uy <= 5, r2 = uy >= 2;
                uint32_t c0 = 255u, c1 = 255u, c2 = 255u;                          // column minima over those rows
                if (r0) { c0 = a[0]; c1 = a[1]; c2 = a[2]; }
                if (r1) { c0 = min(c0, a[3]); c1 = min(c1, a[4]); c2 = min(c2, a[5]); }
                if (r2) { c0 = min(c0, a[6]); c1 = min(c1, a[7]); c2 = min(c2, a[8]); }
                const uint32_t m01 = min(c0, c1), m12 = min(c1, c2);
                t[0] = c0; t[1] = m01; t[5] = m12; t[6] = c2;                       // columns [ux-4, ux] clipped to [0, 2]
                if (uy < 2 || uy > 4) { const uint32_t m = min(m01, c2); t[2] = m; t[3] = m; t[4] = m; }
                else {
                    // the equal cell lies in row uy - 2: the whole grid minus that cell, and the cell's own cost
                    const int rr_ = uy - 2;
                    const uint32_t f0 = min(min(a[0], a[1]), a[2]), f1 = min(min(a[3], a[4]), a[5]), f2 = min(min(a[6], a[7]), a[8]);
                    const uint32_t o = rr_ == 0 ? min(f1, f2) : rr_ == 1 ? min(f0, f2) : min(f0, f1);
                    const uint32_t b0 = rr_ == 0 ? a[0] : rr_ == 1 ? a[3] : a[6], b1 = rr_ == 0 ? a[1] : rr_ == 1 ? a[4] : a[7],
                                   b2 = rr_ == 0 ? a[2] : rr_ == 1 ? a[5] : a[8];
                    const uint32_t s0 = rr_ == 0 ? z[0] : rr_ == 1 ? z[3] : z[6], s1 = rr_ == 0 ? z[1] : rr_ == 1 ? z[4] : z[7],
                                   s2 = rr_ == 0 ? z[2] : rr_ == 1 ? z[5] : z[8];
                    t[2] = min(o, min(b1, b2)) | ((s0 & 0xFFu) << 16) | 0x01000000u;
                    t[3] = min(o, min(b0, b2)) | ((s1 & 0xFFu) << 16) | 0x01000000u;
                    t[4] = min(o, min(b0, b1)) | ((s2 & 0xFFu) << 16) | 0x01000000u;
                }
            }
            asm volatile("bar.sync 1, 448;" ::: "memory");         // the fourteen warps of this phase only
            const int pixCur = I1[p];
            if (tid < 4 * NGD) {
                const int dir = tid / NGD, d = tid - dir * NGD;    // 0 L1, 1 L2, 2 L3, 3 L4
                const bool start = dir == 0 ? startX : dir == 1 ? (startX || startY) : dir == 2 ? startY : (startY || startR);
                int out;
                if (start) out = ccost[cp][d];
                else {
                    const int pixPre = dir == 0 ? I1[p - 1] : I1[p - W + (dir - 2)];
                    const int P2 = abs(pixCur - pixPre) > 50 ? prm.P2 / 8 : prm.P2;          // :101-105
                    const uint32_t pm = (uint32_t)preMin[dir];
                    const uint32_t far_ = (pm + (uint32_t)P2) & 0xFFu;
                    uint32_t acc = far_, se = 0;                                               // low half of acc: the running P1 minimum
                    const int mx = cmx[cp][d], my = cmy[cp][d];
                    const uint32_t* Td = &T[dir][0][0];
                    const int2* cr = corner[dir];
#pragma unroll
                    for (int g = 0; g < 12; ++g) {
                        const int2 c0 = cr[g];
                        const uint32_t ux = min((uint32_t)(mx - c0.x), 7u), uy = min((uint32_t)(my - c0.y), 7u);
                        const uint32_t e = Td[g * NG_TP + uy * 8 + ux];
                        acc = __vminu2(acc, e);
                        se = e >= 0x01000000u ? e : se;                                        // later grids overwrite: last match wins (:71-72)
                    }
                    const uint32_t near_ = acc & 0xFFFFu, same = se ? (se >> 16) & 0xFFu : far_;
                    out = ccost[cp][d] + (int)min(min(far_, same), near_) - (int)pm;           // int, not truncated (:80)
                }
                Lc[dir][d] = out;
            }
        } else if (warp == 14 && lane < NFRESH && p1 < Nn) {
            // the L1 ring slot pixel p+1 takes its hints from was written by pixel p-1 (:276-277): only these wait for it
            make_candidate(lane, p1, p3 == 2 ? 0 : p3 + 1, (p4 + 1) & 3, x1n, y1n);
        }
        __syncthreads();

        // ---- phase Y ----------------------------------------------------------------------------------------------
        if (warp < 4) {
            const int dir = warp;
            const bool start = dir == 0 ? startX : dir == 1 ? (startX || startY) : dir == 2 ? startY : (startY || startR);
            Top2* slot = (dir == 0) ? &top1[cur1] : &toprow[((size_t)(dir - 1) * 2 + curRow) * W + x];
            // stale content is part of the reference's behaviour; the row slots were fetched three pixels ago (hint[])
            const Top2 old = (dir == 0) ? top1[cur1] : hint[p4][dir - 1];
            Top2 nw = old;
            if (start) nw.cost[0] = 0;
            else {
                uint32_t k1 = 0xFFFFFFFFu, k2 = 0xFFFFFFFFu;
                for (int d = lane; d < NGD; d += 32) {
                    const int cst = Lc[dir][d];
                    if (cst < 255) {
                        const uint32_t key = ((uint32_t)(cst + 1024) << 8) | (uint32_t)d;
                        if (key < k1) { k2 = k1; k1 = key; } else if (key < k2) k2 = key;
                    }
                }
                const uint32_t g1 = __reduce_min_sync(0xffffffffu, k1);
                const uint32_t g2 = __reduce_min_sync(0xffffffffu, k1 == g1 ? k2 : k1);
                nw.cost[0] = 255; nw.cost[1] = 255;
                if (g1 != 0xFFFFFFFFu) {
                    const int d1 = g1 & 0xFF;
                    nw.mvx[1] = old.mvx[0]; nw.mvy[1] = old.mvy[0];
                    nw.mvx[0] = cmx[cp][d1]; nw.mvy[0] = cmy[cp][d1]; nw.cost[0] = Lc[dir][d1];
                    if (g2 != 0xFFFFFFFFu) {
                        const int d2 = g2 & 0xFF;
                        nw.mvx[1] = cmx[cp][d2]; nw.mvy[1] = cmy[cp][d2]; nw.cost[1] = Lc[dir][d2];
                    }
                }
            }
            __syncwarp();
            if (lane == 0) {
                *slot = nw;
                if (dir == 0) preMin[0] = nw.cost[0] & 0xFF;          // L1's predecessor of pixel p+1 is this pixel
            }
        } else if (warp == 4) {
            unsigned long long key = ~0ull;
            for (int d = lane; d < NGD; d += 32) {
                const uint32_t a = (uint32_t)(Lc[0][d] + Lc[2][d]) + (uint32_t)(Lc[1][d] + Lc[3][d]);   // :357-362
                if (prm.Sp32) prm.Sp32[(pair * N + p) * NGD + d] = a;
                const unsigned long long kk = ((unsigned long long)a << 32) | (uint32_t)d;
                key = kk < key ? kk : key;
            }
            for (int o = 16; o; o >>= 1) {
                const unsigned long long other = __shfl_xor_sync(0xffffffffu, key, o);
                key = other < key ? other : key;
            }
            if (lane == 0) {
                const int d = (int)(key & 0xFFFFFFFFu);
                prm.minC[pair * N + p] = (uint32_t)(key >> 32);
                prm.flow[(size_t)pair * 2 * N + p] = (double)cmx[cp][d];
                prm.flow[(size_t)pair * 2 * N + N + p] = (double)cmy[cp][d];
            }
        } else if (warp >= 5 && warp <= 8) {
            // ring commits of pixel p: candidate mvs + L2,L3,L4 costs to the row rings, L1 entries for the next pixel
            const size_t cell = (size_t)curRow * W + x;
            for (int i = tid - 160; i < NGD; i += 128) {
                *reinterpret_cast<int2*>(mvrow + (cell * NGD + i) * 2) = make_int2(cmx[cp][i], cmy[cp][i]);
                pxy[0][i] = make_int2(cmx[cp][i], cmy[cp][i]); pz[0][i] = Lc[0][i];
#pragma unroll
                for (int q = 0; q < 3; ++q) Lrow[((size_t)q * 2 * W + cell) * NGD + i] = (int16_t)Lc[q + 1][i];
                if (prm.Cent) {
                    int* e = prm.Cent + ((pair * N + p) * NGD + i) * 3;
                    e[0] = cmx[cp][i]; e[1] = cmy[cp][i]; e[2] = ccost[cp][i];
                }
            }
            if (tid == 287 && p3n < Nn) gen_rnd(p3n, p3);                                   // (p + 3) % 3
        } else if (warp >= 12 && tid < 384 + NGD - NFRESH && p2 < Nn) {
            // the 90 candidates of pixel p+2 whose hints are a row old (fetched during pixel p-1) or random
            make_candidate(tid - 384 + NFRESH, p2, p3 == 0 ? 2 : p3 - 1, (p4 + 2) & 3, x2n, y2n);
        }
        // prefetched rows -> shared memory (pxy / pz[1..3] were last read in phase X of this pixel)
        if (pf_ok) { pxy[1 + pfq][pfd] = pf_mv; pz[1 + pfq][pfd] = pf_c; }
        if (tid >= 330 && tid < 333) preMin[1 + tid - 330] = pf_min;
        if (p3n < Nn) {
            if (tid >= 324 && tid < 327) hint[(p4 + 3) & 3][tid - 324] = pf_hint;
            if (tid >= 352 && tid < 377) c1win[p3][tid - 352] = pf_c1;                    // (p + 3) % 3
        }
        __syncthreads();
        x = x1n; y = y1n; p3 = p3 == 2 ? 0 : p3 + 1; p4 = (p4 + 1) & 3;
    }
}

size_t ng_scratch_bytes(int n, int W)
{
    return align256((size_t)n * 2 * W * NGD * 2 * sizeof(int)) + align256((size_t)n * 3 * 2 * W * NGD * sizeof(int16_t)) +
           align256((size_t)n * 3 * 2 * W * sizeof(Top2)) + align256((size_t)n * 31 * 4) + 1024;
}

int launch_ng(fsgm_ctx* c, int n, const uint8_t* I1, const uint32_t* cen1, const uint32_t* cen2, int W, int H, int P1, int P2,
              const uint32_t* host_rng_states /* n x 31, host */, const int* d_rand_stream,
              uint32_t* Sp32, int* Cent, uint32_t* minC, double* flow)
{
    StageScope ss(c, ST_NG);
    NgParams p{};
    p.I1 = I1; p.cen1 = cen1; p.cen2 = cen2; p.W = W; p.H = H; p.P1 = P1; p.P2 = P2;
    p.Sp32 = Sp32; p.Cent = Cent; p.minC = minC; p.flow = flow; p.rand_stream = d_rand_stream;
    uint32_t* d_state = nullptr;
    const size_t b_mv = (size_t)n * 2 * W * NGD * 2 * sizeof(int), b_L = (size_t)n * 3 * 2 * W * NGD * sizeof(int16_t),
                 b_top = (size_t)n * 3 * 2 * W * sizeof(Top2);
    FSGM_TRY(arena_alloc(c, b_mv, (void**)&p.mvrow));
    FSGM_TRY(arena_alloc(c, b_L, (void**)&p.Lrow));
    FSGM_TRY(arena_alloc(c, b_top, (void**)&p.toprow));
    FSGM_TRY(arena_alloc(c, (size_t)n * 31 * 4, (void**)&d_state));
    FSGM_CUDA(c, cudaMemsetAsync(p.mvrow, 0, b_mv, c->stream));
    FSGM_CUDA(c, cudaMemsetAsync(p.Lrow, 0, b_L, c->stream));
    FSGM_CUDA(c, cudaMemsetAsync(p.toprow, 0, b_top, c->stream));
    if (!d_rand_stream) {
        FSGM_CUDA(c, cudaMemcpyAsync(d_state, host_rng_states, (size_t)n * 31 * 4, cudaMemcpyHostToDevice, c->stream));
        FSGM_CUDA(c, cudaStreamSynchronize(c->stream));       // host_rng_states is a caller temporary
    }
    p.rng_state = d_state;
    if (W >= 4 && (size_t)W * H < ((size_t)1 << 31)) {  // pipelined phases need p+1 .. p+3 to lie outside the cells pixel p commits; 32-bit pixel counters
        // resident pairs per SM: as many as the batch can use (a single pair runs fastest with all the registers)
        // measured on 1242 x 48 strips: 108 / 152 / 146 pairs/s at 1 / 2 / 3 pairs per SM (the third costs spills at 40 registers)
        const int occ = c->ng_occupancy > 0 ? c->ng_occupancy : (n > c->sm_count ? 2 : 1);
        if (occ >= 3) ng_pipe_kernel<3><<<n, NG_THREADS, 0, c->stream>>>(p);
        else if (occ == 2) ng_pipe_kernel<2><<<n, NG_THREADS, 0, c->stream>>>(p);
        else ng_pipe_kernel<1><<<n, NG_THREADS, 0, c->stream>>>(p);
    }
    else ng_kernel<<<n, NG_THREADS, 0, c->stream>>>(p);
    FSGM_LAUNCHED(c);
    return FSGM_OK;
}

}  // namespace fsgm
