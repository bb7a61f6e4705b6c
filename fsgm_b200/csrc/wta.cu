// Winner-take-all + subpixel + vz-index -> disparity — replaces the tail of sgm()
// (reference calc_cost_sgm.cpp:259-308) and convert_vzInd_to_disp (:414-426).
//
// Sp(p,d) = sum_r L_r(p,d) is formed on the fly from the per-direction u8 volumes (1 B read per voxel and
// direction; nothing is written back unless the caller asks for the u16 Sp volume for stage parity).
// Reference behaviours reproduced on purpose:
//   * argmin = FIRST minimum (strict <, :267)            -> key = (sum << 16 | label), warp min
//   * refinement only for 1 < idx (label 1 is never refined, :293)
//   * idx == D-1 reads Sp[p][D], i.e. the NEXT pixel's label 0 (:296); past the last pixel the reference
//     reads out of bounds — defined here as 0, which is what a zero-padded allocation gives
//   * bestD = (unsigned)(sub*256), then the vz conversion truncates a second time (:303,:423), both with
//     x86 double->unsigned semantics (through a signed 64-bit conversion)
// fp64 uses explicit round-to-nearest ops so nothing is contracted into an FMA.
#include "fsgm_internal.h"

namespace fsgm {

constexpr int WTA_WARPS = 8;
constexpr int WTA_MAXD = 512;

__device__ __forceinline__ uint32_t x86_d2u(double v)
{
    if (!(v > -9223372036854775809.0 && v < 9223372036854775808.0)) return 0u;
    return (uint32_t)(unsigned long long)__double2ll_rz(v);
}

struct WtaParams {
    const uint8_t* L[8];
    int n_dirs;
    int W, H, D;
    int subpixel, vz_to_disp;
    const double* O;
    double vMax;
    uint16_t* Sp16;
    const uint16_t* Sp_in;       // if set: sums come from this u16 volume instead of the L_r volumes (direction-split path)
    const uint16_t* next0;       // if set: Sp[.][0] of the pixel after this slab (the reference's read past label D-1)
    uint32_t* bestD;
    uint32_t* minC;
};

template <bool VEC8>
__global__ void __launch_bounds__(WTA_WARPS * 32)
epi_wta_kernel(const WtaParams prm)
{
    __shared__ uint16_t sums[WTA_WARPS][WTA_MAXD + 8];
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    const int D = prm.D, R = prm.n_dirs;
    const size_t N = (size_t)prm.W * prm.H;
    const size_t vol = blockIdx.y * N * D;
    const size_t group = ((size_t)blockIdx.x * WTA_WARPS + wib) * 32;       // 32 consecutive pixels per warp
    if (group >= N) return;
    uint16_t* s = sums[wib];

    uint32_t my_idx = 0, my_min = 0;
    double my_c_1 = 0, my_c = 0, my_c1 = 0;
    bool my_refine = false;

    for (int i = 0; i < 32; ++i) {
        const size_t p = group + i;
        if (p >= N) break;
        uint32_t key = 0xFFFFFFFFu;
        if (prm.Sp_in) {
            for (int d = lane; d < D; d += 32) {
                const uint32_t a = __ldg(prm.Sp_in + vol + p * D + d);
                s[d] = (uint16_t)a;
                key = min(key, (a << 16) | (uint32_t)d);
            }
        } else if (VEC8) {
            for (int d0 = lane * 8; d0 < D; d0 += 256) {
                uint32_t a0 = 0, a1 = 0, a2 = 0, a3 = 0;                   // u16x2 accumulators for 8 labels
                for (int r = 0; r < R; ++r) {
                    uint2 v = __ldg(reinterpret_cast<const uint2*>(prm.L[r] + vol + p * D + d0));
                    a0 += __byte_perm(v.x, 0, 0x4140); a1 += __byte_perm(v.x, 0, 0x4342);
                    a2 += __byte_perm(v.y, 0, 0x4140); a3 += __byte_perm(v.y, 0, 0x4342);
                }
                uint32_t* sw = reinterpret_cast<uint32_t*>(s + d0);
                sw[0] = a0; sw[1] = a1; sw[2] = a2; sw[3] = a3;
                if (prm.Sp16) {
                    uint4* o = reinterpret_cast<uint4*>(prm.Sp16 + vol + p * D + d0);
                    *o = make_uint4(a0, a1, a2, a3);
                }
                uint32_t acc[4] = {a0, a1, a2, a3};
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    key = min(key, ((acc[j] & 0xFFFFu) << 16) | (uint32_t)(d0 + 2 * j));
                    key = min(key, (acc[j] & 0xFFFF0000u) | (uint32_t)(d0 + 2 * j + 1));
                }
            }
        } else {
            for (int d = lane; d < D; d += 32) {
                uint32_t a = 0;
                for (int r = 0; r < R; ++r) a += __ldg(prm.L[r] + vol + p * D + d);
                s[d] = (uint16_t)a;
                if (prm.Sp16) prm.Sp16[vol + p * D + d] = (uint16_t)a;
                key = min(key, (a << 16) | (uint32_t)d);
            }
        }
        key = __reduce_min_sync(0xffffffffu, key);
        __syncwarp();
        const uint32_t idx = key & 0xFFFFu, best = key >> 16;
        const bool refine = prm.subpixel && idx > 1;
        uint32_t c_1 = 0, c1 = 0;
        if (refine) {                                    // warp-uniform
            c_1 = s[idx - 1];
            if (idx + 1 < (uint32_t)D) c1 = s[idx + 1];
            else if (prm.Sp_in) {
                c1 = (p + 1 < N) ? __ldg(prm.Sp_in + vol + (p + 1) * D) : (prm.next0 ? __ldg(prm.next0) : 0u);
            } else if (p + 1 < N) {
                uint32_t v = (lane < R) ? __ldg(prm.L[lane] + vol + (p + 1) * D) : 0u;
                c1 = __reduce_add_sync(0xffffffffu, v);
            } else c1 = prm.next0 ? __ldg(prm.next0) : 0u;
        }
        if (lane == i) { my_idx = idx; my_min = best; my_c_1 = c_1; my_c = best; my_c1 = c1; my_refine = refine; }
        __syncwarp();
    }

    const size_t p = group + lane;
    if (p >= N) return;
    uint32_t q = my_idx;                                 // raw label, or Q24.8 when subpixel is on
    if (prm.subpixel) {
        if (my_refine) {
            double num = __dsub_rn(my_c1, my_c_1);
            double den = (my_c1 < my_c_1) ? __dsub_rn(my_c, my_c_1) : __dsub_rn(my_c, my_c1);
            double sub = __dadd_rn((double)my_idx, __ddiv_rn(__ddiv_rn(num, den), 2.0));
            q = x86_d2u(__dmul_rn(sub, 256.0));
        } else q = my_idx * 256u;
    }
    if (prm.vz_to_disp) {                                // the reference converts whatever sgm() left in bestD (:593)
        double d = __ddiv_rn((double)q, 256.0);
        double rr = __dmul_rn(__ddiv_rn(d, (double)(D + 1)), prm.vMax);
        double vz = __ddiv_rn(rr, __dsub_rn(1.0, rr));
        q = x86_d2u(__dmul_rn(__dmul_rn(prm.O[blockIdx.y * N + p], vz), 256.0));
    }
    prm.bestD[blockIdx.y * N + p] = q;
    prm.minC[blockIdx.y * N + p] = my_min;
}

int launch_epi_wta(fsgm_ctx* c, int n, uint8_t* const* Lvols, int n_dirs, int W, int H, int D, int subpixel,
                   int vz_to_disp, const double* O, double vMax, uint16_t* Sp16, uint32_t* bestD, uint32_t* minC)
{
    if (D > WTA_MAXD) return fail(c, FSGM_ERR_DOMAIN, "label count must be <= 512");
    StageScope ss(c, ST_WTA);
    WtaParams p{};
    for (int k = 0; k < n_dirs; ++k) p.L[k] = Lvols[k];
    p.n_dirs = n_dirs; p.W = W; p.H = H; p.D = D; p.subpixel = subpixel; p.vz_to_disp = vz_to_disp;
    p.O = O; p.vMax = vMax; p.Sp16 = Sp16; p.bestD = bestD; p.minC = minC;
    const size_t N = (size_t)W * H;
    dim3 grid((unsigned)((N + WTA_WARPS * 32 - 1) / (WTA_WARPS * 32)), n);
    if (D % 8 == 0) epi_wta_kernel<true><<<grid, WTA_WARPS * 32, 0, c->stream>>>(p);
    else epi_wta_kernel<false><<<grid, WTA_WARPS * 32, 0, c->stream>>>(p);
    FSGM_LAUNCHED(c);
    return FSGM_OK;
}

// WTA over one slab of pixels from n_vols u8 partial volumes (the all-to-all form of the direction-split path)
int launch_slab_wta(fsgm_ctx* c, const uint8_t* vols, int n_vols, size_t vol_stride, const uint16_t* next0, size_t npix, int D, int subpixel,
                    int vz_to_disp, const double* O, double vMax, uint32_t* bestD, uint32_t* minC)
{
    if (D > WTA_MAXD) return fail(c, FSGM_ERR_DOMAIN, "label count must be <= 512");
    StageScope ss(c, ST_WTA);
    WtaParams p{};
    for (int k = 0; k < n_vols; ++k) p.L[k] = vols + (size_t)k * vol_stride;      // slabs may be padded past npix pixels
    p.n_dirs = n_vols; p.W = (int)npix; p.H = 1; p.D = D; p.subpixel = subpixel; p.vz_to_disp = vz_to_disp;
    p.O = O; p.vMax = vMax; p.Sp16 = nullptr; p.Sp_in = nullptr; p.next0 = next0; p.bestD = bestD; p.minC = minC;
    dim3 grid((unsigned)((npix + WTA_WARPS * 32 - 1) / (WTA_WARPS * 32)), 1);
    if (D % 8 == 0) epi_wta_kernel<true><<<grid, WTA_WARPS * 32, 0, c->stream>>>(p);
    else epi_wta_kernel<false><<<grid, WTA_WARPS * 32, 0, c->stream>>>(p);
    FSGM_LAUNCHED(c);
    return FSGM_OK;
}

__global__ void add_u8_kernel(uint4* __restrict__ a, const uint4* __restrict__ b, size_t n16)
{
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n16) return;
    uint4 x = a[i]; const uint4 y = b[i];
    x.x += y.x; x.y += y.y; x.z += y.z; x.w += y.w;          // byte sums stay below 256 (caller checks): no carries
    a[i] = x;
}

int launch_add_u8(fsgm_ctx* c, uint8_t* a, const uint8_t* b, size_t bytes)
{
    StageScope ss(c, ST_MISC);
    const size_t n16 = bytes / 16;
    add_u8_kernel<<<(unsigned)((n16 + 255) / 256), 256, 0, c->stream>>>(reinterpret_cast<uint4*>(a), reinterpret_cast<const uint4*>(b), n16);
    FSGM_LAUNCHED(c);
    return FSGM_OK;
}

// Largest byte of a caller-supplied u8 volume (stage entry points: decides between the exact-u16 / biased-fp16 kernels and the
// explicit mod-256 form).  Grid-stride 128-bit loads, byte-wise maxima, one atomicMax per warp; the head and tail that are
// not 16-byte aligned are read byte-wise by the first block.
__global__ void max_u8_kernel(const uint8_t* __restrict__ v, size_t bytes, unsigned* __restrict__ out)
{
    const size_t head = min(bytes, (size_t)((16 - (reinterpret_cast<uintptr_t>(v) & 15)) & 15));
    const uint4* q = reinterpret_cast<const uint4*>(v + head);
    const size_t n16 = (bytes - head) / 16;
    unsigned m = 0;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n16; i += (size_t)gridDim.x * blockDim.x) {
        const uint4 x = q[i];
        m = __vmaxu4(m, __vmaxu4(__vmaxu4(x.x, x.y), __vmaxu4(x.z, x.w)));
    }
    if (blockIdx.x == 0) {
        for (size_t i = threadIdx.x; i < head; i += blockDim.x) m = __vmaxu4(m, v[i]);
        for (size_t i = head + n16 * 16 + threadIdx.x; i < bytes; i += blockDim.x) m = __vmaxu4(m, v[i]);
    }
    m = max(max(m & 255u, (m >> 8) & 255u), max((m >> 16) & 255u, m >> 24));
    m = __reduce_max_sync(0xffffffffu, m);
    if ((threadIdx.x & 31) == 0 && m) atomicMax(out, m);
}

// synchronous (the host needs the value to pick a kernel): cmax = max byte of v[0..bytes)
int launch_max_u8(fsgm_ctx* c, const uint8_t* v, size_t bytes, int* cmax)
{
    if (!c->d_scalar) FSGM_CUDA(c, cudaMalloc(&c->d_scalar, 256));
    StageScope ss(c, ST_MISC);
    FSGM_CUDA(c, cudaMemsetAsync(c->d_scalar, 0, 4, c->stream));
    const unsigned blocks = (unsigned)std::min<size_t>((bytes / 16 + 255) / 256 + 1, (size_t)c->sm_count * 8);
    max_u8_kernel<<<blocks, 256, 0, c->stream>>>(v, bytes, static_cast<unsigned*>(c->d_scalar));
    FSGM_LAUNCHED(c);
    unsigned h = 0;
    FSGM_CUDA(c, cudaMemcpyAsync(&h, c->d_scalar, 4, cudaMemcpyDeviceToHost, c->stream));
    FSGM_CUDA(c, cudaStreamSynchronize(c->stream));
    *cmax = (int)h;
    return FSGM_OK;
}

// WTA from an already-summed u16 volume covering `npix` consecutive pixels (one slab of the direction-split path)
int launch_sp_wta(fsgm_ctx* c, const uint16_t* Sp, const uint16_t* next0, size_t npix, int D, int subpixel,
                  int vz_to_disp, const double* O, double vMax, uint32_t* bestD, uint32_t* minC)
{
    if (D > WTA_MAXD) return fail(c, FSGM_ERR_DOMAIN, "label count must be <= 512");
    StageScope ss(c, ST_WTA);
    WtaParams p{};
    p.n_dirs = 0; p.W = (int)npix; p.H = 1; p.D = D; p.subpixel = subpixel; p.vz_to_disp = vz_to_disp;
    p.O = O; p.vMax = vMax; p.Sp16 = nullptr; p.Sp_in = Sp; p.next0 = next0; p.bestD = bestD; p.minC = minC;
    dim3 grid((unsigned)((npix + WTA_WARPS * 32 - 1) / (WTA_WARPS * 32)), 1);
    epi_wta_kernel<false><<<grid, WTA_WARPS * 32, 0, c->stream>>>(p);
    FSGM_LAUNCHED(c);
    return FSGM_OK;
}

}  // namespace fsgm
