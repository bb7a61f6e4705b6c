// C-ABI layer (include/fsgm.h): context, scratch arena, the gateway-shaped entry points and the
// stage entry points.  Host code is C++; everything that touches pixels is a kernel in the sibling
// .cu files.  There is no CPU fallback anywhere in this library.
#include "fsgm_internal.h"
#include <cstdio>
#include <cstring>
#include <new>
#include <algorithm>

namespace fsgm {

int fail(fsgm_ctx* c, int code, const char* what, const char* detail)
{
    if (c) {
        c->err = what ? what : "";
        if (detail) { c->err += ": "; c->err += detail; }
    }
    return code;
}

int arena_reserve(fsgm_ctx* c, size_t total)
{
    total = align256(total) + 4096;
    if (c->arena_top != 0) return fail(c, FSGM_ERR_ARG, "arena_reserve inside an active scope");
    if (total <= c->arena_bytes) return FSGM_OK;
    FSGM_CUDA(c, cudaStreamSynchronize(c->stream));
    if (c->arena) { cudaFree(c->arena); c->arena = nullptr; c->arena_bytes = 0; }
    cudaError_t e = cudaMalloc(reinterpret_cast<void**>(&c->arena), total);
    if (e != cudaSuccess) { cudaGetLastError(); return fail(c, FSGM_ERR_NOMEM, "cudaMalloc(scratch arena)", cudaGetErrorString(e)); }
    c->arena_bytes = total;
    return FSGM_OK;
}

int arena_alloc(fsgm_ctx* c, size_t bytes, void** out)
{
    size_t need = align256(bytes);
    if (c->arena_top + need > c->arena_bytes) return fail(c, FSGM_ERR_NOMEM, "scratch arena exhausted (reserve too small)");
    *out = c->arena + c->arena_top;
    c->arena_top += need;
    return FSGM_OK;
}

StageScope::StageScope(fsgm_ctx* ctx, int st) : c(ctx), stage(st), launches0(ctx->launches)
{
    if (!c->profiling) return;
    if (c->event_pool.empty()) { cudaEvent_t e; if (cudaEventCreate(&e) != cudaSuccess) return; c->event_pool.push_back(e); }
    a = c->event_pool.back(); c->event_pool.pop_back();
    cudaEventRecord(a, c->stream);
}
StageScope::~StageScope()
{
    c->stage_launches[stage] += c->launches - launches0;
    if (!a) return;
    cudaEvent_t b;
    if (c->event_pool.empty()) { if (cudaEventCreate(&b) != cudaSuccess) { c->event_pool.push_back(a); return; } }
    else { b = c->event_pool.back(); c->event_pool.pop_back(); }
    cudaEventRecord(b, c->stream);
    c->timers.push_back({a, b, stage});
}

static void profile_collect(fsgm_ctx* c)
{
    for (auto& t : c->timers) {
        cudaEventSynchronize(t.b);
        float ms = 0.f;
        if (cudaEventElapsedTime(&ms, t.a, t.b) == cudaSuccess) c->stage_ms[t.stage] += ms;
        c->event_pool.push_back(t.a); c->event_pool.push_back(t.b);
    }
    c->timers.clear();
}

int pipe_reserve(fsgm_ctx* c, size_t bytes_per_slot)
{
    HostPipe& p = c->pipe;
    if (!p.h2d) {
        FSGM_CUDA(c, cudaStreamCreateWithFlags(&p.h2d, cudaStreamNonBlocking));
        FSGM_CUDA(c, cudaStreamCreateWithFlags(&p.d2h, cudaStreamNonBlocking));
        for (int i = 0; i < 2; ++i) {
            FSGM_CUDA(c, cudaEventCreateWithFlags(&p.in_ready[i], cudaEventDisableTiming));
            FSGM_CUDA(c, cudaEventCreateWithFlags(&p.done[i], cudaEventDisableTiming));
            FSGM_CUDA(c, cudaEventCreateWithFlags(&p.out_ready[i], cudaEventDisableTiming));
        }
    }
    bytes_per_slot = align256(bytes_per_slot);
    if (bytes_per_slot <= p.bytes) return FSGM_OK;
    FSGM_CUDA(c, cudaDeviceSynchronize());
    for (int i = 0; i < 2; ++i) { if (p.buf[i]) cudaFree(p.buf[i]); p.buf[i] = nullptr; }
    p.bytes = 0; p.used[0] = p.used[1] = 0;
    for (int i = 0; i < 2; ++i)
        if (cudaMalloc(reinterpret_cast<void**>(&p.buf[i]), bytes_per_slot) != cudaSuccess) {
            cudaGetLastError();
            return fail(c, FSGM_ERR_NOMEM, "cudaMalloc(host-gateway staging)");
        }
    p.bytes = bytes_per_slot;
    return FSGM_OK;
}

int check_dims(fsgm_ctx* c, int n, int W, int H, int D)
{
    if (!c) return FSGM_ERR_ARG;
    if (n < 1 || W < 1 || H < 1) return fail(c, FSGM_ERR_ARG, "n_pairs, width and height must be positive");
    if (D < 1 || D > 512) return fail(c, FSGM_ERR_DOMAIN, "dMax must be in 1..512");
    if ((size_t)W * H > (size_t)1 << 30) return fail(c, FSGM_ERR_DOMAIN, "image too large");
    return FSGM_OK;
}

int enabled_dirs(const fsgm_epi_opts& o, int* dirs)
{
    int k = 0;
    for (int r = 0; r < 8; ++r)
        if (dir_enabled(r, o.total_pass, o.paths == 8)) dirs[k++] = r;
    return k;
}

int check_opts(fsgm_ctx* c, const fsgm_epi_opts* in, fsgm_epi_opts* o)
{
    if (in) *o = *in; else fsgm_epi_opts_default(o);
    if (o->paths != 4 && o->paths != 8) return fail(c, FSGM_ERR_ARG, "opts.paths must be 4 or 8");
    if (o->total_pass < 1 || o->total_pass > 2) return fail(c, FSGM_ERR_DOMAIN, "opts.total_pass must be 1 or 2");
    return FSGM_OK;
}

constexpr int VS_MAX_SMEM = 227 * 1024;

// largest byte of a caller-supplied cost volume, never reported below the 5x5-census bound the kernels are tuned for
static int caller_volume_cmax(fsgm_ctx* c, const uint8_t* d_C, size_t bytes, int* cmax)
{
    FSGM_TRY(launch_max_u8(c, d_C, bytes, cmax));
    *cmax = std::max(*cmax, 24);
    return FSGM_OK;
}

// cluster size of the row-synchronous fast path for this problem, or 0 if it does not apply
static int fast_path_cluster(fsgm_ctx* c, int W, int D, int P1, int P2, int cmax, const fsgm_epi_opts& o)
{
    if (c->force_cluster < 0) return 0;
    if (o.adaptive_p2 || sweep_needs_wrap(P1, P2, cmax)) return 0;
    if (o.total_pass < 1 || o.total_pass > 2) return 0;
    const int ndir = o.paths == 8 ? 3 : 1;
    if (c->force_cluster > 0)                  // A/B knob: the forced size must fit and leave every CTA >= 2 columns
        return vsweep_cluster_ok(W, D, ndir, c->force_cluster, VS_MAX_SMEM) ? c->force_cluster : vsweep_cluster_size(W, D, ndir, VS_MAX_SMEM);
    if (c->best_key[0] != W || c->best_key[1] != D || c->best_key[2] != ndir) {      // occupancy queries: once per problem shape
        c->best_cs = vsweep_best_cluster(W, D, ndir, VS_MAX_SMEM, &c->best_clusters);
        c->best_key[0] = W; c->best_key[1] = D; c->best_key[2] = ndir;
    }
    return c->best_cs;
}

// How many of n pairs go through the cluster kernels.  Each pair occupies one cluster for a whole pass, so the work
// comes in waves of K = resident clusters (15 clusters of 8 CTAs on a 148-SM B200): full waves take the fast path; a
// partial wave costs as much as a full one (measured at KITTI size: ~6.6 ms per wave against ~0.65 ms per pair through
// the generic kernels, which fill the whole GPU with independent scanlines), so it is only used when (almost) full.
static int fast_pairs(fsgm_ctx* c, int n, int cs, int D, int W, int ndir)
{
    if (!cs) return 0;
    if (c->clusters_key[0] != cs || c->clusters_key[1] != W || c->clusters_key[2] != D || c->clusters_key[3] != ndir) {
        const int Wk = (W + cs - 1) / cs;                // resident clusters depend on the smem footprint: key on the whole shape
        int k = vsweep_max_clusters(cs, vsweep_smem_bytes(D, Wk, ndir), vsweep_threads());
        c->clusters_key[0] = cs; c->clusters_key[1] = W; c->clusters_key[2] = D; c->clusters_key[3] = ndir;
        c->clusters_max = k > 0 ? k : 1;
    }
    if (c->force_cluster > 0) return n;                 // explicit A/B request: everything through the cluster kernels
    const int K = c->clusters_max, r = n % K;
    return n - r + (r + 1 >= K ? r : 0);
}

static size_t aggregate_scratch_bytes(fsgm_ctx* c, int n, int W, int H, int D, int P1, int P2, int cmax, const fsgm_epi_opts& o)
{
    const size_t N = (size_t)W * H, V = N * D;
    const int cs = fast_path_cluster(c, W, D, P1, P2, cmax, o);
    const int nf = fast_pairs(c, n, cs, D, W, o.paths == 8 ? 3 : 1), ng = n - nf;
    int dirs[8];
    size_t b = 1024;
    if (nf) b += 2 * align256(nf * V) + align256(nf * V * 2) + align256(nf * N * 8);
    if (ng) b += (size_t)enabled_dirs(o, dirs) * align256(ng * V);
    return b;
}

// sweeps + WTA + subpixel (+ vz): the row-synchronous cluster kernels when they apply, else generic sweeps + WTA kernel.
// Sp16 (optional) receives the summed volume for stage parity.
static int aggregate_and_wta(fsgm_ctx* c, int n, const uint8_t* C, const uint8_t* I1, int W, int H, int D, int P1, int P2, int cmax,
                             const fsgm_epi_opts& o, const double* O, double vMax, uint16_t* Sp16, uint32_t* bestD, uint32_t* minC)
{
    const size_t N = (size_t)W * H, V = N * D;
    const int cs = fast_path_cluster(c, W, D, P1, P2, cmax, o);
    const int ndir = o.paths == 8 ? 3 : 1;
    const int nf = fast_pairs(c, n, cs, D, W, ndir), ng = n - nf;
    if (nf) {
        uint8_t *Lh0, *Lh1 = nullptr; uint16_t *S1, *rec;
        FSGM_TRY(arena_get(c, nf * V, &Lh0));
        FSGM_TRY(arena_get(c, nf * V, &Lh1));
        FSGM_TRY(arena_get(c, nf * V, &S1));
        FSGM_TRY(arena_get(c, nf * N * 4, &rec));
        const int hd[2] = {0, 4};
        uint8_t* Lh[2] = {Lh0, Lh1};
        FSGM_TRY(launch_sweeps(c, nf, C, I1, W, H, D, P1, P2, 0, cmax, hd, o.total_pass == 2 ? 2 : 1, Lh));
        // FAST operand configuration of the two cluster passes (vsweep.cu): byte volume between them, biased WTA records
        const bool fast = o.total_pass == 2 && !Sp16 && vsweep_fast_ok(ndir, P2);
        if (fast) {
            FSGM_TRY(launch_vsweep(c, nf, cs, ndir, false, C, nullptr, nullptr, nullptr, S1, nullptr, nullptr, W, H, D, P1, P2, 0, true));
            FSGM_TRY(launch_vsweep(c, nf, cs, ndir, true, C, Lh0, Lh1, S1, nullptr, minC, rec, W, H, D, P1, P2, 1, true));
        } else if (o.total_pass == 2) {
            FSGM_TRY(launch_vsweep(c, nf, cs, ndir, false, C, Lh0, Lh1, nullptr, S1, nullptr, nullptr, W, H, D, P1, P2, 0, false));
            FSGM_TRY(launch_vsweep(c, nf, cs, ndir, true, C, nullptr, nullptr, S1, Sp16, minC, rec, W, H, D, P1, P2, 1, false));
        } else {
            FSGM_TRY(launch_vsweep(c, nf, cs, ndir, true, C, Lh0, nullptr, nullptr, Sp16, minC, rec, W, H, D, P1, P2, 0, false));
        }
        FSGM_TRY(launch_vs_finalize(c, nf, rec, minC, O, W, H, D, o.subpixel, o.vz_to_disp, vMax, fast ? 1 : 0, bestD));
    }
    if (ng) {
        int dirs[8];
        const int nd = enabled_dirs(o, dirs);
        uint8_t* L[8];
        for (int k = 0; k < nd; ++k) FSGM_TRY(arena_get(c, ng * V, &L[k]));
        const size_t po = (size_t)nf * N;                       // pair offset of the generic part
        FSGM_TRY(launch_sweeps(c, ng, C + po * D, I1 ? I1 + po : nullptr, W, H, D, P1, P2, o.adaptive_p2 ? 25 : 0, cmax, dirs, nd, L));
        FSGM_TRY(launch_epi_wta(c, ng, L, nd, W, H, D, o.subpixel, o.vz_to_disp, O ? O + po : nullptr, vMax,
                                Sp16 ? Sp16 + po * D : nullptr, bestD + po, minC + po));
    }
    return FSGM_OK;
}

// scratch needed by the epipolar pipeline for `n` pairs
static size_t epi_scratch_bytes(fsgm_ctx* c, int n, int W, int H, int D, int P1, int P2, const fsgm_epi_opts& o)
{
    const size_t N = (size_t)W * H, V = N * D;
    const bool fused = (D == 64 || D == 128 || D == 256);
    return 2 * align256(n * N * 4) + align256(D * 8) + (fused ? 1 : 2) * align256(n * V) +
           aggregate_scratch_bytes(c, n, W, H, D, P1, P2, 24, o) + 4096;
}

// The whole hot path on device-resident inputs (asynchronous on c->stream).
static int epi_pipeline_dev(fsgm_ctx* c, int n, const uint8_t* I1, const uint8_t* I2, int W, int H, int D, double vMax,
                            const double* Pd0, const double* dirn, const double* O, int P1, int P2,
                            const fsgm_epi_opts& o, uint32_t* bestD, uint32_t* minC)
{
    const size_t N = (size_t)W * H, V = N * D;
    uint32_t *cen1, *cen2; uint8_t *raw = nullptr, *C; double* vz;
    FSGM_TRY(arena_get(c, (size_t)D, &vz));
    FSGM_TRY(arena_get(c, n * N, &cen1));
    FSGM_TRY(arena_get(c, n * N, &cen2));
    FSGM_TRY(arena_get(c, n * V, &C));
    FSGM_TRY(launch_census(c, n, I1, W, H, cen1));
    FSGM_TRY(launch_census(c, n, I2, W, H, cen2));
    FSGM_TRY(launch_vz_table(c, D, vMax, vz));
    bool fused = false;
    FSGM_TRY(launch_epi_cost_fused(c, n, vMax, cen1, cen2, W, H, D, Pd0, dirn, O, C, &fused));
    if (!fused) {
        FSGM_TRY(arena_get(c, n * V, &raw));
        FSGM_TRY(launch_epi_cost(c, n, vz, cen1, cen2, W, H, D, vMax, Pd0, dirn, O, raw, C));
    }
    FSGM_TRY(aggregate_and_wta(c, n, C, I1, W, H, D, P1, P2, /*cmax=*/24, o, O, vMax, nullptr, bestD, minC));
    return FSGM_OK;
}


// ---- wave-pipelined epipolar path -----------------------------------------------------------------------------
// The cluster kernels keep 15 x 8 = 120 of the 148 SMs busy and a pair occupies a cluster for a whole pass, so a batch
// runs in waves of K pairs.  While wave i is in its two cluster passes (stream A), the front-end of wave i+1 (census,
// cost volume, horizontal sweeps) runs on a second stream and fills the SMs the clusters leave idle.
struct StreamSwap {
    fsgm_ctx* c; cudaStream_t saved;
    StreamSwap(fsgm_ctx* ctx, cudaStream_t s) : c(ctx), saved(ctx->stream) { c->stream = s; }
    ~StreamSwap() { c->stream = saved; }
};

static int epi_pipeline_waves(fsgm_ctx* c, int n, int cs, const uint8_t* I1, const uint8_t* I2, int W, int H, int D, double vMax,
                              const double* Pd0, const double* dirn, const double* O, int P1, int P2,
                              const fsgm_epi_opts& o, uint32_t* bestD, uint32_t* minC)
{
    const size_t N = (size_t)W * H, V = N * D;
    const int ndir = o.paths == 8 ? 3 : 1;
    const int nf = fast_pairs(c, n, cs, D, W, ndir), ng = n - nf, K = c->clusters_max;
    if (!c->aux_stream) {
        FSGM_CUDA(c, cudaStreamCreateWithFlags(&c->aux_stream, cudaStreamNonBlocking));
        FSGM_CUDA(c, cudaEventCreateWithFlags(&c->ev_entry, cudaEventDisableTiming));
        FSGM_CUDA(c, cudaEventCreateWithFlags(&c->ev_front[0], cudaEventDisableTiming));
        FSGM_CUDA(c, cudaEventCreateWithFlags(&c->ev_front[1], cudaEventDisableTiming));
    }
    uint32_t *cen1, *cen2; uint8_t *C, *Lh0, *Lh1; uint16_t *S1, *rec;
    FSGM_TRY(arena_get(c, n * N, &cen1));
    FSGM_TRY(arena_get(c, n * N, &cen2));
    FSGM_TRY(arena_get(c, n * V, &C));
    FSGM_TRY(arena_get(c, nf * V, &Lh0));
    FSGM_TRY(arena_get(c, nf * V, &Lh1));
    FSGM_TRY(arena_get(c, nf * V, &S1));
    FSGM_TRY(arena_get(c, nf * N * 4, &rec));
    cudaStream_t A = c->stream, B = c->aux_stream;
    FSGM_CUDA(c, cudaEventRecord(c->ev_entry, A));            // everything queued so far (the caller's inputs) precedes stream B
    FSGM_CUDA(c, cudaStreamWaitEvent(B, c->ev_entry, 0));
    const int hd[2] = {0, 4};
    const bool fast = vsweep_fast_ok(ndir, P2);         // FAST operand configuration of the two cluster passes (vsweep.cu)
    auto front = [&](int p0, int m, cudaStream_t st) -> int {
        StreamSwap sw(c, st);
        FSGM_TRY(launch_census(c, m, I1 + p0 * N, W, H, cen1 + p0 * N));
        FSGM_TRY(launch_census(c, m, I2 + p0 * N, W, H, cen2 + p0 * N));
        bool fused = false;
        FSGM_TRY(launch_epi_cost_fused(c, m, vMax, cen1 + p0 * N, cen2 + p0 * N, W, H, D, Pd0 + p0 * 2 * N, dirn + p0 * 2 * N,
                                       O + p0 * N, C + p0 * V, &fused));
        if (!fused) return fail(c, FSGM_ERR_DOMAIN, "wave pipeline needs the fused cost kernel");
        uint8_t* Lh[2] = {Lh0 + p0 * V, Lh1 + p0 * V};
        return launch_sweeps(c, m, C + p0 * V, I1 + p0 * N, W, H, D, P1, P2, 0, 24, hd, 2, Lh);
    };
    auto back = [&](int p0, int m) -> int {
        if (fast) {
            FSGM_TRY(launch_vsweep(c, m, cs, ndir, false, C + p0 * V, nullptr, nullptr, nullptr, S1 + p0 * V, nullptr, nullptr,
                                   W, H, D, P1, P2, 0, true));
            FSGM_TRY(launch_vsweep(c, m, cs, ndir, true, C + p0 * V, Lh0 + p0 * V, Lh1 + p0 * V, S1 + p0 * V, nullptr, minC + p0 * N,
                                   rec + p0 * N * 4, W, H, D, P1, P2, 1, true));
        } else {
            FSGM_TRY(launch_vsweep(c, m, cs, ndir, false, C + p0 * V, Lh0 + p0 * V, Lh1 + p0 * V, nullptr, S1 + p0 * V, nullptr, nullptr,
                                   W, H, D, P1, P2, 0, false));
            FSGM_TRY(launch_vsweep(c, m, cs, ndir, true, C + p0 * V, nullptr, nullptr, S1 + p0 * V, nullptr, minC + p0 * N,
                                   rec + p0 * N * 4, W, H, D, P1, P2, 1, false));
        }
        return launch_vs_finalize(c, m, rec + p0 * N * 4, minC + p0 * N, O + p0 * N, W, H, D, o.subpixel, o.vz_to_disp, vMax, fast ? 1 : 0,
                                  bestD + p0 * N);
    };
    // partial wave (fewer pairs than resident clusters are left): generic one-warp-per-scanline kernels.  It is queued on stream B under
    // the cluster passes of the last full wave (a fixed batch of 32 pairs per GPU — config E on 8 GPUs — is two waves + 2 pairs).
    auto tail = [&](cudaStream_t st) -> int {
        StreamSwap sw(c, st);
        const size_t po = (size_t)nf * N;
        FSGM_TRY(launch_census(c, ng, I1 + po, W, H, cen1 + po));
        FSGM_TRY(launch_census(c, ng, I2 + po, W, H, cen2 + po));
        bool fused = false;
        FSGM_TRY(launch_epi_cost_fused(c, ng, vMax, cen1 + po, cen2 + po, W, H, D, Pd0 + 2 * po, dirn + 2 * po, O + po, C + po * D, &fused));
        fsgm_epi_opts og = o;
        const int saved = c->force_cluster;
        c->force_cluster = -1;                                // generic path for these pairs
        int rc = aggregate_and_wta(c, ng, C + po * D, I1 + po, W, H, D, P1, P2, 24, og, O + po, vMax, nullptr, bestD + po, minC + po);
        c->force_cluster = saved;
        return rc;
    };
    const int waves = (nf + K - 1) / K;
    if (waves > 0) FSGM_TRY(front(0, std::min(K, nf), A));
    for (int i = 0; i < waves; ++i) {
        const int p0 = i * K, m = std::min(K, nf - p0);
        if (i + 1 < waves) {
            const int q0 = (i + 1) * K, qm = std::min(K, nf - q0);
            FSGM_TRY(front(q0, qm, B));
            FSGM_CUDA(c, cudaEventRecord(c->ev_front[(i + 1) & 1], B));
        }
        if (i > 0) FSGM_CUDA(c, cudaStreamWaitEvent(A, c->ev_front[i & 1], 0));
        if (i + 1 == waves && ng) FSGM_TRY(tail(B));          // stream B has nothing left to do for the last wave: the partial wave goes there
        FSGM_TRY(back(p0, m));
    }
    if (ng) {
        if (waves == 0) FSGM_TRY(tail(A));
        else {
            FSGM_CUDA(c, cudaEventRecord(c->ev_front[waves & 1], B));
            FSGM_CUDA(c, cudaStreamWaitEvent(A, c->ev_front[waves & 1], 0));
        }
    }
    return FSGM_OK;
}

}  // namespace fsgm

using namespace fsgm;

extern "C" {

int fsgm_abi_version(void) { return 3; }

void fsgm_epi_opts_default(fsgm_epi_opts* o)
{
    if (!o) return;
    o->paths = 4; o->total_pass = 2; o->subpixel = 1; o->adaptive_p2 = 0; o->vz_to_disp = 1; o->fb_check = 0; o->fb_thr = 2;
}

int fsgm_create(int device, fsgm_ctx** out)
{
    if (!out) return FSGM_ERR_ARG;
    *out = nullptr;
    int count = 0;
    if (cudaGetDeviceCount(&count) != cudaSuccess || device < 0 || device >= count) { cudaGetLastError(); return FSGM_ERR_CUDA; }
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, device) != cudaSuccess) return FSGM_ERR_CUDA;
    if (prop.major != 10) return FSGM_ERR_CUDA;          // kernels are built for sm_100a only; no fallback
    if (cudaSetDevice(device) != cudaSuccess) return FSGM_ERR_CUDA;
    fsgm_ctx* c = new (std::nothrow) fsgm_ctx();
    if (!c) return FSGM_ERR_NOMEM;
    c->device = device;
    c->sm_count = prop.multiProcessorCount;
    if (cudaStreamCreateWithFlags(&c->own_stream, cudaStreamNonBlocking) != cudaSuccess) { delete c; return FSGM_ERR_CUDA; }
    c->stream = c->own_stream;
    *out = c;
    return FSGM_OK;
}

void fsgm_destroy(fsgm_ctx* c)
{
    if (!c) return;
    cudaSetDevice(c->device);
    cudaStreamSynchronize(c->stream);
    cudaDeviceSynchronize();
    if (c->arena) cudaFree(c->arena);
    if (c->geo_params) cudaFree(c->geo_params);
    if (c->d_scalar) cudaFree(c->d_scalar);
    dist_release(c);
    if (c->geo_host) cudaFreeHost(c->geo_host);
    for (auto e : c->geo_ev) if (e) cudaEventDestroy(e);
    for (auto& t : c->timers) { cudaEventDestroy(t.a); cudaEventDestroy(t.b); }
    for (auto e : c->event_pool) cudaEventDestroy(e);
    for (int i = 0; i < 2; ++i) {
        if (c->pipe.buf[i]) cudaFree(c->pipe.buf[i]);
        if (c->pipe.in_ready[i]) { cudaEventDestroy(c->pipe.in_ready[i]); cudaEventDestroy(c->pipe.done[i]); cudaEventDestroy(c->pipe.out_ready[i]); }
    }
    if (c->pipe.h2d) { cudaStreamDestroy(c->pipe.h2d); cudaStreamDestroy(c->pipe.d2h); }
    if (c->aux_stream) { cudaStreamDestroy(c->aux_stream); cudaEventDestroy(c->ev_entry); cudaEventDestroy(c->ev_front[0]); cudaEventDestroy(c->ev_front[1]); }
    if (c->own_stream) cudaStreamDestroy(c->own_stream);
    delete c;
}

int fsgm_set_stream(fsgm_ctx* c, void* s)
{
    if (!c) return FSGM_ERR_ARG;
    FSGM_CUDA(c, cudaStreamSynchronize(c->stream));
    c->stream = s ? static_cast<cudaStream_t>(s) : c->own_stream;
    return FSGM_OK;
}

int fsgm_synchronize(fsgm_ctx* c)
{
    if (!c) return FSGM_ERR_ARG;
    if (c->pipe.h2d) { FSGM_CUDA(c, cudaStreamSynchronize(c->pipe.h2d)); }
    FSGM_CUDA(c, cudaStreamSynchronize(c->stream));
    if (c->pipe.d2h) { FSGM_CUDA(c, cudaStreamSynchronize(c->pipe.d2h)); }
    return FSGM_OK;
}

int fsgm_epi_wave_pairs(fsgm_ctx* c, int W, int D, int P1, int P2, const fsgm_epi_opts* opts)
{
    if (!c || W < 1 || D < 1) return 0;
    fsgm_epi_opts o;
    if (check_opts(c, opts, &o) != FSGM_OK) return 0;
    if (cudaSetDevice(c->device) != cudaSuccess) return 0;
    const int cs = fast_path_cluster(c, W, D, P1, P2, 24, o);
    if (!cs) return 0;
    fast_pairs(c, 1, cs, D, W, o.paths == 8 ? 3 : 1);
    return c->clusters_max;
}

int fsgm_debug_max_clusters(int cs, size_t smem, int threads) { return fsgm::vsweep_max_clusters(cs, smem, threads); }

int fsgm_tune(fsgm_ctx* c, int key, int value)
{
    if (!c) return FSGM_ERR_ARG;
    if (key == 1) { c->force_cluster = value; return FSGM_OK; }
    if (key == 2) { c->no_overlap = value != 0; return FSGM_OK; }
    if (key == 4) { c->no_p2p = value != 0; return FSGM_OK; }
    if (key == 5) { c->pyd_cluster = value; return FSGM_OK; }
    if (key == 6) { c->pyd_direct_cost = value != 0; return FSGM_OK; }
    if (key == 7) { c->pydng_generic = value != 0; return FSGM_OK; }
    if (key == 9) { c->stage_waves = value < 1 ? 3 : value > 64 ? 64 : value; return FSGM_OK; }
    if (key == 8) { c->fc_rows = value < 0 ? 0 : value > 4096 ? 4096 : value; return FSGM_OK; }
    if (key == 3) { c->ng_occupancy = value < 0 ? 0 : value > 3 ? 3 : value; return FSGM_OK; }
    return fail(c, FSGM_ERR_ARG, "unknown tuning key");
}

const char* fsgm_last_error(const fsgm_ctx* c) { return c ? c->err.c_str() : "null context"; }
uint64_t fsgm_launch_count(const fsgm_ctx* c) { return c ? c->launches : 0; }
size_t fsgm_scratch_bytes(const fsgm_ctx* c) { return c ? c->arena_bytes : 0; }

int fsgm_profile_enable(fsgm_ctx* c, int on)
{
    if (!c) return FSGM_ERR_ARG;
    profile_collect(c);
    c->profiling = on != 0;
    return FSGM_OK;
}
int fsgm_profile_reset(fsgm_ctx* c)
{
    if (!c) return FSGM_ERR_ARG;
    profile_collect(c);
    for (int i = 0; i < ST_COUNT; ++i) { c->stage_ms[i] = 0; c->stage_launches[i] = 0; }
    return FSGM_OK;
}
int fsgm_profile_read(fsgm_ctx* c, int stage, double* ms, uint64_t* launches)
{
    if (!c || stage < 0 || stage >= ST_COUNT) return FSGM_ERR_ARG;
    profile_collect(c);
    if (ms) *ms = c->stage_ms[stage];
    if (launches) *launches = c->stage_launches[stage];
    return FSGM_OK;
}
const char* fsgm_stage_name(int stage)
{
    static const char* names[ST_COUNT] = { "census", "epi_cost", "sweep", "wta", "pyd_cost", "pyd_sweep", "pyd_wta",
                                           "ng", "pydng_cost", "pydng_sweep", "pydng_wta", "misc", "vsweep", "pyramid", "geometry", "exchange" };
    return (stage >= 0 && stage < ST_COUNT) ? names[stage] : nullptr;
}
int fsgm_stage_count(void) { return ST_COUNT; }

// ---------------------------------------------------------------------------------------------
// stage entry points
// ---------------------------------------------------------------------------------------------
int fsgm_census_dev(fsgm_ctx* c, int n_images, const uint8_t* d_img, int W, int H, uint32_t* d_census)
{
    FSGM_TRY(check_dims(c, n_images, W, H, 1));
    if (!d_img || !d_census) return fail(c, FSGM_ERR_ARG, "null pointer");
    FSGM_CUDA(c, cudaSetDevice(c->device));
    return launch_census(c, n_images, d_img, W, H, d_census);
}

int fsgm_epi_cost_dev(fsgm_ctx* c, int n, const uint32_t* d_cen1, const uint32_t* d_cen2, int W, int H, int D, double vMax,
                      const double* d_Pd0, const double* d_dir, const double* d_O, uint8_t* d_raw, uint8_t* d_C)
{
    FSGM_TRY(check_dims(c, n, W, H, D));
    if (!d_cen1 || !d_cen2 || !d_Pd0 || !d_dir || !d_O || !d_C) return fail(c, FSGM_ERR_ARG, "null pointer");
    FSGM_CUDA(c, cudaSetDevice(c->device));
    const size_t V = (size_t)W * H * D;
    FSGM_TRY(arena_reserve(c, align256(D * 8) + (d_raw ? 0 : align256(n * V))));
    ArenaScope scope(c);
    uint8_t* raw = d_raw;
    double* vz;
    FSGM_TRY(arena_get(c, (size_t)D, &vz));
    if (!raw) FSGM_TRY(arena_get(c, n * V, &raw));
    FSGM_TRY(launch_vz_table(c, D, vMax, vz));
    if (!d_raw) {                                   // nobody wants the pre-box volume: fused kernel when it applies
        bool fused = false;
        FSGM_TRY(launch_epi_cost_fused(c, n, vMax, d_cen1, d_cen2, W, H, D, d_Pd0, d_dir, d_O, d_C, &fused));
        if (fused) return FSGM_OK;
    }
    return launch_epi_cost(c, n, vz, d_cen1, d_cen2, W, H, D, vMax, d_Pd0, d_dir, d_O, raw, d_C);
}

int fsgm_sweep_dev(fsgm_ctx* c, int n, const uint8_t* d_C, const uint8_t* d_I1, int W, int H, int D,
                   int P1, int P2, int adaptive_thr, int direction, uint8_t* d_L)
{
    FSGM_TRY(check_dims(c, n, W, H, D));
    if (!d_C || !d_L || (adaptive_thr > 0 && !d_I1)) return fail(c, FSGM_ERR_ARG, "null pointer");
    if (direction < 0 || direction > 7) return fail(c, FSGM_ERR_ARG, "direction must be 0..7");
    FSGM_CUDA(c, cudaSetDevice(c->device));
    uint8_t* L[1] = { d_L };
    // stage-level callers may pass any u8 volume: assume nothing about its range (cmax = 255)
    return launch_sweeps(c, n, d_C, d_I1, W, H, D, P1, P2, adaptive_thr, 255, &direction, 1, L);
}

int fsgm_epi_aggregate_dev(fsgm_ctx* c, int n, const uint8_t* d_C, const uint8_t* d_I1, int W, int H, int D,
                           int P1, int P2, const fsgm_epi_opts* opts, uint16_t* d_Sp,
                           const double* d_O, double vMax, uint32_t* d_bestD, uint32_t* d_minC)
{
    FSGM_TRY(check_dims(c, n, W, H, D));
    fsgm_epi_opts o;
    FSGM_TRY(check_opts(c, opts, &o));
    if (!d_C || !d_bestD || !d_minC || (o.vz_to_disp && !d_O) || (o.adaptive_p2 && !d_I1)) return fail(c, FSGM_ERR_ARG, "null pointer");
    FSGM_CUDA(c, cudaSetDevice(c->device));
    // The caller owns this volume: measure its largest byte instead of assuming the 5x5-census bound.  Up to 24 (anything
    // fsgm_epi_cost_dev produces) every kernel family applies; above it the row-synchronous cluster kernels (whose byte-sum
    // and biased-fp16 forms are derived for C <= 24) are skipped and the generic sweeps pick the exact or the explicit
    // mod-256 form from the measured bound, as the reference's unsigned char arithmetic requires.
    int cmax = 0;
    FSGM_TRY(caller_volume_cmax(c, d_C, (size_t)n * W * H * D, &cmax));
    const int saved = c->force_cluster;
    if (cmax > 24 || (reinterpret_cast<uintptr_t>(d_C) & 15)) c->force_cluster = -1;    // bulk copies need 16-byte alignment
    int rc = arena_reserve(c, aggregate_scratch_bytes(c, n, W, H, D, P1, P2, cmax, o));
    if (rc == FSGM_OK) {
        ArenaScope scope(c);
        rc = aggregate_and_wta(c, n, d_C, d_I1, W, H, D, P1, P2, cmax, o, d_O, vMax, d_Sp, d_bestD, d_minC);
    }
    c->force_cluster = saved;
    return rc;
}

// Direction-split building blocks (one large pair, the R directions spread over GPUs):
//  fsgm_epi_partial_dev : sweeps for the listed directions, summed into a u16 partial volume [N][D]
//  fsgm_epi_wta_sp_dev  : WTA/subpixel/vz over a slab of an already reduced u16 volume
int fsgm_epi_partial_dev(fsgm_ctx* c, const uint8_t* d_C, const uint8_t* d_I1, int W, int H, int D, int P1, int P2,
                         int adaptive_p2, const int* directions, int n_dirs, uint16_t* d_Sp_partial)
{
    FSGM_TRY(check_dims(c, 1, W, H, D));
    if (!d_C || !d_Sp_partial || !directions || (adaptive_p2 && !d_I1)) return fail(c, FSGM_ERR_ARG, "null pointer");
    if (n_dirs < 0 || n_dirs > 8) return fail(c, FSGM_ERR_ARG, "n_dirs must be 0..8");
    for (int k = 0; k < n_dirs; ++k)
        if (directions[k] < 0 || directions[k] > 7) return fail(c, FSGM_ERR_ARG, "direction must be 0..7");
    FSGM_CUDA(c, cudaSetDevice(c->device));
    const size_t N = (size_t)W * H, V = N * D;
    if (n_dirs == 0) { FSGM_CUDA(c, cudaMemsetAsync(d_Sp_partial, 0, V * 2, c->stream)); return FSGM_OK; }
    FSGM_TRY(arena_reserve(c, (size_t)n_dirs * align256(V) + 2 * align256(N * 4)));
    ArenaScope scope(c);
    uint8_t* L[8]; uint32_t *b, *m;
    for (int k = 0; k < n_dirs; ++k) FSGM_TRY(arena_get(c, V, &L[k]));
    FSGM_TRY(arena_get(c, N, &b));
    FSGM_TRY(arena_get(c, N, &m));
    int cmax = 0;
    FSGM_TRY(caller_volume_cmax(c, d_C, V, &cmax));
    FSGM_TRY(launch_sweeps(c, 1, d_C, d_I1, W, H, D, P1, P2, adaptive_p2 ? 25 : 0, cmax, directions, n_dirs, L));
    return launch_epi_wta(c, 1, L, n_dirs, W, H, D, 0, 0, nullptr, 0.0, d_Sp_partial, b, m);
}

// u8 form of the partial volume: valid when the rank's directions cannot exceed 255 together (n_dirs*(24+P2) <= 255, i.e. up
// to two directions at P2 = 64).  Exchanged with an all-to-all instead of a u16 reduce-scatter: half the NVLink bytes.
int fsgm_epi_partial_u8_dev(fsgm_ctx* c, const uint8_t* d_C, const uint8_t* d_I1, int W, int H, int D, int P1, int P2,
                            int adaptive_p2, const int* directions, int n_dirs, uint8_t* d_partial)
{
    FSGM_TRY(check_dims(c, 1, W, H, D));
    if (!d_C || !d_partial || !directions || (adaptive_p2 && !d_I1)) return fail(c, FSGM_ERR_ARG, "null pointer");
    if (n_dirs < 0 || n_dirs > 8) return fail(c, FSGM_ERR_ARG, "n_dirs must be 0..8");
    if (D % 16) return fail(c, FSGM_ERR_DOMAIN, "dMax must be a multiple of 16 for the u8 partial form");
    if ((reinterpret_cast<uintptr_t>(d_partial) & 15)) return fail(c, FSGM_ERR_ARG, "d_partial must be 16-byte aligned");
    FSGM_CUDA(c, cudaSetDevice(c->device));
    const size_t V = (size_t)W * H * D;
    int cmax = 0;
    FSGM_TRY(caller_volume_cmax(c, d_C, V, &cmax));
    if (sweep_needs_wrap(P1, P2, cmax) || n_dirs * (cmax + P2) > 255) return fail(c, FSGM_ERR_DOMAIN, "partial sums do not fit 8 bits");
    if (n_dirs == 0) { FSGM_CUDA(c, cudaMemsetAsync(d_partial, 0, V, c->stream)); return FSGM_OK; }
    FSGM_TRY(arena_reserve(c, (size_t)(n_dirs - 1) * align256(V)));
    ArenaScope scope(c);
    uint8_t* L[8];
    L[0] = d_partial;
    for (int k = 1; k < n_dirs; ++k) FSGM_TRY(arena_get(c, V, &L[k]));
    FSGM_TRY(launch_sweeps(c, 1, d_C, d_I1, W, H, D, P1, P2, adaptive_p2 ? 25 : 0, cmax, directions, n_dirs, L));
    for (int k = 1; k < n_dirs; ++k) FSGM_TRY(launch_add_u8(c, d_partial, L[k], V));
    return FSGM_OK;
}

int fsgm_epi_wta_slabs_dev(fsgm_ctx* c, const uint8_t* d_slabs, int n_slabs, const uint16_t* d_next_label0, size_t n_pixels, int D,
                           int subpixel, int vz_to_disp, const double* d_O, double vMax, uint32_t* d_bestD, uint32_t* d_minC)
{
    if (!c) return FSGM_ERR_ARG;
    if (!d_slabs || !d_bestD || !d_minC || (vz_to_disp && !d_O) || n_pixels < 1 || n_slabs < 1 || n_slabs > 8)
        return fail(c, FSGM_ERR_ARG, "bad argument");
    if (D < 1 || D > 512) return fail(c, FSGM_ERR_DOMAIN, "dMax must be in 1..512");
    FSGM_CUDA(c, cudaSetDevice(c->device));
    return launch_slab_wta(c, d_slabs, n_slabs, n_pixels * (size_t)D, d_next_label0, n_pixels, D, subpixel, vz_to_disp, d_O, vMax, d_bestD, d_minC);
}

int fsgm_epi_wta_sp_dev(fsgm_ctx* c, const uint16_t* d_Sp, const uint16_t* d_next_label0, size_t n_pixels, int D,
                        int subpixel, int vz_to_disp, const double* d_O, double vMax, uint32_t* d_bestD, uint32_t* d_minC)
{
    if (!c) return FSGM_ERR_ARG;
    if (!d_Sp || !d_bestD || !d_minC || (vz_to_disp && !d_O) || n_pixels < 1) return fail(c, FSGM_ERR_ARG, "bad argument");
    if (D < 1 || D > 512) return fail(c, FSGM_ERR_DOMAIN, "dMax must be in 1..512");
    FSGM_CUDA(c, cudaSetDevice(c->device));
    return launch_sp_wta(c, d_Sp, d_next_label0, n_pixels, D, subpixel, vz_to_disp, d_O, vMax, d_bestD, d_minC);
}

// ---------------------------------------------------------------------------------------------
// gateway 1: calc_cost_sgm
// ---------------------------------------------------------------------------------------------
int fsgm_calc_cost_sgm_dev(fsgm_ctx* c, int n, const uint8_t* d_I1, const uint8_t* d_I2, int W, int H, int D, double vMax,
                           const double* d_Pd0, const double* d_dir, const double* d_O, int P1, int P2,
                           const fsgm_epi_opts* opts, uint32_t* d_bestD, uint32_t* d_minC)
{
    FSGM_TRY(check_dims(c, n, W, H, D));
    fsgm_epi_opts o;
    FSGM_TRY(check_opts(c, opts, &o));
    if (!d_I1 || !d_I2 || !d_Pd0 || !d_dir || !d_O || !d_bestD || !d_minC) return fail(c, FSGM_ERR_ARG, "null pointer");
    FSGM_CUDA(c, cudaSetDevice(c->device));
    int dirs[8];
    const int nd = enabled_dirs(o, dirs);
    // process the batch in chunks that keep the scratch arena under ~1/3 of the device memory (queried once per
    // context: cudaMemGetInfo is a slow, synchronising call)
    if (!c->mem_total) {
        size_t free_b = 0;
        FSGM_CUDA(c, cudaMemGetInfo(&free_b, &c->mem_total));
        c->mem_budget = free_b / 3;
    }
    const size_t budget = std::max(c->arena_bytes, c->mem_budget);
    const size_t per_pair = epi_scratch_bytes(c, 1, W, H, D, P1, P2, o);
    int chunk = (int)std::min<size_t>(n, std::max<size_t>(1, budget / per_pair));
    chunk = std::min(chunk, 65535);                   // pairs ride on grid.y / grid.z of the kernels
    {   // keep chunks at whole waves of the cluster kernels when they apply (see fast_pairs)
        const int cs = fast_path_cluster(c, W, D, P1, P2, 24, o);
        if (cs && chunk < n) {
            fast_pairs(c, chunk, cs, D, W, o.paths == 8 ? 3 : 1);
            if (chunk >= c->clusters_max) chunk -= chunk % c->clusters_max;
        }
    }
    size_t need = 0;                                  // chunks differ in their fast/generic split: reserve the largest
    for (int i0 = 0; i0 < n; i0 += chunk) need = std::max(need, epi_scratch_bytes(c, std::min(chunk, n - i0), W, H, D, P1, P2, o));
    FSGM_TRY(arena_reserve(c, need));
    const size_t N = (size_t)W * H;
    for (int i0 = 0; i0 < n; i0 += chunk) {
        const int m = std::min(chunk, n - i0);
        ArenaScope scope(c);
        const int cs = c->no_overlap ? 0 : fast_path_cluster(c, W, D, P1, P2, 24, o);
        const bool waves = cs && o.total_pass == 2 && (D == 64 || D == 128 || D == 256) &&
                           fast_pairs(c, m, cs, D, W, o.paths == 8 ? 3 : 1) >= 2 * c->clusters_max - 1;
        if (waves)
            FSGM_TRY(epi_pipeline_waves(c, m, cs, d_I1 + i0 * N, d_I2 + i0 * N, W, H, D, vMax, d_Pd0 + i0 * 2 * N, d_dir + i0 * 2 * N,
                                        d_O + i0 * N, P1, P2, o, d_bestD + i0 * N, d_minC + i0 * N));
        else
        FSGM_TRY(epi_pipeline_dev(c, m, d_I1 + i0 * N, d_I2 + i0 * N, W, H, D, vMax, d_Pd0 + i0 * 2 * N, d_dir + i0 * 2 * N,
                                  d_O + i0 * N, P1, P2, o, d_bestD + i0 * N, d_minC + i0 * N));
    }
    return FSGM_OK;
}

// Enqueue-only form: returns once every copy and kernel of the batch is queued.  Host buffers must stay valid (and the
// outputs unread) until fsgm_synchronize().  Back-to-back calls overlap: the H2D copies of call k+1 run while the
// kernels of call k execute (the staging slots and their events persist in the context).
int fsgm_calc_cost_sgm_batch_async(fsgm_ctx* c, int n, const uint8_t* I1, const uint8_t* I2, int W, int H, int D, double vMax,
                                   const double* Pd0, const double* dirn, const double* O, int P1, int P2,
                                   const fsgm_epi_opts* opts, uint32_t* bestD, uint32_t* minC)
{
    FSGM_TRY(check_dims(c, n, W, H, D));
    fsgm_epi_opts o;
    FSGM_TRY(check_opts(c, opts, &o));
    if (!I1 || !I2 || !Pd0 || !dirn || !O || !bestD || !minC) return fail(c, FSGM_ERR_ARG, "null pointer");
    FSGM_CUDA(c, cudaSetDevice(c->device));
    const size_t N = (size_t)W * H;
    // Chunks of pairs flow through a 2-slot device staging buffer: H2D of chunk i+1 and D2H of chunk i-1 overlap
    // the kernels of chunk i (truly asynchronous only when the caller's buffers are pinned).
    const size_t in_pair = align256(N) * 2 + align256(2 * N * 8) * 2 + align256(N * 8);
    const size_t out_pair = 2 * align256(N * 4);
    int chunk = (int)std::max<size_t>(1, std::min<size_t>((size_t)n, (size_t(96) << 20) / (in_pair + out_pair) + 1));
    {   // when the cluster kernels apply, feed them whole waves (one pair = one cluster for a whole pass)
        const int cs = fast_path_cluster(c, W, D, P1, P2, 24, o);
        if (cs) {                                    // two waves per chunk so that the wave pipeline has something to overlap
            fast_pairs(c, n, cs, D, W, o.paths == 8 ? 3 : 1);
            chunk = std::min(n, std::max(chunk, (c->no_overlap ? 1 : c->stage_waves) * c->clusters_max));
        }
    }
    FSGM_TRY(pipe_reserve(c, (size_t)chunk * (in_pair + out_pair)));
    HostPipe& p = c->pipe;
    int rc = FSGM_OK;
    int* used = p.used;
    for (int i0 = 0; i0 < n && rc == FSGM_OK; i0 += chunk, ++p.turn) {
        const int m = std::min(chunk, n - i0), slot = p.turn & 1;
        char* base = p.buf[slot];
        uint8_t* dI1 = reinterpret_cast<uint8_t*>(base);                 base += align256(m * N);
        uint8_t* dI2 = reinterpret_cast<uint8_t*>(base);                 base += align256(m * N);
        double* dPd0 = reinterpret_cast<double*>(base);                  base += align256(m * 2 * N * 8);
        double* dDir = reinterpret_cast<double*>(base);                  base += align256(m * 2 * N * 8);
        double* dO = reinterpret_cast<double*>(base);                    base += align256(m * N * 8);
        uint32_t* dBest = reinterpret_cast<uint32_t*>(base);             base += align256(m * N * 4);
        uint32_t* dMin = reinterpret_cast<uint32_t*>(base);
        if (used[slot]) cudaStreamWaitEvent(p.h2d, p.out_ready[slot], 0);          // slot free once its outputs left
        auto H2D = [&](void* d, const void* h, size_t b) { return cudaMemcpyAsync(d, h, b, cudaMemcpyHostToDevice, p.h2d); };
        if (H2D(dI1, I1 + i0 * N, m * N) || H2D(dI2, I2 + i0 * N, m * N) || H2D(dPd0, Pd0 + i0 * 2 * N, m * 2 * N * 8) ||
            H2D(dDir, dirn + i0 * 2 * N, m * 2 * N * 8) || H2D(dO, O + i0 * N, m * N * 8) ||
            cudaEventRecord(p.in_ready[slot], p.h2d) || cudaStreamWaitEvent(c->stream, p.in_ready[slot], 0)) {
            rc = fail(c, FSGM_ERR_CUDA, "H2D copy", cudaGetErrorString(cudaGetLastError()));
            break;
        }
        rc = fsgm_calc_cost_sgm_dev(c, m, dI1, dI2, W, H, D, vMax, dPd0, dDir, dO, P1, P2, &o, dBest, dMin);
        if (rc != FSGM_OK) break;
        if (cudaEventRecord(p.done[slot], c->stream) || cudaStreamWaitEvent(p.d2h, p.done[slot], 0) ||
            cudaMemcpyAsync(bestD + i0 * N, dBest, m * N * 4, cudaMemcpyDeviceToHost, p.d2h) ||
            cudaMemcpyAsync(minC + i0 * N, dMin, m * N * 4, cudaMemcpyDeviceToHost, p.d2h) ||
            cudaEventRecord(p.out_ready[slot], p.d2h))
            rc = fail(c, FSGM_ERR_CUDA, "D2H copy", cudaGetErrorString(cudaGetLastError()));
        used[slot] = 1;
    }
    if (rc != FSGM_OK) { cudaStreamSynchronize(p.d2h); cudaStreamSynchronize(c->stream); cudaStreamSynchronize(p.h2d); }
    return rc;
}

int fsgm_calc_cost_sgm_batch(fsgm_ctx* c, int n, const uint8_t* I1, const uint8_t* I2, int W, int H, int D, double vMax,
                             const double* Pd0, const double* dirn, const double* O, int P1, int P2,
                             const fsgm_epi_opts* opts, uint32_t* bestD, uint32_t* minC)
{
    FSGM_TRY(fsgm_calc_cost_sgm_batch_async(c, n, I1, I2, W, H, D, vMax, Pd0, dirn, O, P1, P2, opts, bestD, minC));
    return fsgm_synchronize(c);
}

int fsgm_forward_backward_check_dev(fsgm_ctx* c, int n, const uint32_t* d_bestD, int W, int H, const double* d_Pd0, const double* d_dir,
                                    const double* d_O, double vMax, int nlab, int thr, int use_vzind, uint8_t* d_conf, uint32_t* d_bestD2)
{
    FSGM_TRY(check_dims(c, n, W, H, 1));
    if (!d_bestD || !d_Pd0 || !d_dir || !d_O || !d_conf || !d_bestD2) return fail(c, FSGM_ERR_ARG, "null pointer");
    if (nlab < 1) return fail(c, FSGM_ERR_ARG, "n must be positive");
    FSGM_CUDA(c, cudaSetDevice(c->device));
    return launch_fb_check(c, n, d_bestD, W, H, d_Pd0, d_dir, d_O, vMax, nlab, thr, use_vzind, d_conf, d_bestD2);
}

int fsgm_convert_vzind_to_disp_dev(fsgm_ctx* c, int n, uint32_t* d_bestD, int W, int H, const double* d_O, double vMax, int nlab)
{
    FSGM_TRY(check_dims(c, n, W, H, 1));
    if (!d_bestD || !d_O) return fail(c, FSGM_ERR_ARG, "null pointer");
    if (nlab < 1) return fail(c, FSGM_ERR_ARG, "n must be positive");
    FSGM_CUDA(c, cudaSetDevice(c->device));
    return launch_vz_to_disp(c, d_bestD, d_O, (size_t)n * W * H, vMax, nlab);
}

// One pair with the reference's commented-out call (calc_cost_sgm.cpp:589-590) switched back on: labels without the vz
// conversion, the check, then the conversion (:592-594).  Synchronous, plain copies: this is not the throughput path.
static int epi_with_fb_check(fsgm_ctx* c, const uint8_t* I1, const uint8_t* I2, int W, int H, int D, double vMax,
                             const double* Pd0, const double* dirn, const double* O, int P1, int P2, fsgm_epi_opts o,
                             uint32_t* bestD, uint32_t* minC, uint8_t* conf, uint32_t* bestD2)
{
    FSGM_TRY(check_dims(c, 1, W, H, D));
    if (!I1 || !I2 || !Pd0 || !dirn || !O || !bestD || !minC) return fail(c, FSGM_ERR_ARG, "null pointer");
    FSGM_CUDA(c, cudaSetDevice(c->device));
    const size_t N = (size_t)W * H;
    const size_t in_b = align256(N) * 2 + align256(2 * N * 8) * 2 + align256(N * 8);
    const size_t out_b = 3 * align256(N * 4) + align256(N);
    FSGM_TRY(pipe_reserve(c, in_b + out_b));
    FSGM_TRY(fsgm_synchronize(c));                               // an earlier asynchronous call may still own the slot
    char* base = c->pipe.buf[0];
    uint8_t* dI1 = reinterpret_cast<uint8_t*>(base);   base += align256(N);
    uint8_t* dI2 = reinterpret_cast<uint8_t*>(base);   base += align256(N);
    double* dPd0 = reinterpret_cast<double*>(base);    base += align256(2 * N * 8);
    double* dDir = reinterpret_cast<double*>(base);    base += align256(2 * N * 8);
    double* dO = reinterpret_cast<double*>(base);      base += align256(N * 8);
    uint32_t* dBest = reinterpret_cast<uint32_t*>(base); base += align256(N * 4);
    uint32_t* dMin = reinterpret_cast<uint32_t*>(base);  base += align256(N * 4);
    uint32_t* dB2 = reinterpret_cast<uint32_t*>(base);   base += align256(N * 4);
    uint8_t* dConf = reinterpret_cast<uint8_t*>(base);
    cudaStream_t s = c->stream;
    FSGM_CUDA(c, cudaMemcpyAsync(dI1, I1, N, cudaMemcpyHostToDevice, s));
    FSGM_CUDA(c, cudaMemcpyAsync(dI2, I2, N, cudaMemcpyHostToDevice, s));
    FSGM_CUDA(c, cudaMemcpyAsync(dPd0, Pd0, 2 * N * 8, cudaMemcpyHostToDevice, s));
    FSGM_CUDA(c, cudaMemcpyAsync(dDir, dirn, 2 * N * 8, cudaMemcpyHostToDevice, s));
    FSGM_CUDA(c, cudaMemcpyAsync(dO, O, N * 8, cudaMemcpyHostToDevice, s));
    const int vz = o.vz_to_disp;
    o.vz_to_disp = 0; o.fb_check = 0;
    FSGM_TRY(fsgm_calc_cost_sgm_dev(c, 1, dI1, dI2, W, H, D, vMax, dPd0, dDir, dO, P1, P2, &o, dBest, dMin));
    FSGM_TRY(launch_fb_check(c, 1, dBest, W, H, dPd0, dDir, dO, vMax, D + 1, o.fb_thr, /*USE_VZIND :4*/ 1, dConf, dB2));
    if (vz) FSGM_TRY(launch_vz_to_disp(c, dBest, dO, N, vMax, D + 1));
    FSGM_CUDA(c, cudaMemcpyAsync(bestD, dBest, N * 4, cudaMemcpyDeviceToHost, s));
    FSGM_CUDA(c, cudaMemcpyAsync(minC, dMin, N * 4, cudaMemcpyDeviceToHost, s));
    if (conf) FSGM_CUDA(c, cudaMemcpyAsync(conf, dConf, N, cudaMemcpyDeviceToHost, s));
    if (bestD2) FSGM_CUDA(c, cudaMemcpyAsync(bestD2, dB2, N * 4, cudaMemcpyDeviceToHost, s));
    FSGM_CUDA(c, cudaStreamSynchronize(s));
    return FSGM_OK;
}

int fsgm_calc_cost_sgm(fsgm_ctx* c, const uint8_t* I1, const uint8_t* I2, int W, int H, int D, double vMax,
                       const double* Pd0, const double* dirn, const double* O, int P1, int P2,
                       const fsgm_epi_opts* opts, uint32_t* bestD, uint32_t* minC, uint8_t* conf, uint32_t* bestD2)
{
    if (opts && opts->fb_check) return epi_with_fb_check(c, I1, I2, W, H, D, vMax, Pd0, dirn, O, P1, P2, *opts, bestD, minC, conf, bestD2);
    int rc = fsgm_calc_cost_sgm_batch(c, 1, I1, I2, W, H, D, vMax, Pd0, dirn, O, P1, P2, opts, bestD, minC);
    if (rc != FSGM_OK) return rc;
    // outputs 3 and 4 of the gateway are allocated but never written by the reference (:571-572, :589-590)
    if (conf) std::memset(conf, 0, (size_t)W * H);
    if (bestD2) std::memset(bestD2, 0, (size_t)W * H * sizeof(uint32_t));
    return FSGM_OK;
}

// ---------------------------------------------------------------------------------------------
// N2: dense epipolar prologue / epilogue (rotation_motion.m, epipolar_geometry.m:104-119, epipolar_sgm_of.m:46-51)
// ---------------------------------------------------------------------------------------------
int fsgm_epipolar_geometry_dev(fsgm_ctx* c, int n, const double* F, const double* Hm, const double* epipole, const int* direction,
                               int W, int H, double* d_Pd0, double* d_dir, double* d_O, double* d_Rflow)
{
    FSGM_TRY(check_dims(c, n, W, H, 1));
    if (!F || !Hm || !epipole || !d_Pd0 || !d_dir || !d_O) return fail(c, FSGM_ERR_ARG, "null pointer");
    if (n > 65535) return fail(c, FSGM_ERR_DOMAIN, "at most 65535 pairs per geometry call");
    FSGM_CUDA(c, cudaSetDevice(c->device));
    return launch_geo_prologue(c, n, F, Hm, epipole, direction, W, H, d_Pd0, d_dir, d_O, d_Rflow);
}

int fsgm_epipolar_flow_dev(fsgm_ctx* c, int n, const uint32_t* d_bestD, const double* d_dir, const double* d_Rflow, int W, int H,
                           double* d_flow)
{
    FSGM_TRY(check_dims(c, n, W, H, 1));
    if (!d_bestD || !d_dir || !d_Rflow || !d_flow) return fail(c, FSGM_ERR_ARG, "null pointer");
    if (n > 65535) return fail(c, FSGM_ERR_DOMAIN, "at most 65535 pairs per geometry call");
    FSGM_CUDA(c, cudaSetDevice(c->device));
    return launch_geo_epilogue(c, n, d_bestD, d_dir, d_Rflow, W, H, d_flow);
}

// d_work: 60 bytes per pixel and pair of caller-provided scratch (Pd0 16, direction 16, offset 8, Rflow 16, bestD 4);
// exactly one of d_flow (f64 planes) and d_flow32 (float, interleaved u,v) is set
static int sgm_of_dev(fsgm_ctx* c, int n, const uint8_t* d_I0, const uint8_t* d_I1, int W, int H,
                      const double* F, const double* Hm, const double* epipole, const int* direction,
                      int D, double vMax, int P1, int P2, const fsgm_epi_opts* opts, void* d_work, double* d_flow, float* d_flow32,
                      uint32_t* d_minC)
{
    FSGM_TRY(check_dims(c, n, W, H, D));
    if (!d_I0 || !d_I1 || !F || !Hm || !epipole || !d_work || (!d_flow && !d_flow32) || !d_minC) return fail(c, FSGM_ERR_ARG, "null pointer");
    fsgm_epi_opts o;
    FSGM_TRY(check_opts(c, opts, &o));
    if (!o.vz_to_disp) return fail(c, FSGM_ERR_ARG, "epipolar_sgm_of needs pixel disparities (opts.vz_to_disp = 1)");
    if (n > 65535) return fail(c, FSGM_ERR_DOMAIN, "at most 65535 pairs per call of the device form (the host form chunks)");
    FSGM_CUDA(c, cudaSetDevice(c->device));
    const size_t N = (size_t)W * H;
    char* base = static_cast<char*>(d_work);
    double* Pd0 = reinterpret_cast<double*>(base);   base += (size_t)n * 2 * N * 8;
    double* dirn = reinterpret_cast<double*>(base);  base += (size_t)n * 2 * N * 8;
    double* Rf = reinterpret_cast<double*>(base);    base += (size_t)n * 2 * N * 8;
    double* O = reinterpret_cast<double*>(base);     base += (size_t)n * N * 8;
    uint32_t* best = reinterpret_cast<uint32_t*>(base);
    FSGM_TRY(launch_geo_prologue(c, n, F, Hm, epipole, direction, W, H, Pd0, dirn, O, Rf));
    FSGM_TRY(fsgm_calc_cost_sgm_dev(c, n, d_I0, d_I1, W, H, D, vMax, Pd0, dirn, O, P1, P2, &o, best, d_minC));
    if (d_flow32) return launch_geo_epilogue_f32(c, n, best, dirn, Rf, W, H, d_flow32);
    return launch_geo_epilogue(c, n, best, dirn, Rf, W, H, d_flow);
}

int fsgm_epipolar_sgm_of_dev(fsgm_ctx* c, int n, const uint8_t* d_I0, const uint8_t* d_I1, int W, int H,
                             const double* F, const double* Hm, const double* epipole, const int* direction,
                             int D, double vMax, int P1, int P2, const fsgm_epi_opts* opts, void* d_work, double* d_flow, uint32_t* d_minC)
{
    return sgm_of_dev(c, n, d_I0, d_I1, W, H, F, Hm, epipole, direction, D, vMax, P1, P2, opts, d_work, d_flow, nullptr, d_minC);
}

int fsgm_epipolar_sgm_of_f32_dev(fsgm_ctx* c, int n, const uint8_t* d_I0, const uint8_t* d_I1, int W, int H,
                                 const double* F, const double* Hm, const double* epipole, const int* direction,
                                 int D, double vMax, int P1, int P2, const fsgm_epi_opts* opts, void* d_work, float* d_flow, uint32_t* d_minC)
{
    return sgm_of_dev(c, n, d_I0, d_I1, W, H, F, Hm, epipole, direction, D, vMax, P1, P2, opts, d_work, nullptr, d_flow, d_minC);
}

size_t fsgm_epipolar_sgm_of_work_bytes(int n_pairs, int W, int H) { return (size_t)n_pairs * W * H * 60; }

// Host-pointer form, enqueue-only like fsgm_calc_cost_sgm_batch_async: 2 bytes per pixel go up, 20 come back.
// The float form returns 8 (flow, interleaved like CV_32FC2) + 4 (minC, optional) bytes per pixel.
static int sgm_of_batch_async(fsgm_ctx* c, int n, const uint8_t* I0, const uint8_t* I1, int W, int H,
                              const double* F, const double* Hm, const double* epipole, const int* direction,
                              int D, double vMax, int P1, int P2, const fsgm_epi_opts* opts, double* flow, float* flow32, uint32_t* minC)
{
    FSGM_TRY(check_dims(c, n, W, H, D));
    fsgm_epi_opts o;
    FSGM_TRY(check_opts(c, opts, &o));
    if (!I0 || !I1 || !F || !Hm || !epipole || (!flow && !flow32) || (!minC && !flow32)) return fail(c, FSGM_ERR_ARG, "null pointer");
    FSGM_CUDA(c, cudaSetDevice(c->device));
    const size_t N = (size_t)W * H;
    const size_t per_pair = 2 * align256(N) + 60 * N + 256 + align256(2 * N * 8) + align256(N * 4);
    int chunk = (int)std::max<size_t>(1, std::min<size_t>((size_t)n, (size_t(96) << 20) / per_pair + 1));
    {
        const int cs = fast_path_cluster(c, W, D, P1, P2, 24, o);
        if (cs) {
            fast_pairs(c, n, cs, D, W, o.paths == 8 ? 3 : 1);
            chunk = std::min(n, std::max(chunk, (c->no_overlap ? 1 : c->stage_waves) * c->clusters_max));
        }
    }
    FSGM_TRY(pipe_reserve(c, (size_t)chunk * per_pair));
    HostPipe& p = c->pipe;
    int rc = FSGM_OK;
    for (int i0 = 0; i0 < n && rc == FSGM_OK; i0 += chunk, ++p.turn) {
        const int m = std::min(chunk, n - i0), slot = p.turn & 1;
        char* base = p.buf[slot];
        uint8_t* dI0 = reinterpret_cast<uint8_t*>(base);   base += align256(m * N);
        uint8_t* dI1 = reinterpret_cast<uint8_t*>(base);   base += align256(m * N);
        void* work = base;                                 base += align256(m * N * 60);
        double* dFlow = reinterpret_cast<double*>(base);   base += align256(m * 2 * N * 8);
        uint32_t* dMin = reinterpret_cast<uint32_t*>(base);
        if (p.used[slot]) cudaStreamWaitEvent(p.h2d, p.out_ready[slot], 0);
        if (cudaMemcpyAsync(dI0, I0 + i0 * N, m * N, cudaMemcpyHostToDevice, p.h2d) ||
            cudaMemcpyAsync(dI1, I1 + i0 * N, m * N, cudaMemcpyHostToDevice, p.h2d) ||
            cudaEventRecord(p.in_ready[slot], p.h2d) || cudaStreamWaitEvent(c->stream, p.in_ready[slot], 0)) {
            rc = fail(c, FSGM_ERR_CUDA, "H2D copy", cudaGetErrorString(cudaGetLastError()));
            break;
        }
        // the slot's scratch is rewritten by this chunk's kernels: they must not start before the slot's previous outputs left
        if (p.used[slot]) cudaStreamWaitEvent(c->stream, p.out_ready[slot], 0);
        rc = sgm_of_dev(c, m, dI0, dI1, W, H, F + (size_t)i0 * 9, Hm + (size_t)i0 * 9, epipole + (size_t)i0 * 2,
                        direction ? direction + i0 : nullptr, D, vMax, P1, P2, &o, work, flow32 ? nullptr : dFlow,
                        flow32 ? reinterpret_cast<float*>(dFlow) : nullptr, dMin);
        if (rc != FSGM_OK) break;
        if (cudaEventRecord(p.done[slot], c->stream) || cudaStreamWaitEvent(p.d2h, p.done[slot], 0) ||
            (flow32 ? cudaMemcpyAsync(flow32 + (size_t)i0 * 2 * N, dFlow, m * 2 * N * 4, cudaMemcpyDeviceToHost, p.d2h)
                    : cudaMemcpyAsync(flow + (size_t)i0 * 2 * N, dFlow, m * 2 * N * 8, cudaMemcpyDeviceToHost, p.d2h)) ||
            (minC ? cudaMemcpyAsync(minC + i0 * N, dMin, m * N * 4, cudaMemcpyDeviceToHost, p.d2h) : cudaSuccess) ||
            cudaEventRecord(p.out_ready[slot], p.d2h))
            rc = fail(c, FSGM_ERR_CUDA, "D2H copy", cudaGetErrorString(cudaGetLastError()));
        p.used[slot] = 1;
    }
    if (rc != FSGM_OK) { cudaStreamSynchronize(p.d2h); cudaStreamSynchronize(c->stream); cudaStreamSynchronize(p.h2d); }
    return rc;
}

int fsgm_epipolar_sgm_of_batch_async(fsgm_ctx* c, int n, const uint8_t* I0, const uint8_t* I1, int W, int H,
                                     const double* F, const double* Hm, const double* epipole, const int* direction,
                                     int D, double vMax, int P1, int P2, const fsgm_epi_opts* opts, double* flow, uint32_t* minC)
{
    return sgm_of_batch_async(c, n, I0, I1, W, H, F, Hm, epipole, direction, D, vMax, P1, P2, opts, flow, nullptr, minC);
}

int fsgm_epipolar_sgm_of_f32_batch_async(fsgm_ctx* c, int n, const uint8_t* I0, const uint8_t* I1, int W, int H,
                                         const double* F, const double* Hm, const double* epipole, const int* direction,
                                         int D, double vMax, int P1, int P2, const fsgm_epi_opts* opts, float* flow, uint32_t* minC)
{
    return sgm_of_batch_async(c, n, I0, I1, W, H, F, Hm, epipole, direction, D, vMax, P1, P2, opts, nullptr, flow, minC);
}

int fsgm_epipolar_sgm_of(fsgm_ctx* c, const uint8_t* I0, const uint8_t* I1, int W, int H, const double* F, const double* Hm,
                         const double* epipole, int direction, int D, double vMax, int P1, int P2, const fsgm_epi_opts* opts,
                         double* flow, uint32_t* minC)
{
    FSGM_TRY(fsgm_epipolar_sgm_of_batch_async(c, 1, I0, I1, W, H, F, Hm, epipole, &direction, D, vMax, P1, P2, opts, flow, minC));
    return fsgm_synchronize(c);
}

}  // extern "C"
