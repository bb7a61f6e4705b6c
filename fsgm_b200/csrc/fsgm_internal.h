// Internal declarations shared by the kernels and the C-ABI layer.  Not installed.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stddef.h>
#include <string>
#include <vector>
#include <utility>
#include "../../include/fsgm.h"

namespace fsgm {
// pipeline stages, for the optional per-stage CUDA-event timing (fsgm_profile_*)
enum Stage { ST_CENSUS = 0, ST_EPI_COST, ST_SWEEP, ST_WTA, ST_PYD_COST, ST_PYD_SWEEP, ST_PYD_WTA,
             ST_NG, ST_PYDNG_COST, ST_PYDNG_SWEEP, ST_PYDNG_WTA, ST_MISC, ST_VSWEEP, ST_PYRAMID, ST_GEOMETRY, ST_EXCHANGE, ST_COUNT };
struct StageTimer { cudaEvent_t a, b; int stage; };
// double-buffered device staging + copy streams for the host-pointer gateways
struct HostPipe {
    cudaStream_t h2d = nullptr, d2h = nullptr;
    cudaEvent_t in_ready[2] = {nullptr, nullptr}, done[2] = {nullptr, nullptr}, out_ready[2] = {nullptr, nullptr};
    char* buf[2] = {nullptr, nullptr};
    size_t bytes = 0;
    int used[2] = {0, 0};        // slot has been used since the last (re)allocation: its out_ready event is meaningful
    unsigned turn = 0;           // running chunk counter: slots alternate across calls
};
}  // namespace fsgm

struct fsgm_ctx {
    int device = 0;
    cudaStream_t own_stream = nullptr;
    cudaStream_t stream = nullptr;        // the stream kernels are launched on
    char* arena = nullptr;                // scratch, grown on demand, reused across calls
    size_t arena_bytes = 0;
    size_t arena_top = 0;
    uint64_t launches = 0;
    int sm_count = 0;
    size_t mem_total = 0, mem_budget = 0;   // cudaMemGetInfo, queried once
    cudaStream_t aux_stream = nullptr;      // second stream: front-end of wave i+1 under the cluster kernels of wave i
    cudaEvent_t ev_entry = nullptr, ev_front[2] = {nullptr, nullptr};
    int no_overlap = 0;                     // tuning knob (fsgm_tune key 2)
    int stage_waves = 3;                    // tuning knob (fsgm_tune key 9): cluster waves per staging chunk of the host-pointer gateways
    int fc_rows = 0;                        // tuning knob (fsgm_tune key 8): rows per CTA of the fused cost kernel, 0 = chosen from the grid size
    unsigned attr_mask = 0;                 // which kernels already had their max-dynamic-smem attribute set on this device
    int clusters_key[4] = {0, 0, 0, 0}, clusters_max = 0;   // resident clusters for the last queried (cluster size, W, D, ndir)
    int best_key[3] = {0, 0, 0}, best_cs = 0, best_clusters = 0;   // cached vsweep_best_cluster() decision for (W, D, ndir)
    int pyd_direct_cost = 0;                // tuning knob (fsgm_tune key 6): 1 = the lane = path pipeline builds its cost volume with the direct kernel only
    int pyd_cluster = 0;                    // tuning knob (fsgm_tune key 5): cluster size of the pyd row-synchronous kernels, 0 auto, -1 off
    std::vector<std::pair<int, int>> pv_occ;   // cached cudaOccupancyMaxActiveClusters answers of pydv_kernel
    int pydng_generic = 0;                  // tuning knob (fsgm_tune key 7): 1 = calc_pyd_cost_sgm_ng searches cell by cell at every step (no grid tables)
    int ng_occupancy = 0;                   // tuning knob (fsgm_tune key 3): resident pairs per SM of the ng kernel, 0 = by batch size
    int force_cluster = 0;                  // tuning/test knob: 0 auto, -1 generic path only, 1/2/4/8 forced cluster size
    std::string err;
    // profiling
    bool profiling = false;
    std::vector<fsgm::StageTimer> timers;      // recorded, not yet read back
    std::vector<cudaEvent_t> event_pool;
    double stage_ms[fsgm::ST_COUNT] = {};
    uint64_t stage_launches[fsgm::ST_COUNT] = {};
    fsgm::HostPipe pipe;
    // multi-GPU (dist.cu): the context's NCCL communicator, one rank per context
    void* nccl_comm = nullptr;
    bool nccl_owned = false;
    int rank = 0, world = 1;
    // peer-mapped receive buffers of the direction split: p2p_peer[j] = rank j's buffer (cudaIpc mapping; [rank] = p2p_local)
    void* p2p_local = nullptr;
    size_t p2p_bytes = 0;
    void* p2p_peer[16] = {};
    int p2p_state = 0;                      // 0 untried, 1 usable, -1 peer access unavailable (NCCL exchange instead)
    int no_p2p = 0;                         // tuning knob (fsgm_tune key 4): 1 = never use the peer-store form
    void* d_scalar = nullptr;               // 256 B of device memory for small read-backs (launch_max_u8)
    void* geo_params = nullptr;             // per-pair F, H, epipole, direction of the dense-geometry prologue (geometry.cu)
    size_t geo_cap = 0;
    void* geo_host = nullptr;               // pinned staging ring for them (4 slots of geo_host_cap pairs)
    size_t geo_host_cap = 0;
    cudaEvent_t geo_ev[4] = {nullptr, nullptr, nullptr, nullptr};
    unsigned geo_turn = 0;
};

namespace fsgm {

// ---- error plumbing ---------------------------------------------------------------------------
int  fail(fsgm_ctx* c, int code, const char* what, const char* detail = nullptr);
#define FSGM_CUDA(ctx, expr)                                                         \
    do { cudaError_t e__ = (expr);                                                   \
         if (e__ != cudaSuccess) return fsgm::fail((ctx), FSGM_ERR_CUDA, #expr, cudaGetErrorString(e__)); } while (0)
#define FSGM_TRY(expr) do { int rc__ = (expr); if (rc__ != FSGM_OK) return rc__; } while (0)
#define FSGM_LAUNCHED(ctx)                                                           \
    do { (ctx)->launches++; cudaError_t e__ = cudaGetLastError();                    \
         if (e__ != cudaSuccess) return fsgm::fail((ctx), FSGM_ERR_CUDA, "kernel launch", cudaGetErrorString(e__)); } while (0)

// ---- per-stage timing: brackets the enclosed launches with events when profiling is on -----------------
struct StageScope {
    fsgm_ctx* c; int stage; cudaEvent_t a = nullptr; uint64_t launches0;
    StageScope(fsgm_ctx* ctx, int st);
    ~StageScope();
};
int  pipe_reserve(fsgm_ctx* c, size_t bytes_per_slot);

// ---- scratch arena: bump allocator, reset per top-level call ----------------------------------
struct ArenaScope {                       // restores the arena top on scope exit
    fsgm_ctx* c; size_t saved;
    explicit ArenaScope(fsgm_ctx* ctx) : c(ctx), saved(ctx->arena_top) {}
    ~ArenaScope() { c->arena_top = saved; }
};
int arena_reserve(fsgm_ctx* c, size_t total_bytes);            // make sure the arena holds this much
int arena_alloc(fsgm_ctx* c, size_t bytes, void** out);        // 256-B aligned slice; fails if not reserved
template <class T> inline int arena_get(fsgm_ctx* c, size_t count, T** out) {
    return arena_alloc(c, count * sizeof(T), reinterpret_cast<void**>(out));
}
inline size_t align256(size_t b) { return (b + 255) & ~size_t(255); }

// ---- direction table (same order as include/fsgm.h documents) ---------------------------------
//   r: 0 L1(+1,0)  1 L3(0,+1)  2 L2(+1,+1)  3 L4(-1,+1)   4..7 = negated (pass 1)
__host__ __device__ inline int dir_dx(int r) { const int t[8] = {1, 0, 1, -1, -1, 0, -1, 1}; return t[r]; }
__host__ __device__ inline int dir_dy(int r) { const int t[8] = {0, 1, 1, 1, 0, -1, -1, -1}; return t[r]; }
inline bool dir_enabled(int r, int total_pass, bool diag) {
    if (r >= 4 && total_pass < 2) return false;
    if ((r & 3) >= 2 && !diag) return false;
    return total_pass >= 1;
}

// ---- argument checks shared by the entry points (api.cu) -----------------------------------------------
int check_dims(fsgm_ctx* c, int n, int W, int H, int D);
int check_opts(fsgm_ctx* c, const fsgm_epi_opts* in, fsgm_epi_opts* o);
int enabled_dirs(const fsgm_epi_opts& o, int* dirs);
void dist_release(fsgm_ctx* c);                                 // dist.cu: drops the context's NCCL communicator

// ---- kernel launchers (definitions in the .cu files) -------------------------------------------
int launch_census(fsgm_ctx* c, int n_images, const uint8_t* img, int W, int H, uint32_t* cen);
int launch_epi_cost_fused(fsgm_ctx* c, int n, double vMax, const uint32_t* cen1, const uint32_t* cen2, int W, int H, int D,
                          const double* Pd0, const double* dirn, const double* O, uint8_t* C, bool* done);
int launch_vz_table(fsgm_ctx* c, int D, double vMax, double* d_vz);
int launch_epi_cost(fsgm_ctx* c, int n, const double* d_vz, const uint32_t* cen1, const uint32_t* cen2, int W, int H, int D, double vMax,
                    const double* Pd0, const double* dirn, const double* O, uint8_t* raw, uint8_t* C);
// all enabled directions in one launch; Lvols[k] is the output volume of the k-th enabled direction
// cmax = upper bound on the values in C (24 for anything built from 5x5 census; 255 = unknown) — selects the
// exact-u16 or the explicit mod-256 instantiation (see aggregate.cu)
int launch_sweeps(fsgm_ctx* c, int n, const uint8_t* C, const uint8_t* I1, int W, int H, int D,
                  int P1, int P2, int adaptive_thr, int cmax, const int* dirs, int n_dirs, uint8_t* const* Lvols);
bool sweep_needs_wrap(int P1, int P2, int cmax);
int launch_sweeps_scatter(fsgm_ctx* c, const uint8_t* C, int W, int H, int D, int P1, int P2, const int* dirs, const int* slots,
                          int n_dirs, uint8_t* const* peer, int world, size_t stripe_pixels, size_t local_pixels);
// sum of n_dirs L volumes -> WTA -> subpixel -> (optionally) vz->disparity; Sp16 (u16 [n][N][D]) may be null
int launch_epi_wta(fsgm_ctx* c, int n, uint8_t* const* Lvols, int n_dirs, int W, int H, int D, int subpixel,
                   int vz_to_disp, const double* O, double vMax, uint16_t* Sp16, uint32_t* bestD, uint32_t* minC);

int launch_slab_wta(fsgm_ctx* c, const uint8_t* vols, int n_vols, size_t vol_stride, const uint16_t* next0, size_t npix, int D, int subpixel,
                    int vz_to_disp, const double* O, double vMax, uint32_t* bestD, uint32_t* minC);
int launch_add_u8(fsgm_ctx* c, uint8_t* a, const uint8_t* b, size_t bytes);
int launch_max_u8(fsgm_ctx* c, const uint8_t* v, size_t bytes, int* cmax);   // synchronous read-back
int launch_sp_wta(fsgm_ctx* c, const uint16_t* Sp, const uint16_t* next0, size_t npix, int D, int subpixel,
                  int vz_to_disp, const double* O, double vMax, uint32_t* bestD, uint32_t* minC);

// ---- forward/backward check and the stand-alone vz conversion (fbcheck.cu) ---------------------------------
int launch_fb_check(fsgm_ctx* c, int n_pairs, const uint32_t* D1, int W, int H, const double* Pd0, const double* dirn, const double* O,
                    double vMax, int n, int thr, int use_vzind, uint8_t* conf, uint32_t* D2);
int launch_vz_to_disp(fsgm_ctx* c, uint32_t* D, const double* O, size_t total, double vMax, int n);

// ---- row-synchronous cluster path for the non-horizontal directions (vsweep.cu) --------------------------
int vsweep_cluster_size(int W, int D, int ndir, int max_smem);
bool vsweep_cluster_ok(int W, int D, int ndir, int cs, int max_smem);
int vsweep_best_cluster(int W, int D, int ndir, int max_smem, int* clusters);
int vsweep_max_clusters(int cs, size_t smem, int threads);
size_t vsweep_smem_bytes(int D, int Wk, int ndir);
int vsweep_threads();
int launch_vsweep(fsgm_ctx* c, int n, int cs, int ndir, bool final_, const uint8_t* C, const uint8_t* addA, const uint8_t* addB,
                  const uint16_t* Sin, uint16_t* Sout, uint32_t* minC, uint16_t* rec, int W, int H, int D, int P1, int P2, int up, bool fast);
bool vsweep_fast_ok(int ndir, int P2);
int launch_vs_finalize(fsgm_ctx* c, int n, const uint16_t* rec, const uint32_t* minC, const double* O, int W, int H, int D,
                       int subpixel, int vz_to_disp, double vMax, int biased, uint32_t* bestD);

// ---- pyramidal 2-D-window variant (pyd.cu) ----------------------------------------------------------
int launch_pyd_cost(fsgm_ctx* c, int n, const uint32_t* cen1, const uint32_t* cen2, int W, int H,
                    const double* preMv, int mvW, int mvH, int agg, int rx, int ry, uint8_t* C, int pitch = 0, int soa = 0,
                    const uint32_t* list = nullptr, const uint32_t* list_count = nullptr);
int launch_pyd_cost_sep(fsgm_ctx* c, int n, const uint32_t* cen1, const uint32_t* cen2, int W, int H,
                        const double* preMv, int mvW, int mvH, int agg, int rx, int ry, uint8_t* C, uint32_t* list, uint32_t* count);
int launch_pyd_sweeps(fsgm_ctx* c, int n, const uint8_t* C, const uint8_t* I1, const double* preMv, int mvW, int mvH,
                      int W, int H, int Sx, int Sy, int P1, int P2, int adaptive, const int* dirs, int n_dirs, uint8_t* const* Lvols,
                      int pitch = 0);
int launch_pyd_wta(fsgm_ctx* c, int n, uint8_t* const* Lvols, const int* weights, int n_dirs, int W, int H, int Sx, int Sy,
                   int subpixel, uint16_t* Sp16, uint32_t* bestD, uint32_t* minC, double* mvSub);

// ---- row-synchronous cluster path of the pyramidal variant (pydv.cu) -----------------------------------------------
bool pydv_applicable(int Sx, int Sy, int P1, int P2, int diag, int passes, int adaptive);
int pydv_pick_cluster(fsgm_ctx* c, int n, int W, int Sx, int forced);
int launch_pyd_shift_flags(fsgm_ctx* c, int n, const double* preMv, int mvW, int mvH, int W, int H, uint8_t* flags);
int launch_pydv(fsgm_ctx* c, int n, int cs, bool final_, const uint8_t* C, const uint8_t* H0, const uint8_t* H1, const uint8_t* S1in,
                uint8_t* S1out, uint4* rec, const uint8_t* flags, const double* preMv, int mvW, int mvH, int W, int H, int Sx, int Sy,
                int P1, int P2);
int launch_pydv_finalize(fsgm_ctx* c, int n, const uint4* rec, int W, int H, int Sx, int Sy, int subpixel,
                         uint32_t* bestD, uint32_t* minC, double* mvSub);

// ---- lane = path aggregation of the pyramidal variant (pydl.cu): volumes as [y][label column][x][16-byte frame] --------------
bool pydl_applicable(int Sx, int Sy, int P1, int P2, int n_dirs, const int* weights);
int launch_pydl_desc(fsgm_ctx* c, int n, const double* preMv, int mvW, int mvH, int W, int H, int Sx, int Sy, const int* dirs, int n_dirs,
                     uint32_t* const* out, int force_generic);
int launch_pydl_sweeps(fsgm_ctx* c, int n, const uint8_t* C, const uint8_t* I1, const double* preMv, int mvW, int mvH, int W, int H,
                       int Sx, int Sy, int P1, int P2, int adaptive, const int* dirs, int n_dirs, uint32_t* const* desc, uint8_t* const* Lvols);
int launch_pydl_wta(fsgm_ctx* c, int n, uint8_t* const* Lvols, int n_dirs, int W, int H, int Sx, int Sy, int subpixel,
                    uint32_t* bestD, uint32_t* minC, double* mvSub);

// ---- dense epipolar prologue / epilogue (geometry.cu): F, H, epipole -> Pd0, direction, offset, Rflow; labels -> flow ------
int launch_geo_prologue(fsgm_ctx* c, int n, const double* F, const double* Hm, const double* epi, const int* direction, int W, int H,
                        double* Pd0, double* dirn, double* O, double* Rflow);
int launch_geo_epilogue(fsgm_ctx* c, int n, const uint32_t* bestD, const double* dirn, const double* Rflow, int W, int H, double* flow);
int launch_geo_epilogue_f32(fsgm_ctx* c, int n, const uint32_t* bestD, const double* dirn, const double* Rflow, int W, int H, float* flow);

// ---- pyramid driver (pyramid.cu): impyramid 'reduce', label -> mv, 2 x nearest upsample ------------------------
int launch_pyr_reduce(fsgm_ctx* c, int n_images, const uint8_t* in, int W, int H, uint8_t* out);
int launch_pyr_label_to_mv(fsgm_ctx* c, int n, const uint32_t* label, const double* preMv, int mvW, int mvH, const double* mvSub,
                           int W, int H, int rx, int ry, double* mv);
int launch_pyr_upsample2(fsgm_ctx* c, int n, const double* mv, int W, int H, double* out);

}  // namespace fsgm
