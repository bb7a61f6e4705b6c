// Multi-GPU entry points of the C ABI: one fsgm_ctx per GPU / rank, an NCCL communicator owned by (or lent to)
// the context, and the two partitionings the path has:
//
//   * a batch of independent pairs  -> fsgm_shard_range(): rank r owns a contiguous block of pairs, no data-path collective;
//     fsgm_dist_allgather_u32() is there for callers that want the sharded outputs everywhere.
//   * ONE large pair, the scan directions of sgm() (calc_cost_sgm.cpp:114-257) split over the ranks ->
//     fsgm_calc_cost_sgm_dirsplit_dev(): every rank builds the cost volume (cheap, redundant), sweeps only its own directions
//     into a partial volume, the partial volumes are summed per pixel slab over NVLink, every rank runs winner-take-all +
//     subpixel (:259-308, :414-426) on its slab and the 8 B/pixel outputs are all-gathered.
//       - u8 exchange (every rank's directions fit a byte together, i.e. world >= 4 at P2 = 64): the partial volume is a u8
//         volume; slab j of every rank goes to rank j in ONE grouped ncclSend/ncclRecv exchange (an all-to-all: half the
//         bytes of the u16 reduce-scatter) and the received slabs are summed inside the WTA kernel;
//       - u16 exchange otherwise: ncclReduceScatter(sum) on the u16 pairs typed ncclUint32 — per-voxel totals are at most
//         8*255 < 65536, so no carry crosses a half-word and the integer sum is exact.
//       - peer-store form (default when the GPUs can map each other's memory, i.e. one NVSwitch box): no exchange step at all.
//         Every rank's receive buffer is mapped into the other ranks (cudaIpc), and the sweep kernel writes each pixel's L row
//         straight into the memory of the rank that owns the pixel's slab (sweep_fast_kernel<.., SCATTER>, aggregate.cu): the
//         NVLink transfer runs under the sweep's own arithmetic, tile by tile, instead of after it.  One 4-byte all-gather
//         orders the stores before the owners' WTA.  The two exchange forms above remain as the fallback (no peer access,
//         mod-256 parameter domain, adaptive P2) and for A/B (fsgm_tune key 4).
//     The reference's read of the NEXT pixel's label 0 for argmin == dMax-1 (:293-296) crosses slab boundaries: the first voxel
//     of every slab is all-gathered (one word per rank) and handed to the previous rank's WTA.
//
// NCCL is bound at run time (dlopen of libnccl.so.2: inside a torch process that is the copy torch already loaded), so
// libfsgm.so has no link-time dependency on it and single-GPU users never touch it.  Every failure is reported as FSGM_ERR_NCCL.
#include "fsgm_internal.h"
#include <nccl.h>
#include <dlfcn.h>
#include <algorithm>
#include <cstring>
#include <mutex>

namespace fsgm {

namespace {

struct NcclApi {
    void* lib = nullptr;
    ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
    ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    ncclResult_t (*AllGather)(const void*, void*, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*ReduceScatter)(const void*, void*, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*Send)(const void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*Recv)(void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*GroupStart)() = nullptr;
    ncclResult_t (*GroupEnd)() = nullptr;
    const char* (*GetErrorString)(ncclResult_t) = nullptr;
    std::string err;
};

NcclApi* nccl_api()
{
    static NcclApi api;
    static std::once_flag once;
    std::call_once(once, [] {
        const char* names[] = { "libnccl.so.2", "libnccl.so" };
        for (const char* n : names) { api.lib = dlopen(n, RTLD_NOW | RTLD_GLOBAL); if (api.lib) break; }
        if (!api.lib) { api.err = std::string("cannot load libnccl.so.2: ") + (dlerror() ? dlerror() : ""); return; }
        auto sym = [&](const char* name) -> void* {
            void* p = dlsym(api.lib, name);
            if (!p && api.err.empty()) api.err = std::string("libnccl lacks ") + name;
            return p;
        };
        api.GetUniqueId = reinterpret_cast<decltype(api.GetUniqueId)>(sym("ncclGetUniqueId"));
        api.CommInitRank = reinterpret_cast<decltype(api.CommInitRank)>(sym("ncclCommInitRank"));
        api.CommDestroy = reinterpret_cast<decltype(api.CommDestroy)>(sym("ncclCommDestroy"));
        api.AllGather = reinterpret_cast<decltype(api.AllGather)>(sym("ncclAllGather"));
        api.ReduceScatter = reinterpret_cast<decltype(api.ReduceScatter)>(sym("ncclReduceScatter"));
        api.Send = reinterpret_cast<decltype(api.Send)>(sym("ncclSend"));
        api.Recv = reinterpret_cast<decltype(api.Recv)>(sym("ncclRecv"));
        api.GroupStart = reinterpret_cast<decltype(api.GroupStart)>(sym("ncclGroupStart"));
        api.GroupEnd = reinterpret_cast<decltype(api.GroupEnd)>(sym("ncclGroupEnd"));
        api.GetErrorString = reinterpret_cast<decltype(api.GetErrorString)>(sym("ncclGetErrorString"));
    });
    return &api;
}

int nccl_fail(fsgm_ctx* c, const char* what, ncclResult_t r)
{
    NcclApi* a = nccl_api();
    return fail(c, FSGM_ERR_NCCL, what, (a->GetErrorString && r != ncclSuccess) ? a->GetErrorString(r) : a->err.c_str());
}
#define FSGM_NCCL(ctx, expr) do { ncclResult_t r__ = (expr); if (r__ != ncclSuccess) return nccl_fail((ctx), #expr, r__); } while (0)

int need_comm(fsgm_ctx* c, NcclApi** api)
{
    if (!c) return FSGM_ERR_ARG;
    *api = nccl_api();
    if (!(*api)->err.empty()) return fail(c, FSGM_ERR_NCCL, "NCCL unavailable", (*api)->err.c_str());
    if (!c->nccl_comm) return fail(c, FSGM_ERR_NCCL, "no communicator: call fsgm_dist_init or fsgm_dist_adopt_comm first");
    return FSGM_OK;
}

// word[0] = sum over the n_vols slabs of the first byte (u8 exchange) / the first u16 of the reduced slab (u16 exchange)
__global__ void first_voxel_kernel(const uint8_t* slabs, int n_vols, size_t stride, const uint16_t* sp, uint32_t* out)
{
    uint32_t a = 0;
    if (sp) a = sp[0];
    else for (int k = 0; k < n_vols; ++k) a += slabs[(size_t)k * stride];
    *out = a;
}

struct P2PMsg { cudaIpcMemHandle_t h; int ok; int pad[15]; };
static_assert(sizeof(P2PMsg) == 128, "P2PMsg is exchanged as 128 bytes");

void p2p_drop(fsgm_ctx* c)
{
    for (int j = 0; j < 16; ++j) {
        if (c->p2p_peer[j] && c->p2p_peer[j] != c->p2p_local) cudaIpcCloseMemHandle(c->p2p_peer[j]);
        c->p2p_peer[j] = nullptr;
    }
    if (c->p2p_local) cudaFree(c->p2p_local);
    c->p2p_local = nullptr; c->p2p_bytes = 0;
}

// Collective: make sure every rank owns a receive buffer of >= bytes that all other ranks have mapped.  Returns FSGM_OK with
// *usable = false when peer mapping is not available anywhere (the decision is agreed between the ranks: all or none).
int p2p_ensure(fsgm_ctx* c, NcclApi* a, size_t bytes, bool* usable)
{
    *usable = false;
    if (c->p2p_state < 0 || c->no_p2p || c->world > 16) return FSGM_OK;
    if (bytes <= c->p2p_bytes) { *usable = true; return FSGM_OK; }
    ncclComm_t comm = static_cast<ncclComm_t>(c->nccl_comm);
    const int world = c->world, rank = c->rank;
    FSGM_CUDA(c, cudaStreamSynchronize(c->stream));
    p2p_drop(c);
    P2PMsg mine{};
    void* local = nullptr;
    mine.ok = cudaMalloc(&local, bytes) == cudaSuccess && cudaIpcGetMemHandle(&mine.h, local) == cudaSuccess;
    cudaGetLastError();
    P2PMsg* stage = nullptr;                                   // [1 + world] messages in device memory
    if (cudaMalloc(reinterpret_cast<void**>(&stage), sizeof(P2PMsg) * (world + 1)) != cudaSuccess) {
        cudaGetLastError(); if (local) cudaFree(local);
        return fail(c, FSGM_ERR_NOMEM, "cudaMalloc(peer handle staging)");
    }
    std::vector<P2PMsg> all(world);
    auto agree = [&](bool* everyone) -> int {                 // all-gather `mine`, AND the ok flags
        FSGM_CUDA(c, cudaMemcpyAsync(stage, &mine, sizeof mine, cudaMemcpyHostToDevice, c->stream));
        FSGM_NCCL(c, a->AllGather(stage, stage + 1, sizeof(P2PMsg), ncclUint8, comm, c->stream));
        FSGM_CUDA(c, cudaMemcpyAsync(all.data(), stage + 1, sizeof(P2PMsg) * world, cudaMemcpyDeviceToHost, c->stream));
        FSGM_CUDA(c, cudaStreamSynchronize(c->stream));
        *everyone = true;
        for (int j = 0; j < world; ++j) *everyone &= all[j].ok != 0;
        return FSGM_OK;
    };
    bool ok_all = false;
    int rc = agree(&ok_all);
    if (rc == FSGM_OK && ok_all) {
        c->p2p_local = local; c->p2p_bytes = bytes; local = nullptr;
        for (int j = 0; j < world && mine.ok; ++j) {
            if (j == rank) { c->p2p_peer[j] = c->p2p_local; continue; }
            void* ptr = nullptr;
            if (cudaIpcOpenMemHandle(&ptr, all[j].h, cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) { cudaGetLastError(); mine.ok = 0; }
            else c->p2p_peer[j] = ptr;
        }
        rc = agree(&ok_all);                                   // second round: did every rank map every buffer?
    }
    cudaFree(stage);
    if (local) cudaFree(local);
    if (rc != FSGM_OK) { p2p_drop(c); return rc; }
    if (!ok_all) { p2p_drop(c); c->p2p_state = -1; return FSGM_OK; }
    c->p2p_state = 1;
    *usable = true;
    return FSGM_OK;
}

// ---- striped ownership of the peer-store form: global stripe j (stripe_pixels pixels) lives on rank j % world as its local
// stripe j / world, in chunks of stripe_pixels + 1 pixels ------------------------------------------------------------------------
// offsetFromPosD0 of a rank's local pixels (zero for the extra / unused ones)
__global__ void stripe_gather_f64_kernel(const double* __restrict__ src, size_t N, unsigned stripe_pixels, int rank, int world,
                                         size_t local_pixels, double* __restrict__ dst)
{
    const size_t lp = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (lp >= local_pixels) return;
    const size_t chunk = (size_t)stripe_pixels + 1, ls = lp / chunk, off = lp - ls * chunk;
    const size_t g = (ls * world + rank) * stripe_pixels + off;
    dst[lp] = (off < stripe_pixels && g < N) ? src[g] : 0.0;
}

// all-gathered local outputs [world][local_pixels] -> the image
__global__ void stripe_scatter_u32_kernel(const uint32_t* __restrict__ a, const uint32_t* __restrict__ b, size_t N, unsigned stripe_pixels,
                                          int world, size_t local_pixels, uint32_t* __restrict__ oa, uint32_t* __restrict__ ob)
{
    const size_t g = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= N) return;
    const size_t j = g / stripe_pixels, owner = j % world, ls = j / world;
    const size_t src = owner * local_pixels + ls * ((size_t)stripe_pixels + 1) + (g - j * stripe_pixels);
    oa[g] = a[src]; ob[g] = b[src];
}

}  // namespace

void dist_release(fsgm_ctx* c)
{
    p2p_drop(c);
    c->p2p_state = 0;
    if (c->nccl_comm && c->nccl_owned) { NcclApi* a = nccl_api(); if (a->CommDestroy) a->CommDestroy(static_cast<ncclComm_t>(c->nccl_comm)); }
    c->nccl_comm = nullptr; c->nccl_owned = false; c->rank = 0; c->world = 1;
}

}  // namespace fsgm

using namespace fsgm;

extern "C" {

int fsgm_shard_range(int n, int rank, int world, int* first, int* count)
{
    if (n < 0 || world < 1 || rank < 0 || rank >= world || !first || !count) return FSGM_ERR_ARG;
    const int base = n / world, rem = n % world;
    *first = rank * base + std::min(rank, rem);
    *count = base + (rank < rem ? 1 : 0);
    return FSGM_OK;
}

// The enabled directions in the order they are dealt to the ranks (rank r takes positions r, r + world, ...).  Horizontal
// scanlines are the longest serial chains and the fewest warps, diagonals the cheapest: swapping neighbours in the reversed
// half (0 1 2 3 5 4 7 6) gives every rank one direction of each kind at world = 2 and a horizontal + a vertical one at
// world = 4, so nobody waits for a rank that drew both horizontal sweeps.
static int split_order(const fsgm_epi_opts& o, int* dirs)
{
    const int nd = enabled_dirs(o, dirs);
    for (int k = nd / 2; k + 1 < nd; k += 2) std::swap(dirs[k], dirs[k + 1]);
    return nd;
}

int fsgm_dirsplit_plan(int width, int height, int dMax, int paths, int P1, int P2, int rank, int world, fsgm_dirsplit_info* out)
{
    if (!out || width < 1 || height < 1 || dMax < 1 || (paths != 4 && paths != 8) || world < 1 || rank < 0 || rank >= world) return FSGM_ERR_ARG;
    const size_t N = (size_t)width * height;
    size_t slab = (N + world - 1) / world;
    slab += slab & 1;                                   // even: slab * dMax u16 values pack into whole 32-bit words
    out->slab_pixels = slab;
    out->padded_pixels = slab * world;
    fsgm_epi_opts o; fsgm_epi_opts_default(&o); o.paths = paths;
    int dirs[8];
    const int nd = split_order(o, dirs);
    out->n_dirs = 0;
    for (int k = rank; k < nd; k += world) out->dirs[out->n_dirs++] = dirs[k];
    const int k_max = (nd + world - 1) / world;         // most directions any rank owns
    out->exchange_u8 = (!sweep_needs_wrap(P1, P2, 24) && k_max * (24 + P2) <= 255 && dMax % 16 == 0) ? 1 : 0;
    const size_t lo = std::min(N, (size_t)rank * slab), hi = std::min(N, (size_t)(rank + 1) * slab);
    out->first_pixel = lo; out->n_pixels = hi - lo;
    return FSGM_OK;
}

int fsgm_dist_unique_id(void* id, size_t bytes)
{
    if (!id || bytes < sizeof(ncclUniqueId)) return FSGM_ERR_ARG;
    NcclApi* a = nccl_api();
    if (!a->err.empty()) return FSGM_ERR_NCCL;
    ncclUniqueId u;
    if (a->GetUniqueId(&u) != ncclSuccess) return FSGM_ERR_NCCL;
    std::memcpy(id, &u, sizeof u);
    return FSGM_OK;
}

int fsgm_dist_init(fsgm_ctx* c, const void* id, size_t bytes, int rank, int world)
{
    if (!c) return FSGM_ERR_ARG;
    if (!id || bytes < sizeof(ncclUniqueId) || world < 1 || rank < 0 || rank >= world) return fail(c, FSGM_ERR_ARG, "bad rank / world / id");
    NcclApi* a = nccl_api();
    if (!a->err.empty()) return fail(c, FSGM_ERR_NCCL, "NCCL unavailable", a->err.c_str());
    FSGM_CUDA(c, cudaSetDevice(c->device));
    dist_release(c);
    ncclUniqueId u;
    std::memcpy(&u, id, sizeof u);
    ncclComm_t comm = nullptr;
    FSGM_NCCL(c, a->CommInitRank(&comm, world, u, rank));
    c->nccl_comm = comm; c->nccl_owned = true; c->rank = rank; c->world = world;
    return FSGM_OK;
}

int fsgm_dist_adopt_comm(fsgm_ctx* c, void* nccl_comm, int rank, int world)
{
    if (!c) return FSGM_ERR_ARG;
    if (!nccl_comm || world < 1 || rank < 0 || rank >= world) return fail(c, FSGM_ERR_ARG, "bad rank / world / communicator");
    NcclApi* a = nccl_api();
    if (!a->err.empty()) return fail(c, FSGM_ERR_NCCL, "NCCL unavailable", a->err.c_str());
    dist_release(c);
    c->nccl_comm = nccl_comm; c->nccl_owned = false; c->rank = rank; c->world = world;
    return FSGM_OK;
}

int fsgm_dist_finalize(fsgm_ctx* c)
{
    if (!c) return FSGM_ERR_ARG;
    cudaSetDevice(c->device);
    cudaStreamSynchronize(c->stream);
    dist_release(c);
    return FSGM_OK;
}

int fsgm_dist_rank(const fsgm_ctx* c) { return c ? c->rank : -1; }
int fsgm_dist_world(const fsgm_ctx* c) { return c ? c->world : 0; }

int fsgm_dist_allgather_u32(fsgm_ctx* c, const uint32_t* d_send, size_t count, uint32_t* d_recv)
{
    NcclApi* a;
    FSGM_TRY(need_comm(c, &a));
    if (!d_send || !d_recv) return fail(c, FSGM_ERR_ARG, "null pointer");
    FSGM_CUDA(c, cudaSetDevice(c->device));
    StageScope ss(c, ST_EXCHANGE);
    FSGM_NCCL(c, a->AllGather(d_send, d_recv, count, ncclUint32, static_cast<ncclComm_t>(c->nccl_comm), c->stream));
    return FSGM_OK;
}

int fsgm_calc_cost_sgm_dirsplit_dev(fsgm_ctx* c, const uint8_t* d_I1, const uint8_t* d_I2, int W, int H, int D, double vMax,
                                    const double* d_Pd0, const double* d_dir, const double* d_O, int P1, int P2,
                                    const fsgm_epi_opts* opts, uint32_t* d_bestD, uint32_t* d_minC)
{
    FSGM_TRY(check_dims(c, 1, W, H, D));
    fsgm_epi_opts o;
    FSGM_TRY(check_opts(c, opts, &o));
    if (!d_I1 || !d_I2 || !d_Pd0 || !d_dir || !d_O || !d_bestD || !d_minC) return fail(c, FSGM_ERR_ARG, "null pointer");
    if (c->world <= 1 && !c->nccl_comm)                  // a single rank: the plain call
        return fsgm_calc_cost_sgm_dev(c, 1, d_I1, d_I2, W, H, D, vMax, d_Pd0, d_dir, d_O, P1, P2, &o, d_bestD, d_minC);
    NcclApi* a;
    FSGM_TRY(need_comm(c, &a));
    ncclComm_t comm = static_cast<ncclComm_t>(c->nccl_comm);
    FSGM_CUDA(c, cudaSetDevice(c->device));
    const int world = c->world, rank = c->rank;
    const size_t N = (size_t)W * H, V = N * D;
    // the plan (slabs, this rank's directions, exchange form); total_pass = 1 drops the reversed directions from the rank's list
    fsgm_dirsplit_info plan;
    if (fsgm_dirsplit_plan(W, H, D, o.paths, P1, P2, rank, world, &plan) != FSGM_OK) return fail(c, FSGM_ERR_ARG, "dirsplit plan");
    int my[8], nmy = 0;
    {
        int dirs[8];
        const int nd = split_order(o, dirs);
        for (int k = rank; k < nd; k += world) my[nmy++] = dirs[k];
    }
    const bool adaptive = o.adaptive_p2 != 0;
    const bool u8x = plan.exchange_u8 != 0;
    const size_t slab = plan.slab_pixels, npad = plan.padded_pixels, cnt = plan.n_pixels;
    const bool fused_cost = (D == 64 || D == 128 || D == 256);
    size_t need = 2 * align256(N * 4) + align256((size_t)D * 8) + (fused_cost ? 1 : 2) * align256(V) + 4 * align256(npad * 4) + 4096;
    need += (size_t)(3 + 2 * world) * align256(((size_t)8 * W + 1) * (((size_t)(H + 7) / 8 + world - 1) / world) * 8);   // striped peer-store form
    if (u8x) need += 2 * align256(npad * D) + (size_t)std::max(0, nmy - 1) * align256(V);
    else need += (size_t)std::max(1, nmy) * align256(V) + align256(npad * D * 2) + align256(slab * D * 2) + 2 * align256(N * 4);
    FSGM_TRY(arena_reserve(c, need));
    ArenaScope scope(c);
    // ---- cost volume (every rank, redundant) ------------------------------------------------------------------
    uint32_t *cen1, *cen2; uint8_t *raw = nullptr, *C; double* vz;
    FSGM_TRY(arena_get(c, (size_t)D, &vz));
    FSGM_TRY(arena_get(c, N, &cen1));
    FSGM_TRY(arena_get(c, N, &cen2));
    FSGM_TRY(arena_get(c, V, &C));
    FSGM_TRY(launch_census(c, 1, d_I1, W, H, cen1));
    FSGM_TRY(launch_census(c, 1, d_I2, W, H, cen2));
    FSGM_TRY(launch_vz_table(c, D, vMax, vz));
    bool fused = false;
    FSGM_TRY(launch_epi_cost_fused(c, 1, vMax, cen1, cen2, W, H, D, d_Pd0, d_dir, d_O, C, &fused));
    if (!fused) {
        FSGM_TRY(arena_get(c, V, &raw));
        FSGM_TRY(launch_epi_cost(c, 1, vz, cen1, cen2, W, H, D, vMax, d_Pd0, d_dir, d_O, raw, C));
    }
    uint32_t *slabB, *slabM, *allB, *allM, *firsts;
    FSGM_TRY(arena_get(c, slab, &slabB));
    FSGM_TRY(arena_get(c, slab, &slabM));
    FSGM_TRY(arena_get(c, npad, &allB));
    FSGM_TRY(arena_get(c, npad, &allM));
    FSGM_TRY(arena_get(c, (size_t)world + 1, &firsts));
    const uint16_t* next0 = (rank + 1 < world) ? reinterpret_cast<const uint16_t*>(firsts + 1 + rank + 1) : nullptr;   // low half of the word
    int all_dirs[8];
    const int nd_all = enabled_dirs(o, all_dirs);
    bool scatter = !adaptive && !sweep_needs_wrap(P1, P2, 24) && D % 16 == 0;
    // peer-store form: stripes of 8 image rows dealt round-robin to the ranks
    constexpr int STRIPE_ROWS = 8;
    const size_t sp = (size_t)STRIPE_ROWS * W, chunk = sp + 1;
    const size_t nls = ((size_t)(H + STRIPE_ROWS - 1) / STRIPE_ROWS + world - 1) / world, LP = nls * chunk;
    if (scatter) FSGM_TRY(p2p_ensure(c, a, (size_t)nd_all * LP * D, &scatter));
    if (scatter) {
        // ---- the sweeps write every L row into its stripe owner's buffer; no exchange step ------------------------------------------
        double* localO; uint32_t *locB, *locM, *gatB, *gatM;
        FSGM_TRY(arena_get(c, LP, &localO));
        FSGM_TRY(arena_get(c, LP, &locB));
        FSGM_TRY(arena_get(c, LP, &locM));
        FSGM_TRY(arena_get(c, LP * world, &gatB));
        FSGM_TRY(arena_get(c, LP * world, &gatM));
        int slots[8];
        for (int i = 0; i < nmy; ++i) slots[i] = rank + i * world;           // position of my i-th direction in the dealt order
        uint8_t* peers[16];
        for (int j = 0; j < world; ++j) peers[j] = static_cast<uint8_t*>(c->p2p_peer[j]);
        {
            // the pixel after the last one reads as zero (the reference reads past its buffer there): nobody writes that row
            const size_t j = (N - 1) / sp;
            if ((int)(j % world) == rank)
                for (int k = 0; k < nd_all; ++k)
                    FSGM_CUDA(c, cudaMemsetAsync(static_cast<uint8_t*>(c->p2p_local) + ((size_t)k * LP + (j / world) * chunk + (N - j * sp)) * D, 0, D, c->stream));
        }
        if (nmy) FSGM_TRY(launch_sweeps_scatter(c, C, W, H, D, P1, P2, my, slots, nmy, peers, world, sp, LP));
        {
            StageScope ss(c, ST_EXCHANGE);
            // every rank's stores have landed once every rank has passed this point (kernel completion makes them visible)
            FSGM_NCCL(c, a->AllGather(firsts, firsts + 1, 1, ncclUint32, comm, c->stream));
            stripe_gather_f64_kernel<<<(unsigned)((LP + 255) / 256), 256, 0, c->stream>>>(d_O, N, (unsigned)sp, rank, world, LP, localO);
            FSGM_LAUNCHED(c);
        }
        FSGM_TRY(launch_slab_wta(c, static_cast<const uint8_t*>(c->p2p_local), nd_all, LP * D, nullptr, LP, D, o.subpixel,
                                 o.vz_to_disp, localO, vMax, locB, locM));
        {
            StageScope ss(c, ST_EXCHANGE);
            FSGM_NCCL(c, a->AllGather(locB, gatB, LP, ncclUint32, comm, c->stream));
            FSGM_NCCL(c, a->AllGather(locM, gatM, LP, ncclUint32, comm, c->stream));
            stripe_scatter_u32_kernel<<<(unsigned)((N + 255) / 256), 256, 0, c->stream>>>(gatB, gatM, N, (unsigned)sp, world, LP, d_bestD, d_minC);
            FSGM_LAUNCHED(c);
        }
        return FSGM_OK;
    } else if (u8x) {
        // ---- this rank's directions summed into one u8 volume, exchanged slab-wise, summed inside the WTA kernel -------------------
        uint8_t *part, *recv, *L[8];
        FSGM_TRY(arena_get(c, npad * D, &part));
        FSGM_TRY(arena_get(c, npad * D, &recv));
        L[0] = part;
        for (int k = 1; k < nmy; ++k) FSGM_TRY(arena_get(c, V, &L[k]));
        if (npad > N) FSGM_CUDA(c, cudaMemsetAsync(part + V, 0, (npad - N) * D, c->stream));
        if (nmy == 0) FSGM_CUDA(c, cudaMemsetAsync(part, 0, V, c->stream));
        else {
            FSGM_TRY(launch_sweeps(c, 1, C, d_I1, W, H, D, P1, P2, adaptive ? 25 : 0, 24, my, nmy, L));
            for (int k = 1; k < nmy; ++k) FSGM_TRY(launch_add_u8(c, part, L[k], V));
        }
        {
            StageScope ss(c, ST_EXCHANGE);
            const size_t sb = slab * D;
            FSGM_NCCL(c, a->GroupStart());
            for (int j = 0; j < world; ++j) {
                if (j == rank) continue;
                FSGM_NCCL(c, a->Send(part + (size_t)j * sb, sb, ncclUint8, j, comm, c->stream));
                FSGM_NCCL(c, a->Recv(recv + (size_t)j * sb, sb, ncclUint8, j, comm, c->stream));
            }
            FSGM_NCCL(c, a->GroupEnd());
            FSGM_CUDA(c, cudaMemcpyAsync(recv + (size_t)rank * sb, part + (size_t)rank * sb, sb, cudaMemcpyDeviceToDevice, c->stream));
            first_voxel_kernel<<<1, 1, 0, c->stream>>>(recv, world, sb, nullptr, firsts);
            FSGM_LAUNCHED(c);
            FSGM_NCCL(c, a->AllGather(firsts, firsts + 1, 1, ncclUint32, comm, c->stream));
        }
        if (cnt) FSGM_TRY(launch_slab_wta(c, recv, world, slab * D, next0, cnt, D, o.subpixel, o.vz_to_disp, d_O + plan.first_pixel, vMax, slabB, slabM));
    } else {
        // ---- u16 partial volume, reduce-scatter over pixel slabs on the u16 pairs viewed as 32-bit words ------------------------------
        uint8_t* L[8]; uint16_t *part, *mine; uint32_t *tb, *tm;
        for (int k = 0; k < std::max(1, nmy); ++k) FSGM_TRY(arena_get(c, V, &L[k]));
        FSGM_TRY(arena_get(c, npad * D, &part));
        FSGM_TRY(arena_get(c, slab * D, &mine));
        FSGM_TRY(arena_get(c, N, &tb));
        FSGM_TRY(arena_get(c, N, &tm));
        if (npad > N) FSGM_CUDA(c, cudaMemsetAsync(part + V, 0, (npad - N) * D * 2, c->stream));
        if (nmy == 0) FSGM_CUDA(c, cudaMemsetAsync(part, 0, V * 2, c->stream));
        else {
            FSGM_TRY(launch_sweeps(c, 1, C, d_I1, W, H, D, P1, P2, adaptive ? 25 : 0, 24, my, nmy, L));
            FSGM_TRY(launch_epi_wta(c, 1, L, nmy, W, H, D, 0, 0, nullptr, 0.0, part, tb, tm));      // the summing pass of the WTA kernel
        }
        {
            StageScope ss(c, ST_EXCHANGE);
            FSGM_NCCL(c, a->ReduceScatter(part, mine, slab * D / 2, ncclUint32, ncclSum, comm, c->stream));
            first_voxel_kernel<<<1, 1, 0, c->stream>>>(nullptr, 0, 0, mine, firsts);
            FSGM_LAUNCHED(c);
            FSGM_NCCL(c, a->AllGather(firsts, firsts + 1, 1, ncclUint32, comm, c->stream));
        }
        if (cnt) FSGM_TRY(launch_sp_wta(c, mine, next0, cnt, D, o.subpixel, o.vz_to_disp, d_O + plan.first_pixel, vMax, slabB, slabM));
    }
    {
        StageScope ss(c, ST_EXCHANGE);
        FSGM_NCCL(c, a->AllGather(slabB, allB, slab, ncclUint32, comm, c->stream));
        FSGM_NCCL(c, a->AllGather(slabM, allM, slab, ncclUint32, comm, c->stream));
        FSGM_CUDA(c, cudaMemcpyAsync(d_bestD, allB, N * 4, cudaMemcpyDeviceToDevice, c->stream));
        FSGM_CUDA(c, cudaMemcpyAsync(d_minC, allM, N * 4, cudaMemcpyDeviceToDevice, c->stream));
    }
    return FSGM_OK;
}

}  // extern "C"
