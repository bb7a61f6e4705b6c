/* fsgm.h — C ABI of the B200-native fSGM hot path (census / Hamming cost volume, multi-direction
 * semi-global aggregation, winner-take-all + subpixel).
 *
 * Every gateway below is a drop-in for one MATLAB MEX gateway of the reference
 * (`void mexFunction(int nlhs, mxArray* plhs[], int nrhs, const mxArray* prhs[])`): same operands in
 * the same order, as flat pointers plus explicit sizes, caller-allocated outputs, int status.
 * Array layout is the reference's: row-major, x fastest (the MATLAB drivers permute before the call,
 * epipolar_sgm_of.m:33-43); two-plane double arrays are plane-major (plane 0 = X, plane 1 = Y).
 * INTEGRATION.md shows the MEX / ctypes stubs that bind these symbols.
 *
 * Pointers named d_* are DEVICE pointers; all others are HOST pointers.  There is no CPU fallback:
 * every entry point returns FSGM_ERR_CUDA if no sm_100 device is usable.
 */
#ifndef FSGM_H
#define FSGM_H

#include <stddef.h>
#include <stdint.h>

#if defined(__GNUC__)
#define FSGM_API __attribute__((visibility("default")))
#else
#define FSGM_API
#endif

#ifdef __cplusplus
extern "C" {
#endif

#define FSGM_OK             0
#define FSGM_ERR_ARG       -1   /* null pointer, non-positive size, bad option value           */
#define FSGM_ERR_DOMAIN    -2   /* parameter combination outside what the kernels implement    */
#define FSGM_ERR_CUDA      -3   /* CUDA runtime / launch failure (see fsgm_last_error)          */
#define FSGM_ERR_NOMEM     -4   /* scratch arena could not be grown                            */
#define FSGM_ERR_NCCL      -5   /* NCCL missing, no communicator on the context, or a collective failed           */

typedef struct fsgm_ctx fsgm_ctx;

/* ---- context: one per GPU; owns a stream and a scratch arena reused across calls ------------- */
FSGM_API int         fsgm_create(int device, fsgm_ctx** out);
FSGM_API void        fsgm_destroy(fsgm_ctx* ctx);
/* Launch on a caller-owned cudaStream_t instead of the context's own (NULL restores it). */
FSGM_API int         fsgm_set_stream(fsgm_ctx* ctx, void* cuda_stream);
FSGM_API int         fsgm_synchronize(fsgm_ctx* ctx);
FSGM_API const char* fsgm_last_error(const fsgm_ctx* ctx);
/* Tuning / A-B knobs (results never change).  key 1 = aggregation path of the epipolar variant: 0 auto (default),
 * -1 generic one-warp-per-scanline kernels only, 1..16 = thread-block-cluster size of the row-synchronous kernel.
 * key 2 = 1 disables the two-stream wave pipeline (front-end of wave i+1 under the cluster passes of wave i).
 * key 3 = pairs resident per SM in the calc_cost_sgm_ng kernel (1..3; 0 = chosen from the batch size).
 * key 4 = 1: the direction split never uses the peer-store form (NCCL exchange of the partial volumes instead).
 * key 5 = aggregation path of the pyramidal variant: 0 auto (one thread per path where it applies), -1 one-warp-per-scanline kernels only,
 *         -2 = auto with every shifted step forced through the label-by-label form (test knob), 1..16 = row-synchronous clusters of that size.
 * key 6 = 1: the pyramidal variant builds its cost volume with the direct kernel only (no separable box filter + fix-up list).
 * key 7 = 1: calc_pyd_cost_sgm_ng runs the cell-by-cell compatibility search at every step (no per-grid tables).
 * key 8 = rows per CTA of the fused epipolar cost kernel (0 = chosen from the grid size).
 * key 9 = cluster waves per staging chunk of the host-pointer batch gateways (default 3; a chunk is uploaded while the previous
 *         one computes, and the wave pipeline inside a chunk needs at least two waves to overlap anything). */
FSGM_API int         fsgm_tune(fsgm_ctx* ctx, int key, int value);
/* occupancy probe: resident clusters of `cluster_size` CTAs x `threads` threads with `smem_bytes` dynamic shared memory */
FSGM_API int         fsgm_debug_max_clusters(int cluster_size, size_t smem_bytes, int threads);
FSGM_API int         fsgm_abi_version(void);
/* number of kernels this context has launched so far (bench.py's gpu_launches) */
FSGM_API uint64_t    fsgm_launch_count(const fsgm_ctx* ctx);
FSGM_API size_t      fsgm_scratch_bytes(const fsgm_ctx* ctx);

/* Optional per-stage device timing (CUDA events on the launch stream around every stage's kernels).
 * fsgm_profile_read synchronises the recorded events and returns accumulated milliseconds / kernel launches. */
FSGM_API int         fsgm_profile_enable(fsgm_ctx* ctx, int on);
FSGM_API int         fsgm_profile_reset(fsgm_ctx* ctx);
FSGM_API int         fsgm_profile_read(fsgm_ctx* ctx, int stage, double* ms, uint64_t* launches);
FSGM_API const char* fsgm_stage_name(int stage);
FSGM_API int         fsgm_stage_count(void);

/* ---- options that are compile-time constants inside the reference ------------------------- */
typedef struct fsgm_epi_opts {
    int paths;        /* 4 = as shipped (calc_cost_sgm.cpp:104 enableDiagnalPath=false), 8 = diagonals on */
    int total_pass;   /* calc_cost_sgm.cpp:103, reference value 2; 1 keeps only the forward sweeps           */
    int subpixel;     /* calc_cost_sgm.cpp:560, reference value 1 (always on)                               */
    int adaptive_p2;  /* calc_cost_sgm.cpp:102, reference value 0; threshold 25 (:68-72)                    */
    int vz_to_disp;   /* calc_cost_sgm.cpp:592-594 (USE_VZIND), reference value 1                           */
    int fb_check;     /* calc_cost_sgm.cpp:589-590: the forward/backward check the reference ships commented
                         out; 0 = as shipped (conf, bestD2 stay zero), 1 = run it and fill conf / bestD2    */
    int fb_thr;       /* calc_cost_sgm.cpp:489 default argument thr = 2 (x256 label units)                  */
} fsgm_epi_opts;
FSGM_API void fsgm_epi_opts_default(fsgm_epi_opts* o);   /* {4, 2, 1, 0, 1, 0, 2} — the reference as shipped */
/* Wave size of the epipolar throughput path for this problem shape: the row-synchronous aggregation kernels give every pair one
 * thread-block cluster for a whole pass, so pairs go through in waves of this many (15 at KITTI width, 256 labels, 8 paths).
 * Batches that are a multiple of it have no partial wave (which takes the generic kernels).  0 = those kernels do not apply
 * (label count not 64/128/256, parameters outside the no-wrap domain, adaptive P2, image too wide for the on-chip path state). */
FSGM_API int         fsgm_epi_wave_pairs(fsgm_ctx* ctx, int width, int dMax, int P1, int P2, const fsgm_epi_opts* opts);

/* ---- gateway 1: calc_cost_sgm (calc_cost_sgm.cpp:539-598) ----------------------------------
 * [bestD, minC, conf, bestD2] = calc_cost_sgm(I1, I2, dMax, vMax, pixelPosD0, normlizeDirection,
 *                                             offsetFromPosD0, P1, P2)
 * I1,I2 u8[H][W]; pixelPosD0 f64[2][H][W] (1-based); normlizeDirection f64[2][H][W];
 * offsetFromPosD0 f64[H][W].  bestD u32[H][W] = pixel disparity x256, minC u32[H][W];
 * conf u8[H][W] and bestD2 u32[H][W] may be NULL, otherwise they are zero-filled exactly as the
 * reference leaves them (its forward/backward check is commented out, :589-590) — unless opts->fb_check
 * is set, which runs that check (on the label map, before the vz conversion) and fills both.
 * opts == NULL means the reference as shipped. */
FSGM_API int fsgm_calc_cost_sgm(fsgm_ctx* ctx, const uint8_t* I1, const uint8_t* I2, int width, int height,
                       int dMax, double vMax, const double* pixelPosD0, const double* normlizeDirection,
                       const double* offsetFromPosD0, int P1, int P2, const fsgm_epi_opts* opts,
                       uint32_t* bestD, uint32_t* minC, uint8_t* conf, uint32_t* bestD2);

/* Same call over n_pairs independent pairs; every array is pair-major (pair stride = its per-pair size). */
FSGM_API int fsgm_calc_cost_sgm_batch(fsgm_ctx* ctx, int n_pairs, const uint8_t* I1, const uint8_t* I2, int width, int height,
                             int dMax, double vMax, const double* pixelPosD0, const double* normlizeDirection,
                             const double* offsetFromPosD0, int P1, int P2, const fsgm_epi_opts* opts,
                             uint32_t* bestD, uint32_t* minC);

/* Enqueue-only form of the batch call: returns when all copies and kernels are queued.  The host buffers must stay valid,
 * and the outputs must not be read, until fsgm_synchronize(ctx).  Consecutive calls overlap (the host->device copies of one
 * call run under the kernels of the previous one); use pinned host memory for the copies to be truly asynchronous. */
FSGM_API int fsgm_calc_cost_sgm_batch_async(fsgm_ctx* ctx, int n_pairs, const uint8_t* I1, const uint8_t* I2, int width, int height,
                             int dMax, double vMax, const double* pixelPosD0, const double* normlizeDirection,
                             const double* offsetFromPosD0, int P1, int P2, const fsgm_epi_opts* opts,
                             uint32_t* bestD, uint32_t* minC);

/* Device-resident form: inputs and outputs already in HBM, asynchronous on the context's stream. */
FSGM_API int fsgm_calc_cost_sgm_dev(fsgm_ctx* ctx, int n_pairs, const uint8_t* d_I1, const uint8_t* d_I2, int width, int height,
                           int dMax, double vMax, const double* d_pixelPosD0, const double* d_normlizeDirection,
                           const double* d_offsetFromPosD0, int P1, int P2, const fsgm_epi_opts* opts,
                           uint32_t* d_bestD, uint32_t* d_minC);

/* ---- stage entry points (device pointers; the reference's internal seams) --------------------
 * census()            common.cpp:3-27
 * calc_cost()         calc_cost_sgm.cpp:319-412  (d_raw may be NULL; otherwise also receives the pre-box cost)
 * sgm() sweeps        calc_cost_sgm.cpp:114-257  one direction r = 0..7 in the order
 *                     L1(+x) L3(+y) L2(+x+y) L4(-x+y), then the same four reversed
 * sgm() WTA/subpixel  calc_cost_sgm.cpp:259-308 followed by convert_vzInd_to_disp (:414-426) */
FSGM_API int fsgm_census_dev(fsgm_ctx* ctx, int n_images, const uint8_t* d_img, int width, int height, uint32_t* d_census);
FSGM_API int fsgm_epi_cost_dev(fsgm_ctx* ctx, int n_pairs, const uint32_t* d_cen1, const uint32_t* d_cen2, int width, int height,
                      int dMax, double vMax, const double* d_pixelPosD0, const double* d_normlizeDirection,
                      const double* d_offsetFromPosD0, uint8_t* d_raw, uint8_t* d_C);
FSGM_API int fsgm_sweep_dev(fsgm_ctx* ctx, int n_pairs, const uint8_t* d_C, const uint8_t* d_I1, int width, int height, int dMax,
                   int P1, int P2, int adaptive_thr, int direction, uint8_t* d_L);
/* fsgm_epi_aggregate_dev / fsgm_epi_partial*_dev take a caller-owned cost volume: its largest byte is measured on the device
 * (one extra pass over d_C and a stream synchronisation).  Up to 24 — anything fsgm_epi_cost_dev produces — all kernel families
 * apply; above it the generic sweeps run in the exact or explicit mod-256 form the measured bound requires (the reference's
 * unsigned char arithmetic, common.h:4-8), never the forms derived for C <= 24.  A d_C that is not 16-byte aligned also takes
 * the generic sweeps (the row-synchronous kernels move cost rows with bulk copies). */
FSGM_API int fsgm_epi_aggregate_dev(fsgm_ctx* ctx, int n_pairs, const uint8_t* d_C, const uint8_t* d_I1, int width, int height,
                           int dMax, int P1, int P2, const fsgm_epi_opts* opts, uint16_t* d_Sp /* may be NULL */,
                           const double* d_offsetFromPosD0, double vMax, uint32_t* d_bestD, uint32_t* d_minC);

/* forward_backward_check (calc_cost_sgm.cpp:488-536) with calc_disp_from_first (:429-486) inlined, and
 * convert_vzInd_to_disp (:414-426) as a stage of its own.  d_bestD is the x256 label map BEFORE the vz conversion (that is
 * where the reference's call sits, :589-593), values < 512*256; n = dMax + 1 (:590, :593); use_vzind = 1 is the reference's
 * build (#define USE_VZIND, :4).  d_conf u8 [n_pairs][H][W] (1 = consistent), d_bestD2 u32 [n_pairs][H][W] (the map
 * projected into image 2, 512<<8 where nothing landed). */
FSGM_API int fsgm_forward_backward_check_dev(fsgm_ctx* ctx, int n_pairs, const uint32_t* d_bestD, int width, int height,
                         const double* d_pixelPosD0, const double* d_normlizeDirection, const double* d_offsetFromPosD0,
                         double vMax, int n, int thr, int use_vzind, uint8_t* d_conf, uint32_t* d_bestD2);
FSGM_API int fsgm_convert_vzind_to_disp_dev(fsgm_ctx* ctx, int n_pairs, uint32_t* d_bestD, int width, int height,
                         const double* d_offsetFromPosD0, double vMax, int n);

/* direction-split building blocks for ONE large pair (the R scan directions spread over GPUs, volumes reduced over
 * NVLink) — the stage seams of fsgm_calc_cost_sgm_dirsplit_dev below, for callers that issue the collectives themselves:
 *   partial : sweeps (calc_cost_sgm.cpp:114-257) for the listed directions, summed into u16 [H*W][dMax]
 *   wta_sp  : calc_cost_sgm.cpp:259-308 + :414-426 over n_pixels consecutive pixels of a reduced volume;
 *             d_next_label0 = Sp[.][0] of the pixel after the slab (the value the reference reads for label dMax-1),
 *             NULL = 0 (end of image) */
FSGM_API int fsgm_epi_partial_dev(fsgm_ctx* ctx, const uint8_t* d_C, const uint8_t* d_I1, int width, int height, int dMax,
                         int P1, int P2, int adaptive_p2, const int* directions, int n_dirs, uint16_t* d_Sp_partial);
/* 8-bit form for >= 4 GPUs (each rank owns at most two directions: their sum fits a byte when n_dirs*(24+P2) <= 255):
 *   partial_u8 : the listed directions summed into u8 [H*W][dMax]; ranks exchange pixel slabs with an all-to-all
 *   wta_slabs  : WTA over n_pixels pixels from n_slabs received u8 slabs laid out back to back */
FSGM_API int fsgm_epi_partial_u8_dev(fsgm_ctx* ctx, const uint8_t* d_C, const uint8_t* d_I1, int width, int height, int dMax,
                         int P1, int P2, int adaptive_p2, const int* directions, int n_dirs, uint8_t* d_partial);
FSGM_API int fsgm_epi_wta_slabs_dev(fsgm_ctx* ctx, const uint8_t* d_slabs, int n_slabs, const uint16_t* d_next_label0,
                         size_t n_pixels, int dMax, int subpixel, int vz_to_disp, const double* d_offsetFromPosD0, double vMax,
                         uint32_t* d_bestD, uint32_t* d_minC);
FSGM_API int fsgm_epi_wta_sp_dev(fsgm_ctx* ctx, const uint16_t* d_Sp, const uint16_t* d_next_label0, size_t n_pixels, int dMax,
                         int subpixel, int vz_to_disp, const double* d_offsetFromPosD0, double vMax,
                         uint32_t* d_bestD, uint32_t* d_minC);

/* ---- multi-GPU (one fsgm_ctx per GPU / rank; NCCL over NVLink) ---------------------------------------------------
 * The library binds NCCL at run time (dlopen of libnccl.so.2) and owns, or is lent, one communicator per context.  All failures on
 * this side return FSGM_ERR_NCCL (fsgm_last_error has NCCL's text).
 *   fsgm_dist_unique_id   rank 0 creates the rendezvous id (128 bytes, ncclUniqueId) and hands it to the other ranks by any channel
 *   fsgm_dist_init        ncclCommInitRank on the context's device; collective over all `world` ranks
 *   fsgm_dist_adopt_comm  use a caller-owned ncclComm_t instead (not destroyed by the library)
 *   fsgm_dist_finalize    drop the communicator (fsgm_destroy does it too)
 * Two partitionings exist for this path:
 *   batch of independent pairs : rank r takes the block fsgm_shard_range() gives it and calls the ordinary gateways — no data-path
 *                                collective; fsgm_dist_allgather_u32 is there for callers that want the sharded outputs everywhere
 *   one large pair             : fsgm_calc_cost_sgm_dirsplit_dev — every rank passes the SAME pair; the scan directions of sgm()
 *                                (calc_cost_sgm.cpp:114-257) are split over the ranks, the per-direction volumes are reduced over
 *                                NVLink per pixel slab — by the sweep kernel itself, which stores every pixel's row into the
 *                                peer-mapped memory of the slab's owner while it computes (fallback without peer access: u8
 *                                slabs through one grouped send/recv exchange when every rank's directions fit a byte
 *                                together, else ncclReduceScatter on u16 pairs typed ncclUint32) — every rank runs WTA on
 *                                its slab and the outputs are all-gathered: d_bestD / d_minC are complete and
 *                                identical on every rank, and bit-identical to fsgm_calc_cost_sgm_dev on one GPU. */
#define FSGM_DIST_ID_BYTES 128
typedef struct fsgm_dirsplit_info {
    size_t slab_pixels;      /* pixels per rank after the reduction (even; the last slabs may reach past the image)          */
    size_t padded_pixels;    /* slab_pixels * world                                                                         */
    size_t first_pixel;      /* this rank's slab: pixels [first_pixel, first_pixel + n_pixels) of the row-major image       */
    size_t n_pixels;
    int    dirs[8];          /* this rank's scan directions (order of fsgm_sweep_dev), n_dirs of them                       */
    int    n_dirs;
    int    exchange_u8;      /* 1: u8 slabs, grouped send/recv; 0: u16 reduce-scatter                                       */
} fsgm_dirsplit_info;
FSGM_API int fsgm_shard_range(int n, int rank, int world, int* first, int* count);
FSGM_API int fsgm_dirsplit_plan(int width, int height, int dMax, int paths, int P1, int P2, int rank, int world, fsgm_dirsplit_info* out);
FSGM_API int fsgm_dist_unique_id(void* id, size_t bytes);
FSGM_API int fsgm_dist_init(fsgm_ctx* ctx, const void* id, size_t bytes, int rank, int world);
FSGM_API int fsgm_dist_adopt_comm(fsgm_ctx* ctx, void* nccl_comm, int rank, int world);
FSGM_API int fsgm_dist_finalize(fsgm_ctx* ctx);
FSGM_API int fsgm_dist_rank(const fsgm_ctx* ctx);
FSGM_API int fsgm_dist_world(const fsgm_ctx* ctx);
FSGM_API int fsgm_dist_allgather_u32(fsgm_ctx* ctx, const uint32_t* d_send, size_t count, uint32_t* d_recv /* [world][count] */);
FSGM_API int fsgm_calc_cost_sgm_dirsplit_dev(fsgm_ctx* ctx, const uint8_t* d_I1, const uint8_t* d_I2, int width, int height,
                           int dMax, double vMax, const double* d_pixelPosD0, const double* d_normlizeDirection,
                           const double* d_offsetFromPosD0, int P1, int P2, const fsgm_epi_opts* opts,
                           uint32_t* d_bestD, uint32_t* d_minC);

/* ---- gateway 2: calc_pyd_cost_sgm (calc_pyd_cost_sgm.cpp:439-510) ------------------------------
 * [bestD, minC, mvSub] = calc_pyd_cost_sgm(I1, I2, preMv, halfSearchWinSizeX, halfSearchWinSizeY, aggHalfWinSize,
 *                                          subPixelRefine, P1, P2, enableDiagnalPath, totalPass, adpativeP2)
 * preMv f64[2][mvHeight][mvWidth] with its own stride mvWidth >= width (:388, :493-494).
 * bestD u32[H][W] = raw label sx*(2ry+1)+sy; minC u32[H][W]; mvSub f64[2][H][W] (zeros unless subPixelRefine).
 * Labels are ordered offx-outer (:392-393).  totalPass is the reference's loop bound (0..16 accepted). */
FSGM_API int fsgm_calc_pyd_cost_sgm(fsgm_ctx* ctx, const uint8_t* I1, const uint8_t* I2, int width, int height,
                           const double* preMv, int mvWidth, int mvHeight,
                           int halfSearchWinSizeX, int halfSearchWinSizeY, int aggHalfWinSize, int subPixelRefine,
                           int P1, int P2, int enableDiagnalPath, int totalPass, int adpativeP2,
                           uint32_t* bestD, uint32_t* minC, double* mvSub);
FSGM_API int fsgm_calc_pyd_cost_sgm_dev(fsgm_ctx* ctx, int n_pairs, const uint8_t* d_I1, const uint8_t* d_I2, int width, int height,
                           const double* d_preMv, int mvWidth, int mvHeight,
                           int halfSearchWinSizeX, int halfSearchWinSizeY, int aggHalfWinSize, int subPixelRefine,
                           int P1, int P2, int enableDiagnalPath, int totalPass, int adpativeP2,
                           uint32_t* d_bestD, uint32_t* d_minC, double* d_mvSub);
/* stage seams: calc_cost (:374-437), one sweep of sgm2d (:142-296, step :34-89), sweeps + WTA + subpixel (:114-372) */
FSGM_API int fsgm_pyd_cost_dev(fsgm_ctx* ctx, int n_pairs, const uint32_t* d_cen1, const uint32_t* d_cen2, int width, int height,
                           const double* d_preMv, int mvWidth, int mvHeight, int aggHalfWinSize,
                           int halfSearchWinSizeX, int halfSearchWinSizeY, uint8_t* d_C);
FSGM_API int fsgm_pyd_sweep_dev(fsgm_ctx* ctx, int n_pairs, const uint8_t* d_C, const uint8_t* d_I1, const double* d_preMv,
                           int mvWidth, int mvHeight, int width, int height, int halfSearchWinSizeX, int halfSearchWinSizeY,
                           int P1, int P2, int adpativeP2, int direction, uint8_t* d_L);
FSGM_API int fsgm_pyd_aggregate_dev(fsgm_ctx* ctx, int n_pairs, const uint8_t* d_C, const uint8_t* d_I1, const double* d_preMv,
                           int mvWidth, int mvHeight, int width, int height, int halfSearchWinSizeX, int halfSearchWinSizeY,
                           int subPixelRefine, int P1, int P2, int enableDiagnalPath, int totalPass, int adpativeP2,
                           uint16_t* d_Sp /* may be NULL */, uint32_t* d_bestD, uint32_t* d_minC, double* d_mvSub);

/* ---- dense epipolar prologue / epilogue: the per-pixel part of the MATLAB driver around gateway 1 --------------------------
 * epipolar_geometry.m:104-119 (with rotation_motion.m:7-35) and epipolar_sgm_of.m:46-51.  F (fundamental matrix), H (rotation
 * homography K*R/K), the epipole in image 2 (1-based pixel coordinates, as MATLAB's epi(1:2)) and the expansion/contraction
 * flag `direction` (epipolar_geometry.m:94-98) are HOST arrays: F and H row-major 3x3 per pair ([n_pairs][9]), epipole
 * [n_pairs][2], direction [n_pairs] (NULL = all 0).  Feature matching, LMedS and the SVDs that produce them stay on the host.
 *   geometry : -> pixelPosD0 f64 [n][2][H][W] (1-based), normlizeDirection f64 [n][2][H][W], offsetFromPosD0 f64 [n][H][W],
 *              Rflow f64 [n][2][H][W] (may be NULL)
 *   flow     : flow = (bestD / 256) * normlizeDirection + Rflow, f64 [n][2][H][W]
 *   sgm_of   : geometry -> calc_cost_sgm -> flow in one call; the 40-byte-per-pixel fp64 operands of gateway 1 never cross
 *              PCIe.  d_work = fsgm_epipolar_sgm_of_work_bytes() bytes of device scratch.
 * MATLAB evaluates F*p and H*p through BLAS (summation order / FMA unspecified): the order is fixed here as
 * (m1*x + m2*y) + m3, separately rounded, and restated in oracle/geometry_oracle.py. */
FSGM_API int fsgm_epipolar_geometry_dev(fsgm_ctx* ctx, int n_pairs, const double* F, const double* H, const double* epipole,
                           const int* direction, int width, int height, double* d_pixelPosD0, double* d_normlizeDirection,
                           double* d_offsetFromPosD0, double* d_Rflow);
FSGM_API int fsgm_epipolar_flow_dev(fsgm_ctx* ctx, int n_pairs, const uint32_t* d_bestD, const double* d_normlizeDirection,
                           const double* d_Rflow, int width, int height, double* d_flow);
FSGM_API size_t fsgm_epipolar_sgm_of_work_bytes(int n_pairs, int width, int height);
FSGM_API int fsgm_epipolar_sgm_of_dev(fsgm_ctx* ctx, int n_pairs, const uint8_t* d_I0, const uint8_t* d_I1, int width, int height,
                           const double* F, const double* H, const double* epipole, const int* direction,
                           int dMax, double vMax, int P1, int P2, const fsgm_epi_opts* opts, void* d_work,
                           double* d_flow, uint32_t* d_minC);
/* host images in, host flow out; enqueue-only (fsgm_synchronize before reading), consecutive calls overlap */
FSGM_API int fsgm_epipolar_sgm_of_batch_async(fsgm_ctx* ctx, int n_pairs, const uint8_t* I0, const uint8_t* I1, int width, int height,
                           const double* F, const double* H, const double* epipole, const int* direction,
                           int dMax, double vMax, int P1, int P2, const fsgm_epi_opts* opts, double* flow, uint32_t* minC);
/* float forms: flow as float [n][H][W][2], (u, v) interleaved per pixel — the reference's own return type, CV_32FC2
 * (proj/include/epi_sgm.h:6-12) — each component the fp64 value rounded to nearest.  8 instead of 16 bytes per pixel come
 * back over PCIe; minC may be NULL in the host form (then 8 B/px down in total). */
FSGM_API int fsgm_epipolar_sgm_of_f32_dev(fsgm_ctx* ctx, int n_pairs, const uint8_t* d_I0, const uint8_t* d_I1, int width, int height,
                           const double* F, const double* H, const double* epipole, const int* direction,
                           int dMax, double vMax, int P1, int P2, const fsgm_epi_opts* opts, void* d_work,
                           float* d_flow, uint32_t* d_minC);
FSGM_API int fsgm_epipolar_sgm_of_f32_batch_async(fsgm_ctx* ctx, int n_pairs, const uint8_t* I0, const uint8_t* I1, int width, int height,
                           const double* F, const double* H, const double* epipole, const int* direction,
                           int dMax, double vMax, int P1, int P2, const fsgm_epi_opts* opts, float* flow, uint32_t* minC);
FSGM_API int fsgm_epipolar_sgm_of(fsgm_ctx* ctx, const uint8_t* I0, const uint8_t* I1, int width, int height,
                           const double* F, const double* H, const double* epipole, int direction,
                           int dMax, double vMax, int P1, int P2, const fsgm_epi_opts* opts, double* flow, uint32_t* minC);

/* ---- pyramid driver: the MATLAB loop around gateway 2 (pyramidal_sgm.m:1-77), every level resident on the device ----------
 * [mvCurLevel, mvPyd, minC] = pyramidal_sgm(I0, I1, numPyd)
 * Builds both image pyramids with impyramid(.,'reduce') (:28-29), then from the coarsest level down (:36-75): calc_pyd_cost_sgm
 * with the previous level's flow as prior (zeros at the top, :33), subpixel refinement only at the finest level (:48),
 * label -> motion vector (:57-64), and 2*imresize(mv, 2, 'nearest') as the next prior (:72; its size 2*ceil(n/2) >= n is the
 * reason preMv has its own stride).  impyramid belongs to MATLAB's Image Processing Toolbox, not to the reference tree: its
 * published algorithm is restated in csrc/pyramid.cu (5-tap [1 4 6 4 1]/16, ceil(n/2), symmetric border, uint8 rounding after
 * each of the two passes).  The struct carries the constants the script hard-codes (:14-22). */
typedef struct fsgm_pyd_opts {
    int numPyd;                  /* pyramidal_sgm.m:12  5  (1..16) */
    int P1, P2;                  /* :14-15  6, 32 */
    int aggHalfWinSize;          /* :16  2 */
    int verSearchHalfWinSize;    /* :17  5 */
    int horSearchHalfWinSize;    /* :18  5 */
    int enableDiagonal;          /* :19  1 */
    int totalPass;               /* :20  2 */
    int adaptiveP2;              /* :21  0 */
} fsgm_pyd_opts;
FSGM_API void fsgm_pyd_opts_default(fsgm_pyd_opts* o);
/* level sizes, finest first: widths[0] = width, widths[l] = ceil(widths[l-1] / 2) */
FSGM_API int fsgm_pyramid_dims(int width, int height, int numPyd, int* widths, int* heights);
/* impyramid(img, 'reduce') on n_images u8 images: d_out is u8 [n_images][ceil(H/2)][ceil(W/2)] */
FSGM_API int fsgm_impyramid_reduce_dev(fsgm_ctx* ctx, int n_images, const uint8_t* d_img, int width, int height, uint8_t* d_out);
/* mv f64 [n_pairs][2][H][W] (plane 0 = x, 1 = y) and minC u32 [n_pairs][H][W] of the finest level; mvPyd (optional, may be NULL)
 * receives every level's flow back to back, finest first, level l as f64 [n_pairs][2][Hl][Wl]. */
FSGM_API int fsgm_pyramidal_sgm_dev(fsgm_ctx* ctx, int n_pairs, const uint8_t* d_I0, const uint8_t* d_I1, int width, int height,
                           const fsgm_pyd_opts* opts, double* d_mv, uint32_t* d_minC, double* d_mvPyd);
FSGM_API int fsgm_pyramidal_sgm(fsgm_ctx* ctx, const uint8_t* I0, const uint8_t* I1, int width, int height,
                           const fsgm_pyd_opts* opts, double* mv, uint32_t* minC, double* mvPyd);

/* ---- gateway 3: calc_cost_sgm_ng (calc_cost_sgm_ng.cpp:484-527) ----------------------------------
 * [minC, flow] = calc_cost_sgm_ng(I1, I2, preMv, halfSearchWinSize, aggSize, subPixelRefine, P1, P2)
 * The reference reads and ignores preMv, halfSearchWinSize, aggSize and subPixelRefine (:497-503); they are accepted
 * here (preMv may be NULL) and ignored too.  minC u32[H][W]; flow f64[2][H][W] (integer-valued).
 * The reference draws its random hints from libc rand() (:148-149), 8 values per pixel in raster order, from whatever
 * state the process is in.  Here that hidden state is an input: either a seed (the stream glibc's srand(seed); rand()
 * produces — the pinned oracle uses srand(1)) or an explicit stream of 8*W*H rand() values (e.g. MSVC's). */
typedef struct fsgm_ng_opts {
    unsigned       seed;          /* used when rand_stream == NULL */
    const int32_t* rand_stream;   /* HOST pointer, 8*width*height values, or NULL */
} fsgm_ng_opts;
FSGM_API void fsgm_ng_opts_default(fsgm_ng_opts* o);            /* {1, NULL} */
FSGM_API int  fsgm_glibc_rand_fill(unsigned seed, size_t count, int32_t* out);   /* srand(seed); out[i] = rand() */
FSGM_API int fsgm_calc_cost_sgm_ng(fsgm_ctx* ctx, const uint8_t* I1, const uint8_t* I2, int width, int height,
                          const double* preMv, double halfSearchWinSize, double aggSize, int subPixelRefine,
                          int P1, int P2, const fsgm_ng_opts* opts, uint32_t* minC, double* flow);
/* device form over n_pairs independent pairs (one CTA per pair: the path is serial inside a pair); seeds[n_pairs] is a
 * HOST array (NULL = all 1); d_rand_stream (device, 8*W*H per pair) overrides the seeds; d_Sp (u32 [n][N][108]) and
 * d_Centries (int32 [n][N][108][3] = mvx,mvy,cost) are optional stage outputs */
FSGM_API int fsgm_calc_cost_sgm_ng_dev(fsgm_ctx* ctx, int n_pairs, const uint8_t* d_I1, const uint8_t* d_I2, int width, int height,
                          int P1, int P2, const unsigned* seeds, const int32_t* d_rand_stream,
                          uint32_t* d_minC, double* d_flow, uint32_t* d_Sp, int32_t* d_Centries);

/* ---- gateway 4: calc_pyd_cost_sgm_ng (calc_pyd_cost_sgm_ng.cpp:448-523) -------------------------
 * [minC, flow] = calc_pyd_cost_sgm_ng(I1, I2, preMv, halfSearchWinSize, aggSize, subPixelRefine, P1, P2)
 * preMv f64[2][mvHeight][mvWidth] (hints are clamped to its size, :392-393); candidates = 9*(2r+1)^2 with
 * r = halfSearchWinSize (0..5 supported: up to 9 x 11 x 11 = 1089 candidates); aggregation radius = aggSize/2 (:490). */
FSGM_API int fsgm_calc_pyd_cost_sgm_ng(fsgm_ctx* ctx, const uint8_t* I1, const uint8_t* I2, int width, int height,
                          const double* preMv, int mvWidth, int mvHeight, int halfSearchWinSize, int aggSize,
                          int subPixelRefine, int P1, int P2, uint32_t* minC, double* flow);
/* optional stage outputs: d_Sp u32 [n][N][D]; d_cost u8 [n][N][D]; d_XY int32 [n][N][2][9][2r+1] (entry mv factored as
 * X[h][ox], Y[h][oy]) */
FSGM_API int fsgm_calc_pyd_cost_sgm_ng_dev(fsgm_ctx* ctx, int n_pairs, const uint8_t* d_I1, const uint8_t* d_I2, int width, int height,
                          const double* d_preMv, int mvWidth, int mvHeight, int halfSearchWinSize, int aggSize,
                          int subPixelRefine, int P1, int P2, uint32_t* d_minC, double* d_flow,
                          uint32_t* d_Sp, uint8_t* d_cost, int32_t* d_XY);

#ifdef __cplusplus
}
#endif
#endif /* FSGM_H */
