/* dpr.h - C ABI of libdpr.so, the B200 (sm_100a) implementation of DiffPointRasterisation.jl's hot path.
 *
 * This is the drop-in boundary: the entry points below are what a Julia package extension binds with `ccall`
 * in place of the kernels of the reference's CUDA extension.  Reference locations are under /root/reference.
 *
 * Data layout (identical to the reference's canonical form, i.e. Julia's memory):
 *   points       Vector{SVector{N_in,T}}          = dense (N_in, P)          column-major
 *   rotation     Vector{SMatrix{N_out,N_in,T}}    = dense (N_out, N_in, B)   column-major
 *   translation  Vector{SVector{N_out,T}}         = dense (N_out, B)
 *   background, out_weight  (B)   - NULL means the FillArrays default Zeros / Ones (src/interface.jl:368-394,
 *   point_weight            (P)     ext/DiffPointRasterisationCUDAExt.jl:15-17)
 *   out, ds_dout            (g_1, ..., g_N_out, B) column-major: first grid axis contiguous, pose stride prod(g)
 *   d_points (N_in, P), d_rotation (N_out, N_in, B), d_translation (N_out, B), d_background (B),
 *   d_out_weight (B), d_point_weight (P)
 * All pointers of the device entry points are DEVICE pointers (CuPtr{T} in Julia) valid on the current device.
 * The caller owns every buffer, including the workspace.  Outputs are fully overwritten (the reference zero-fills
 * then accumulates: ext/DiffPointRasterisationCUDAExt.jl:272-276; forward overwrites with the background,
 * src/raster.jl:27), so callers need not pre-zero.  All work is enqueued on `stream`; nothing synchronises.
 *
 * Supported: T in {float, double}; any 1 <= N_in <= 4, 1 <= N_out <= 4 (the reference generates its kernels for any pair,
 * src/raster.jl:36-66, src/util.jl:26-27); tuned kernels for the 2-d outputs (2,2) and (3,2), generic ones elsewhere.
 * Return value: DPR_OK (0) or a negative dpr_status; never throws, never aborts.
 */
#ifndef DPR_H
#define DPR_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef void* dpr_stream_t; /* cudaStream_t / CUstream (CUDA.jl: stream().handle) */

enum dpr_status {
    DPR_OK = 0,
    DPR_ERR_BAD_DIMS = -1,      /* negative sizes, grid extent < 1, prod(grid)*B overflow            */
    DPR_ERR_UNSUPPORTED = -2,   /* N_in or N_out outside 1..4, or element size outside the supported set */
    DPR_ERR_NULL_POINTER = -3,  /* a required pointer is NULL                                          */
    DPR_ERR_WORKSPACE = -4,     /* workspace smaller than dpr_workspace_bytes()                        */
    DPR_ERR_CUDA = -5,          /* a CUDA runtime call failed; see dpr_last_error_message()            */
    DPR_ERR_NO_DEVICE = -6,     /* no CUDA device / not an sm_100 device: there is NO CPU fallback     */
    DPR_ERR_NCCL = -7,          /* a NCCL call failed (multi-GPU entry points only)                    */
    DPR_ERR_BAD_OPTION = -8
};

enum dpr_op { DPR_OP_FORWARD = 0, DPR_OP_PULLBACK = 1 };

/* Library/ABI version (major*10000 + minor*100 + patch). */
int dpr_version(void);
/* Static string for a status code; for DPR_ERR_CUDA the thread's last CUDA error text is in dpr_last_error_message(). */
const char* dpr_status_string(int status);
const char* dpr_last_error_message(void);

/* Scratch the caller must provide (device memory, 256-byte aligned) for one call; may return 0.
 * Replaces nothing in the reference (CUDA.jl allocates implicitly); the glue allocates a CuVector{UInt8}. */
size_t dpr_workspace_bytes(int op, int n_in, int n_out, const int64_t* grid, int64_t n_points, int64_t batch,
                           int sizeof_T);

/* Batched forward splat: replaces the canonical `raster!` method src/raster.jl:5-34 and its kernel
 * src/raster.jl:36-66 (incl. the background broadcast, :27) for CuArray arguments. */
int dpr_raster_forward_f32(int n_in, int n_out, const int64_t* grid, int64_t n_points, int64_t batch,
                           const float* points, const float* rotation, const float* translation,
                           const float* background, const float* out_weight, const float* point_weight,
                           float* out, void* workspace, size_t workspace_bytes, dpr_stream_t stream);
int dpr_raster_forward_f64(int n_in, int n_out, const int64_t* grid, int64_t n_points, int64_t batch,
                           const double* points, const double* rotation, const double* translation,
                           const double* background, const double* out_weight, const double* point_weight,
                           double* out, void* workspace, size_t workspace_bytes, dpr_stream_t stream);

/* Batched pullback: replaces ext/DiffPointRasterisationCUDAExt.jl:231-321 (driver: sum!, 5 x fill!, launch) and
 * its kernel :19-210; same gradients as the CPU method src/raster_pullback.jl:85-148.
 * d_background, d_out_weight and d_point_weight may be NULL (that gradient is then not computed). */
int dpr_raster_pullback_f32(int n_in, int n_out, const int64_t* grid, int64_t n_points, int64_t batch,
                            const float* ds_dout, const float* points, const float* rotation,
                            const float* translation, const float* out_weight, const float* point_weight,
                            float* d_points, float* d_rotation, float* d_translation, float* d_background,
                            float* d_out_weight, float* d_point_weight, void* workspace, size_t workspace_bytes,
                            dpr_stream_t stream);
int dpr_raster_pullback_f64(int n_in, int n_out, const int64_t* grid, int64_t n_points, int64_t batch,
                            const double* ds_dout, const double* points, const double* rotation,
                            const double* translation, const double* out_weight, const double* point_weight,
                            double* d_points, double* d_rotation, double* d_translation, double* d_background,
                            double* d_out_weight, double* d_point_weight, void* workspace, size_t workspace_bytes,
                            dpr_stream_t stream);

/* Host-buffer entry points: same semantics, but every pointer is a HOST pointer (the reference's CPU methods
 * src/raster.jl:5-34 and src/raster_pullback.jl:85-148 take host Arrays).  The library stages pose chunks through
 * device memory it owns (grown on demand, per device, released by dpr_host_release), overlapping H2D copies,
 * kernels and D2H copies on internal streams, and returns when the results are in the host buffers.
 * Pinned buffers (dpr_host_alloc) make the copies asynchronous. */
int dpr_raster_forward_host_f32(int n_in, int n_out, const int64_t* grid, int64_t n_points, int64_t batch,
                                const float* points, const float* rotation, const float* translation,
                                const float* background, const float* out_weight, const float* point_weight,
                                float* out);
int dpr_raster_forward_host_f64(int n_in, int n_out, const int64_t* grid, int64_t n_points, int64_t batch,
                                const double* points, const double* rotation, const double* translation,
                                const double* background, const double* out_weight, const double* point_weight,
                                double* out);
int dpr_raster_pullback_host_f32(int n_in, int n_out, const int64_t* grid, int64_t n_points, int64_t batch,
                                 const float* ds_dout, const float* points, const float* rotation,
                                 const float* translation, const float* out_weight, const float* point_weight,
                                 float* d_points, float* d_rotation, float* d_translation, float* d_background,
                                 float* d_out_weight, float* d_point_weight);
int dpr_raster_pullback_host_f64(int n_in, int n_out, const int64_t* grid, int64_t n_points, int64_t batch,
                                 const double* ds_dout, const double* points, const double* rotation,
                                 const double* translation, const double* out_weight, const double* point_weight,
                                 double* d_points, double* d_rotation, double* d_translation, double* d_background,
                                 double* d_out_weight, double* d_point_weight);
/* Non-blocking variants: the call returns at once with a ticket, the work proceeds on a helper thread (on the device that
 * was current at the call), dpr_host_wait(ticket) blocks until the results are in the host buffers, returns the status and
 * frees the ticket.  A forward and a pullback use separate staging arenas, so an INDEPENDENT forward and pullback
 * (ds_dout not computed from `out`) can overlap: the device -> host copies of `out` and the host -> device copies of
 * ds_dout then share the link's two directions.  The buffers must stay valid until the wait. */
typedef struct dpr_ticket* dpr_ticket_t;
int dpr_raster_forward_host_async_f32(int n_in, int n_out, const int64_t* grid, int64_t n_points, int64_t batch,
                                      const float* points, const float* rotation, const float* translation,
                                      const float* background, const float* out_weight, const float* point_weight,
                                      float* out, dpr_ticket_t* ticket);
int dpr_raster_forward_host_async_f64(int n_in, int n_out, const int64_t* grid, int64_t n_points, int64_t batch,
                                      const double* points, const double* rotation, const double* translation,
                                      const double* background, const double* out_weight, const double* point_weight,
                                      double* out, dpr_ticket_t* ticket);
int dpr_raster_pullback_host_async_f32(int n_in, int n_out, const int64_t* grid, int64_t n_points, int64_t batch,
                                       const float* ds_dout, const float* points, const float* rotation,
                                       const float* translation, const float* out_weight, const float* point_weight,
                                       float* d_points, float* d_rotation, float* d_translation, float* d_background,
                                       float* d_out_weight, float* d_point_weight, dpr_ticket_t* ticket);
int dpr_raster_pullback_host_async_f64(int n_in, int n_out, const int64_t* grid, int64_t n_points, int64_t batch,
                                       const double* ds_dout, const double* points, const double* rotation,
                                       const double* translation, const double* out_weight, const double* point_weight,
                                       double* d_points, double* d_rotation, double* d_translation, double* d_background,
                                       double* d_out_weight, double* d_point_weight, dpr_ticket_t* ticket);
int dpr_host_wait(dpr_ticket_t ticket);
int dpr_host_alloc(void** ptr, size_t bytes);  /* pinned host memory */
int dpr_host_free(void* ptr);
int dpr_host_release(void);                    /* frees the staging arena of the current device */

/* Multi-GPU (new; the reference is single-device): one process per GPU, poses sharded contiguously, points
 * replicated.  The forward needs no exchange; the pullback needs ONE sum over ranks of the pose-summed gradients
 * d_points (src/raster_pullback.jl:141) and d_point_weight (:146), which callers keep in one packed (N_in+1)*P buffer.
 * Rank 0 creates a 128-byte id (ncclUniqueId), ships it to the other ranks by any means, every rank calls
 * dpr_comm_init_rank, then dpr_comm_allreduce_sum_* after its local dpr_raster_pullback_* on the same stream.
 * NCCL is loaded with dlopen at first use; DPR_ERR_NCCL if it is not available.
 * Payloads up to 16 MB are summed by the library's own kernels over peer memory (symmetric buffers mapped with CUDA IPC
 * by dpr_comm_init_rank; flags written over NVLink; one-shot, or reduce-scatter + all-gather for large payloads on four or
 * more ranks; every element is added in rank order, so all ranks receive bit-identical sums); larger payloads, and boxes
 * without peer access, go through ncclAllReduce.  The calls are collective: every rank makes them in the same order. */
typedef struct dpr_comm* dpr_comm_t;
int dpr_comm_unique_id(void* id128);
int dpr_comm_init_rank(dpr_comm_t* comm, int n_ranks, int rank, const void* id128);
int dpr_comm_destroy(dpr_comm_t comm);
int dpr_comm_allreduce_sum_f32(dpr_comm_t comm, float* buf, int64_t count, dpr_stream_t stream);
int dpr_comm_allreduce_sum_f64(dpr_comm_t comm, double* buf, int64_t count, dpr_stream_t stream);
/* 1 when the communicator serves small payloads with its own peer-memory kernels, 0 when everything goes through NCCL */
int dpr_comm_uses_peer_memory(dpr_comm_t comm);

/* Introspection / tuning (benchmarks and tests). */
enum dpr_option {
    DPR_OPT_FORWARD_ALGO = 0,   /* 0 auto, 1 global-reduction kernel, 2 shared-memory tile kernel (2-d grids),
                                   3 tile-binned kernel (3-d grids: CTA per pose x tile, src/raster.jl:27,36-66)  */
    DPR_OPT_PULLBACK_ALGO = 1,  /* 0 auto, 1 generic gather kernel, 2 2-d kernel (paired loads), 3 2-d kernel (scalar loads),
                                   4 2-d kernel with TMA-staged pose images (Float32, image must fit shared memory),
                                   (5, 6: round-1 experiments, removed - rejected with DPR_ERR_BAD_OPTION),
                                   7 tile-binned kernel (3-d grids: CTA per pose x tile with its ds_dout tile on chip) */
    DPR_OPT_TILE_SMEM_BYTES = 2,/* shared-memory budget per CTA for tiles (0 = default)                       */
    DPR_OPT_POINT_SPLIT = 3,    /* forward: force the number of point splits per (pose, slab) (0 = auto)      */
    DPR_OPT_POSE_CHUNK = 4,     /* pullback: force poses per CTA (0 = auto)                                   */
    DPR_OPT_FORWARD_ACCUM = 5,  /* forward tile kernel: 0 auto (fixed-point where eligible), 1 float atomics  */
    DPR_OPT_POINT_SORT = 6,     /* 0 auto, 1 always sort the points first (pullback: spatially; forward: also by
                                   radius for the one-slab Float32 tile kernel), 2 never                      */
    DPR_OPT_TILE3D_TMA = 7,     /* tensor-map TMA (cp.async.bulk.tensor): 0 auto - ds_dout tiles of the 3-d tile pullback and the
                                   bank-skewed (padded) image copies of the 2-d staged pullback, when rows are 16-byte
                                   multiples; 1 never (cooperative tile loads / dense 1-d bulk copies)            */
    DPR_OPT_COMM_P2P = 9,       /* dpr_comm_*: 0 auto (payloads up to 16 MB: one-shot all-reduce over CUDA-IPC mapped peer memory,
                                   set up by dpr_comm_init_rank), 1 NCCL only.  Set before dpr_comm_init_rank.            */
    DPR_OPT_BINNING_CACHE = 8   /* 3-d tile path: 1 = the spatial pre-sort and the per-pose bins stay in the caller's workspace
                                   and are reused by the next call on the same (points, point_weight, rotation,
                                   translation) - e.g. the pullback after the forward (the rrule,
                                   ext/DiffPointRasterisationChainRulesCoreExt.jl:56-61).  Validated on the device by a
                                   128-bit hash of those inputs.  Contract: zero the first 256 bytes of a workspace when
                                   you allocate it and do not touch it between calls.  0 (default) = off.              */
};
int dpr_set_option(int option, int64_t value);
int64_t dpr_get_option(int option);
/* Number of kernels this library has launched in this process (for the bench's gpu_launches claim). */
int64_t dpr_kernel_launch_count(void);
/* Per-kernel timing for benchmarks: while enabled, every kernel launch is bracketed with CUDA events on its stream.
 * dpr_profile_enable(on) also clears the records; dpr_profile_get synchronises on record i and returns its name
 * (static string) and duration in milliseconds.  on = 2 is a checked mode for debugging: every launch is followed by a
 * stream synchronise and a kernel that faulted is named in dpr_last_error_message() and on stderr. */
int dpr_profile_enable(int on);
int dpr_profile_count(void);
int dpr_profile_get(int i, const char** name, float* ms);
/* Name of the kernel path the last forward / pullback call on this thread took (static string). */
const char* dpr_last_path(int op);

#ifdef __cplusplus
}
#endif
#endif /* DPR_H */
