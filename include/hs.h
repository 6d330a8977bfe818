/*
 * hs.h - C ABI of the B200-native Horn-Schunck solver (libhs_b200.so).
 *
 * This is the drop-in boundary for the one hot path of liuyang9609/Cpp-Optical-Flow:
 *     class hornSchunck            HornSchunckOF/hornSchunck.cpp:8-76
 *     call site                    HornSchunckOF/main.cpp:97-98
 * The reference has no FFI layer of its own (the class is #included textually, main.cpp:8);
 * the entry points below are what a binding for that class has to reach, one per public
 * member plus the device-resident / row-slab helpers the multi-GPU host code needs.
 * cpp-optical-flow_b200/adapter/hornSchunck.cpp is the header-only cv::Mat adapter with the
 * reference's class surface; cpp-optical-flow_b200/hs_ctypes.py binds the same symbols from
 * Python (INTEGRATION.md shows both).
 *
 * Conventions
 *   - plain C, no CUDA/torch/OpenCV types; every pointer is a raw host or device address
 *   - every function returns an hs_status (HS_OK == 0) and never throws; the text of the last
 *     failure is available from hs_last_error(ctx) (or hs_last_error(NULL) for hs_create)
 *   - a context is bound to one CUDA device (or to the devices listed in hs_config.device_ids) and
 *     must not be used from two threads at once;
 *     distinct contexts are independent; there is no global mutable state
 *   - there is NO CPU fallback: without a usable CUDA device hs_create fails with HS_ERR_CUDA
 *   - images are row-major with a byte stride ("step" of cv::Mat); strides may exceed the row
 *
 * Semantics (identical to the reference for the same windowSize / maxIterations / alpha):
 *   gradX, gradY = unnormalised 3x3 Sobel of the PREVIOUS frame, BORDER_REFLECT_101  (:27-28)
 *   gradT        = next - prev                                                     (:39)
 *   per sweep    ubar = w x w box mean of u incl. centre, zero padding, anchor w-w/2-1 (:53-54,60)
 *                c = (gradX*ubar + gradY*vbar + gradT) / (alpha^2 + gradX^2 + gradY^2)  (:63-68)
 *                u = ubar - gradX*c,  v = vbar - gradY*c   (Jacobi, both from old averages) (:69-73)
 *   u = v = 0 initially (:49-50), exactly maxIterations sweeps, no convergence test (:56).
 * Arithmetic is fp32 on the device (gradients are exact integers); outputs are widened to fp64
 * on request because the reference's consumers read u.at<double> (plotFlow.cpp:72-75).
 */
#ifndef HS_B200_H
#define HS_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define HS_VERSION 200 /* 0.2.0 */

typedef enum hs_status {
    HS_OK = 0,
    HS_ERR_INVALID_ARG = 1,
    HS_ERR_CUDA = 2,
    HS_ERR_OOM = 3,
    HS_ERR_UNSUPPORTED = 4,
    HS_ERR_STATE = 5, /* call order violated, e.g. iterate before prepare */
    HS_ERR_NCCL = 6   /* HS_EXCHANGE_NCCL: libnccl missing or an NCCL call failed */
} hs_status;

typedef enum hs_dtype { HS_F32 = 0, HS_F64 = 1 } hs_dtype;

/* Arithmetic of the solve.  HS_PREC_F32 (default) is the fast path: exact integer gradients, fp32
 * sweeps in the fused kernel, within 1e-4 px of the reference's fp64 result.  HS_PREC_F64 repeats the
 * reference's own fp64 arithmetic operation by operation (convertTo(CV_64FC1) :23-24, every product
 * and sum of :60-73 rounded separately, no contraction): BIT-IDENTICAL to the fp64 oracle for w <= 7.
 * It is the A/B diagnostic for fp32 rounding and the path for frames that are not 8-bit. */
typedef enum hs_precision { HS_PREC_F32 = 0, HS_PREC_F64 = 1 } hs_precision;

/* Element type of the prev/next frames handed to hs_solve / hs_upload / hs_gradients (cv::Mat depths;
 * the reference accepts any depth, hornSchunck.cpp:23-24).  Anything but HS_FRAME_U8 needs HS_PREC_F64. */
typedef enum hs_frame_dtype {
    HS_FRAME_U8 = 0, HS_FRAME_S8 = 1, HS_FRAME_U16 = 2, HS_FRAME_S16 = 3, HS_FRAME_S32 = 4, HS_FRAME_F32 = 5, HS_FRAME_F64 = 6
} hs_frame_dtype;

/* How one context spreads over several GPUs (hs_config.num_devices > 1, or slab_world > 1). */
typedef enum hs_decomposition {
    HS_DECOMP_BATCH = 0,    /* independent frame pairs: pair i runs on device i mod N, no communication        */
    HS_DECOMP_ROW_SLAB = 1  /* ONE image cut into N bands of rows; k-row halos cross the seams every k sweeps  */
} hs_decomposition;

/* Row slabs: how the halo rows cross a seam. */
typedef enum hs_exchange {
    HS_EXCHANGE_PEER = 0, /* default: the fused kernel stores seam rows straight into the neighbour's halo rows
                             (peer memory over NVLink) and signals per-tile flags; no host work between sweeps */
    HS_EXCHANGE_NCCL = 1  /* A/B: one launch per k sweeps, ncclSend/ncclRecv of the halo rows in between
                             (single process: ncclCommInitAll; libnccl is dlopen'ed, HS_ERR_NCCL if absent)     */
} hs_exchange;

/* hs_config.flags */
#define HS_FLAG_TOP_IS_SEAM 0x1    /* row-slab: rows exist above this context's buffer          */
#define HS_FLAG_BOTTOM_IS_SEAM 0x2 /* row-slab: rows exist below this context's buffer          */
#define HS_FLAG_FORCE_GENERIC 0x4  /* always use the one-sweep-per-launch kernel (A/B testing)  */
#define HS_FLAG_SINGLE_PHASE 0x8   /* one launch per k fused sweeps (no multi-phase dataflow launch) */
#define HS_FLAG_TEXTBOOK 0x10      /* NOT the reference's arithmetic: Horn & Schunck's 2x2x2-cube gradients
                                      and 1/6-1/12 weighted 3x3 average (what BASELINE.json's prose
                                      describes); window_size must be 3; checked against its own NumPy
                                      oracle (oracle/hs_oracle.py::np_flow_textbook), never against the
                                      reference                                                       */

/*
 * Replaces the three public fields + constructor of class hornSchunck (hornSchunck.cpp:10-17)
 * and adds the geometry the C++ code reads off cv::Mat.  Zero-initialise, set struct_size =
 * sizeof(hs_config), then fill what you need; 0 means "default" for every optional field except
 * `device` (see there).
 */
typedef struct hs_config {
    uint32_t struct_size;
    int32_t width;          /* columns of each frame                                   (required) */
    int32_t height;         /* rows held by this context (whole image, or slab + halo) (required) */
    int32_t window_size;    /* windowSize  (hornSchunck.cpp:10,14), any w >= 1         (required) */
    int32_t max_iterations; /* maxIterations (:10,15), >= 0                            (required) */
    double alpha;           /* alpha (:11,16); 0 is legal and yields IEEE nan/inf like upstream   */
    int32_t batch;          /* independent frame pairs solved per call (default 1)                */
    int32_t device;         /* CUDA ordinal.  NOTE: 0 is ordinal 0, not "default" - a zero-initialised
                               config binds to device 0; pass -1 for the calling thread's current device */
    int32_t temporal_k;     /* sweeps fused per phase (temporal blocking depth k); 0 = auto       */
    uint32_t flags;         /* HS_FLAG_*                                                          */
    /* Row-slab decomposition (one context per GPU).  The context's buffer holds `height` rows of
     * a taller image; it produces rows [out_row_begin, out_row_end) and treats the other rows as
     * halo that the host refreshes between hs_iterate calls.  0/0 = the whole buffer. */
    int32_t out_row_begin;
    int32_t out_row_end;
    void* stream;           /* cudaStream_t to run on; NULL = a private non-blocking stream       */
    int32_t global_row0;    /* row slabs: image row of this context's buffer row 0 (default 0)    */
    /* ---- fields added in 0.2 (struct_size tells the library whether they are present) --------- */
    int32_t precision;      /* hs_precision                                                       */
    int32_t frame_dtype;    /* hs_frame_dtype of the frames (default 8-bit unsigned)              */
    /* Several GPUs behind ONE context, driven by the calling host thread (main.cpp:97-98 stays one
     * call): num_devices > 1 makes hs_create build one child context per device.  hs_solve / hs_upload /
     * hs_solve_device / hs_download / hs_sync / hs_get_timing work on such a context; the result is
     * bit-identical to one GPU.  HS_DECOMP_BATCH deals the `batch` pairs to the devices; HS_DECOMP_ROW_SLAB
     * (batch == 1, fused-kernel windows 2..9) splits the rows.  Listing the same ordinal several times
     * runs that many row slabs on one GPU in a single launch (how the seam protocol is tested on 1 GPU). */
    int32_t num_devices;        /* 0 or 1 = single device (`device`)                              */
    const int32_t* device_ids;  /* num_devices ordinals; NULL = 0 .. num_devices-1                */
    int32_t decomposition;      /* hs_decomposition                                               */
    int32_t exchange;           /* hs_exchange (row slabs)                                        */
    /* One PROCESS per GPU (torchrun / MPI style): this context is slab `slab_rank` of `slab_world` of an
     * image `height` rows tall.  hs_get_slab_info says which frame rows to upload; neighbours are wired
     * with hs_slab_export / hs_slab_connect (CUDA IPC), after which hs_iterate exchanges halos itself. */
    int32_t slab_world;
    int32_t slab_rank;
} hs_config;

typedef struct hs_ctx hs_ctx;

/* cudaEvent-measured milliseconds of the last hs_solve / hs_gradients call. */
typedef struct hs_timing {
    float h2d_ms;
    float prepare_ms; /* gradient + coefficient stage                    */
    float iterate_ms; /* all Jacobi sweeps                               */
    float d2h_ms;     /* widen + device->host                            */
    float total_ms;
    int32_t launches;   /* kernels launched by the last call              */
    int32_t temporal_k; /* sweeps fused per phase (temporal blocking depth)*/
    int32_t kernel_id;  /* 0 = generic one-sweep kernel, 1 = fused tile kernel, 2 = fp64 reference-arithmetic kernel */
} hs_timing;

/* Device-side view used by device-resident callers (bench, row-slab host code). All pointers
 * are device addresses owned by the context and stay valid until hs_destroy. */
typedef struct hs_device_view {
    void* prev;           /* uint8, frame_pitch bytes per row; rows = height + seam rows          */
    void* next;
    size_t frame_pitch;   /* bytes                                                                */
    size_t frame_pair_stride; /* bytes between consecutive pairs of the batch                     */
    int32_t frame_rows;   /* height + (top seam ? 1 : 0) + (bottom seam ? 1 : 0)                  */
    int32_t frame_row0;   /* buffer row of the context's row 0 inside prev/next (0 or 1)          */
    float* uv;            /* CURRENT flow plane, {u, v} INTERLEAVED per pixel (uv[2*x] = u, uv[2*x+1] = v):
                             the fused kernel feeds both to Blackwell's packed-fp32 instructions.  The
                             two planes ping-pong: re-query after hs_iterate                       */
    void* reserved;
    size_t flow_pitch;    /* bytes per row of uv (8 bytes per pixel)                              */
    size_t flow_pair_stride; /* bytes between pairs                                               */
    int32_t width, height, batch;
    int32_t halo_rows_top;    /* rows of halo a neighbour must refresh per fused launch: a*k      */
    int32_t halo_rows_bottom; /* (w/2)*k                                                          */
} hs_device_view;

/* ---- lifetime -------------------------------------------------------------------------- */
/* hornSchunck::hornSchunck(int,int,double)  hornSchunck.cpp:13-17 (+ geometry). */
int hs_create(const hs_config* cfg, hs_ctx** out);
void hs_destroy(hs_ctx* ctx);

/* ---- the reference's two public methods --------------------------------------------------- */
/* hornSchunck::getFlow  hornSchunck.cpp:43-75 / main.cpp:98.  prev/next: host uint8, `batch`
 * images each (image i starts at prev + i*prev_image_stride bytes; pass 0 for batch == 1).
 * u/v: host outputs of out_dtype, same batching rule.  Synchronous. */
int hs_solve(hs_ctx* ctx,
             const uint8_t* prev, size_t prev_row_stride, size_t prev_image_stride,
             const uint8_t* next, size_t next_row_stride, size_t next_image_stride,
             void* u, size_t u_row_stride, size_t u_image_stride,
             void* v, size_t v_row_stride, size_t v_image_stride,
             int out_dtype);

/* hs_solve split in two so that independent pairs can overlap: hs_solve_async queues upload, solve and download and
 * returns; hs_solve_wait blocks until the flow of that call is in u, v.  The host<->device copies run on a transfer
 * stream owned by the context, the kernels on its compute stream (hs_config.stream): give several contexts the SAME
 * compute stream and their kernels run back to back in call order while the copies of one call overlap the sweeps
 * of the next (bench.py: `e2e.value_two_contexts_in_flight`, the batch256 workload).  Use page-locked host buffers
 * (hs_host_alloc), keep them untouched until hs_solve_wait, and have at most one call per context in flight; any
 * other work queued on the context (hs_solve_device ...) must have been waited for with hs_sync first.
 * Single-device HS_PREC_F32 whole-image contexts. */
int hs_solve_async(hs_ctx* ctx,
                   const uint8_t* prev, size_t prev_row_stride, size_t prev_image_stride,
                   const uint8_t* next, size_t next_row_stride, size_t next_image_stride,
                   void* u, size_t u_row_stride, size_t u_image_stride,
                   void* v, size_t v_row_stride, size_t v_image_stride,
                   int out_dtype);
int hs_solve_wait(hs_ctx* ctx);

/* preprocess() + getFlow in one call (HornSchunckOF/main.cpp:11-26,84 then :98): prev/next are 8UC3
 * BGR frames (row stride >= 3*width bytes, a multiple of 4); the BGR2GRAY conversion runs on the
 * device with OpenCV's fixed-point luma (bit-exact with cv::cvtColor).  batch == 1 contexts. */
int hs_solve_bgr(hs_ctx* ctx,
                 const uint8_t* prev_bgr, size_t prev_row_stride,
                 const uint8_t* next_bgr, size_t next_row_stride,
                 void* u, size_t u_row_stride, void* v, size_t v_row_stride, int out_dtype);

/* hornSchunck::getGradients  hornSchunck.cpp:19-41.  gx/gy/gt: host outputs of out_dtype, all
 * with the same row stride (batch == 1 contexts only). */
int hs_gradients(hs_ctx* ctx,
                 const uint8_t* prev, size_t prev_row_stride,
                 const uint8_t* next, size_t next_row_stride,
                 void* gx, void* gy, void* gt, size_t out_row_stride, int out_dtype);

/* ---- device-resident pipeline (what hs_solve is made of) ------------------------------------ */
int hs_upload(hs_ctx* ctx,                                    /* host uint8 -> device frames     */
              const uint8_t* prev, size_t prev_row_stride, size_t prev_image_stride,
              const uint8_t* next, size_t next_row_stride, size_t next_image_stride);
int hs_prepare(hs_ctx* ctx);                                  /* gradients+coefficients, u=v=0   */
/* `iterations` more Jacobi sweeps.  Whole-image contexts run them as ONE launch of ceil(n/k) phases
 * (tiles synchronise through per-tile counters); row-slab contexts and small images launch once
 * per k sweeps, so that the host can refresh halos in between. */
int hs_iterate(hs_ctx* ctx, int iterations);
/* Row-slab helper: `sweeps` (<= temporal_k) fused sweeps producing only buffer rows
 * [row_begin, row_end) of the NEXT flow planes (any rows of the buffer, halo rows included); the
 * current planes become the next ones when `flip` is non-zero (pass it on the last partial launch
 * of a step).  Lets the host (a) compute the rows next to a seam first, start their halo exchange
 * and overlap it with the interior rows, and (b) keep a halo several launches deep and advance
 * part of it redundantly, so that halos are exchanged only every few launches. */
int hs_iterate_rows(hs_ctx* ctx, int sweeps, int row_begin, int row_end, int flip);
/* Early exit (contract extension, the reference always runs maxIterations sweeps, hornSchunck.cpp:56):
 * sweeps in chunks of `check_every`; after each chunk the residual max(|u_n - u_{n-1}|, |v_n - v_{n-1}|)
 * of the last sweep is reduced on the device; stops at the first chunk whose residual is <= tolerance,
 * or after max_sweeps.  The flow after *sweeps_done sweeps is bit-identical to hs_iterate(*sweeps_done). */
int hs_iterate_until(hs_ctx* ctx, int max_sweeps, double tolerance, int check_every, int* sweeps_done,
                     double* residual);
int hs_solve_device(hs_ctx* ctx);                             /* prepare + iterate(max_iterations)*/
int hs_download(hs_ctx* ctx,                                  /* device u,v -> host              */
                void* u, size_t u_row_stride, size_t u_image_stride,
                void* v, size_t v_row_stride, size_t v_image_stride, int out_dtype);
int hs_sync(hs_ctx* ctx);                                     /* wait for the context's stream   */
/* Only what the reference's plot consumes: plotFlow::plotBresenhamLine (plotFlow.cpp:68-88) reads
 * u, v at rows/columns that are multiples of `delta` (20 in main.cpp:104).  Gathers those samples of
 * the current device-resident flow as doubles, row-major ny x nx, ny = ceil(rows/delta),
 * nx = ceil(width/delta) (1 197 values instead of 7.5 MB for the bundled 1242x375 pair).
 * Pass u == v == NULL to query ny / nx only. */
int hs_sample_grid(hs_ctx* ctx, int delta, double* u, double* v, int* ny, int* nx);
int hs_get_device_view(hs_ctx* ctx, hs_device_view* out);

/* ---- streaming front-end for frame sequences (the .mp4 branch, HornSchunckOF/main.cpp:53-59) -------
 * Every frame is uploaded once and serves as `next` of one pair and `prev` of the following one.
 * hs_video_push(frame n) queues the solve of pair (n-1, n) and, while that runs, returns the flow
 * of the PREVIOUS pair (n-2, n-1) in u, v: *pair_index = n-2, or -1 when no flow is due yet (then
 * u, v are not touched).  Upload, solve and download of consecutive pairs overlap on two streams.
 * hs_video_flush returns the last pair; hs_video_reset starts a new sequence.  out_dtype must stay
 * the same within a sequence.  batch == 1 contexts; use pinned host buffers for full overlap. */
int hs_video_push(hs_ctx* ctx, const uint8_t* frame, size_t row_stride,
                  void* u, size_t u_row_stride, void* v, size_t v_row_stride, int out_dtype, int* pair_index);
int hs_video_flush(hs_ctx* ctx, void* u, size_t u_row_stride, void* v, size_t v_row_stride, int* pair_index);
int hs_video_reset(hs_ctx* ctx);

/* ---- row slabs across processes (one rank per GPU) ------------------------------------------------ */
/* Geometry of this context's slab inside the full image (valid for slab_world > 1 contexts and for
 * the children of a row-slab group).  Rows are image rows. */
typedef struct hs_slab_info {
    int32_t rank, world;
    int32_t own_begin, own_end;     /* rows this slab produces                                       */
    int32_t buf_begin, buf_end;     /* rows its flow planes hold (own rows + halo)                   */
    int32_t frame_begin, frame_end; /* rows of prev/next to pass to hs_upload (buffer + 1 Sobel row per seam) */
    int32_t temporal_k;
    int32_t halo_top, halo_bottom;  /* rows a neighbour writes into this slab per phase              */
} hs_slab_info;
int hs_get_slab_info(hs_ctx* ctx, hs_slab_info* out);
/* The same plan as pure host arithmetic (no device needed): bands of rows with even first rows, a*k (+1
 * when odd) halo rows above and (w/2)*k below each seam, one more frame row per seam for the Sobel taps. */
int hs_plan_slab(int32_t height, int32_t world, int32_t rank, int32_t window_size, int32_t temporal_k,
                 hs_slab_info* out);

/* Opaque, fixed-size, position-independent description of where a slab's flow planes and seam flags
 * live: ship it to the neighbouring ranks by any means (an all-gather of 256 bytes). */
typedef struct hs_slab_handle { uint8_t bytes[256]; } hs_slab_handle;
int hs_slab_export(hs_ctx* ctx, hs_slab_handle* out);
/* Wire this slab to the slab above (rank-1) and below (rank+1); pass NULL at the image border.  Opens
 * the neighbours' memory (cudaIpcOpenMemHandle; same-process handles are used directly with peer
 * access).  Afterwards hs_iterate(n) runs all n sweeps in ONE launch.  Every rank must (1) call
 * hs_prepare, (2) synchronise and barrier with its neighbours, (3) call hs_iterate with the same n;
 * and must not call hs_prepare again before the neighbours' hs_iterate has completed. */
int hs_slab_connect(hs_ctx* ctx, const hs_slab_handle* up, const hs_slab_handle* down);

/* ---- introspection --------------------------------------------------------------------------- */
int hs_get_timing(const hs_ctx* ctx, hs_timing* out);
/* The temporal-blocking depth hs_create would pick for this configuration on a device with num_sms SMs
 * (pure host arithmetic, no device needed; 0 = no fused kernel for this window).  The rule is documented in
 * DESIGN.md section 4 and checked against the measured k sweeps committed under profiles/. */
int hs_default_temporal_k(const hs_config* cfg, int32_t num_sms);
const char* hs_last_error(const hs_ctx* ctx); /* ctx may be NULL: last hs_create failure of thread */
int hs_version(void);

/* Pinned host memory for callers that want zero-copy-speed transfers (optional). */
int hs_host_alloc(void** ptr, size_t bytes);
int hs_host_free(void* ptr);

#ifdef __cplusplus
}
#endif
#endif /* HS_B200_H */
