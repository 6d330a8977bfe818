// C ABI of libhs_b200.so (include/hs.h): context, device buffers, TMA descriptors, launches.
// Replaces class hornSchunck (/root/reference/HornSchunckOF/hornSchunck.cpp:8-76) behind
// hs_create / hs_solve / hs_destroy.  No CPU fallback anywhere in this file.
#include "../../include/hs.h"

#include <cudaTypedefs.h>

#include <algorithm>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <new>
#include <string>
#include <vector>

#include "hs_kernels.cuh"
#include "hs_kernels_f64.cuh"

namespace {

constexpr int TILE_R = 4;       // rows per thread in the fused kernel
constexpr int TILE_NWARP = 12;  // warps per CTA (staged tile 128 x 48)

thread_local std::string g_create_err;

inline int round_up(int v, int m) { return (v + m - 1) / m * m; }

struct DevGuard {
    int prev = -1;
    bool ok = true;
    explicit DevGuard(int dev) {
        if (cudaGetDevice(&prev) != cudaSuccess) { ok = false; prev = -1; }
        if (ok && prev != dev) ok = (cudaSetDevice(dev) == cudaSuccess);
    }
    ~DevGuard() {
        if (prev >= 0) cudaSetDevice(prev);
    }
};

}  // namespace

struct hs_ctx {
    hs_config cfg{};
    int dev = 0;
    cudaStream_t stream = nullptr;
    bool own_stream = false;

    int W = 0, H = 0, B = 1, w = 0, a = 0, RL = 0, RR = 0, T = 0;
    double alpha = 0;
    int pitch = 0;
    long long plane = 0;
    int oy0 = 0, oy1 = 0, grow0 = 0;
    bool top_seam = false, bot_seam = false;
    bool textbook = false;       // HS_FLAG_TEXTBOOK: cube gradients + weighted 3x3 average (non-parity extra)

    int frows = 0, frow0 = 0;
    size_t fpitch = 0, fimg = 0;
    // HS_PREC_F64 contexts (hs_kernels_f64.cuh): frames of any depth, every plane in fp64
    bool f64 = false;
    int fdt = HS_FRAME_U8, fes = 1;          // frame element type / size in bytes
    double* d64[8] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};   // gx, gy, gt, den, u0, v0, u1, v1
    uint8_t* d_prev = nullptr;
    uint8_t* d_next = nullptr;
    float2* d_uv[2] = {nullptr, nullptr};   // flow planes, {u, v} interleaved per pixel (packed-fp32 operands), ping-pong
    int cur = 0;
    uint32_t* d_cpk = nullptr;   // packed {Ix, Iy, It}
    int* d_done = nullptr;       // per-tile phase counters of a multi-phase launch
    size_t done_cap = 0;
    bool multi_phase = true;     // whole hs_iterate in one launch (single-GPU) vs one launch per k sweeps
    float* d_inv = nullptr;
    void* d_out = nullptr;
    size_t out_bytes = 0;
    uint8_t* d_bgr = nullptr;    // staging for hs_solve_bgr: two 8UC3 frames
    uint8_t* d_flat = nullptr;   // staging for dense host frames whose rows are not 128-byte multiples (do_upload)
    unsigned int* d_resid = nullptr;   // hs_iterate_until: max-abs-difference accumulator
    // streaming front-end (hs_video_*): frame ring, copy stream, double-buffered output staging
    uint8_t* d_ring[3] = {nullptr, nullptr, nullptr};   // [0], [1] are the context's own prev/next
    void* d_vout[2] = {nullptr, nullptr};
    cudaStream_t copy_stream = nullptr;
    cudaEvent_t ev_up = nullptr, ev_k1[3] = {nullptr, nullptr, nullptr}, ev_solved[2] = {nullptr, nullptr};
    cudaEvent_t ev_async[2] = {nullptr, nullptr};   // hs_solve_async: frames uploaded / flow staged
    bool async_pending = false;
    cudaStream_t xfer = nullptr; // stream the host<->device copies of do_upload / do_download go to (null = `stream`)
    int vid_frames = 0;          // frames pushed so far
    int vid_pending = -1;        // pair solved (or being solved) whose flow was not handed out yet
    int vid_dtype = -1;
    size_t bgr_pitch = 0;

    CUtensorMap tm_uv[2], tm_cpk, tm_inv;
    // Row-slab seams (multi-GPU): where each neighbouring slab lives.  Set by slab_link(); a linked
    // context runs all sweeps of hs_iterate in ONE launch and exchanges halos from inside the kernel.
    struct Seam {
        bool on = false;
        float2* uv[2] = {nullptr, nullptr};  // the neighbour's planes (peer-mapped device memory)
        int dy = 0;                         // my buffer row y is the neighbour's buffer row y + dy
        int* inbox = nullptr;               // the neighbour's inbox (it polls it; my seam tiles publish there)
        int nbr_rows = 0, nbr_parity = 0;   // the neighbour's produced rows and the image-row parity of its first one
    };
    Seam up, dn;
    int* d_inbox = nullptr;      // [2][SEAM_JMAX][tiles_x]: phase counts published by the neighbours
    void* arena = nullptr;       // seam contexts: u[2], v[2] and the inbox live in ONE allocation (one IPC handle)
    size_t arena_bytes = 0, inbox_off = 0, plane_off[2] = {0, 0};
    void* ipc_up = nullptr;      // arenas of the neighbours opened with cudaIpcOpenMemHandle (multi-process)
    void* ipc_dn = nullptr;
    bool linked = false;
    int reverse = 0;             // walk tile rows bottom-up (odd slabs)
    // row-slab geometry inside the full image (slab_world > 1 contexts and children of a row-slab group)
    int slab_rank = 0, slab_world = 1, img_h = 0;
    int sl_y0 = 0, sl_y1 = 0, sl_b0 = 0, sl_b1 = 0, sl_f0 = 0, sl_f1 = 0;
    // several GPUs behind one context (hs_config.num_devices > 1): the children, one per device
    std::vector<hs_ctx*> kids;
    int decomp = 0, exchange = 0;
    bool emulate = false;        // all children on ONE device: their slabs run in a single launch
    cudaEvent_t ev_x = nullptr;  // child: "my last prepare / iterate is done" for the neighbours' streams
    struct NcclState* nccl = nullptr;
    int phase_count = 0;         // phases completed since hs_prepare
    int max_ctas = 0;            // 0 = one CTA per SM
    int kernel_id = 0;  // 0 generic, 1 fused tile
    int num_sms = 148;
    bool use_pdl = true;
    int k = 1;

    cudaEvent_t ev[6] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
    hs_timing timing{};
    bool uploaded = false, prepared = false, timing_pending = false;
    std::string err;

    hs::Geom geom() const {
        hs::Geom g;
        g.W = W; g.H = H; g.pitch = pitch; g.plane = plane; g.oy0 = oy0; g.oy1 = oy1; g.grow0 = grow0;
        return g;
    }
};

namespace {

int fail(hs_ctx* c, int code, const char* fmt, ...) {
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof buf, fmt, ap);
    va_end(ap);
    if (c) c->err = buf; else g_create_err = buf;
    return code;
}

#define HS_CUDA(c, call)                                                                        \
    do {                                                                                        \
        cudaError_t e__ = (call);                                                               \
        if (e__ != cudaSuccess)                                                                 \
            return fail((c), e__ == cudaErrorMemoryAllocation ? HS_ERR_OOM : HS_ERR_CUDA,       \
                        "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e__), __FILE__,      \
                        __LINE__);                                                              \
    } while (0)

// cuTensorMapEncodeTiled through the runtime's driver entry point: no -lcuda at link time.
PFN_cuTensorMapEncodeTiled_v12000 tensor_map_encoder() {
    static PFN_cuTensorMapEncodeTiled_v12000 fn = nullptr;
    if (!fn) {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
            q == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<PFN_cuTensorMapEncodeTiled_v12000>(p);
    }
    return fn;
}

// rank-3 map over (x, y, pair) of a pitched plane; box = SX x SY x 1; OOB reads as zero
int make_map(hs_ctx* c, CUtensorMap* m, void* base, CUtensorMapDataType dt, int esize, int box_x,
             int box_y, int per_px = 1) {
    auto enc = tensor_map_encoder();
    if (!enc) return fail(c, HS_ERR_CUDA, "cuTensorMapEncodeTiled entry point not available");
    cuuint64_t dims[3] = {(cuuint64_t)c->W * per_px, (cuuint64_t)c->H, (cuuint64_t)c->B};
    cuuint64_t strides[2] = {(cuuint64_t)c->pitch * per_px * esize, (cuuint64_t)c->plane * per_px * esize};
    cuuint32_t box[3] = {(cuuint32_t)box_x, (cuuint32_t)box_y, 1};
    cuuint32_t estr[3] = {1, 1, 1};
    CUresult r = enc(m, dt, 3, base, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                     CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return fail(c, HS_ERR_CUDA, "cuTensorMapEncodeTiled failed (CUresult %d)", (int)r);
    return HS_OK;
}

// The flow plane ({u, v} interleaved, 8 bytes per pixel) as a rank-4 map (32-byte chunk, chunk of the row, y,
// pair) with SWIZZLE_32B.  The staged tile has the same linear layout as a rank-3 box, except that the two
// 16-byte halves of a 32-byte chunk trade places in every other 128-byte line (address bit 4 ^= bit 7).  A lane's
// four pixels are one 32-byte chunk; with the swizzle, lanes 0-3 and 4-7 of a quarter warp find the half they
// want in different bank groups, so the two LDS.128 of a patch row are conflict-free (they were 2-way
// conflicted: 8 % of the kernel's shared-memory wavefronts).  Columns W .. round_up(W, 4) of the last chunk
// are read from the pad columns of the plane instead of being zero-filled: they are zero by construction
// (memset at hs_prepare; the kernels only ever store zeros beyond column W).
int make_map_uv(hs_ctx* c, CUtensorMap* m, void* base, int box_x_px, int box_y) {
    auto enc = tensor_map_encoder();
    if (!enc) return fail(c, HS_ERR_CUDA, "cuTensorMapEncodeTiled entry point not available");
    cuuint64_t dims[4] = {8, (cuuint64_t)(c->W + 3) / 4, (cuuint64_t)c->H, (cuuint64_t)c->B};
    cuuint64_t strides[3] = {32, (cuuint64_t)c->pitch * 8, (cuuint64_t)c->plane * 8};
    cuuint32_t box[4] = {8, (cuuint32_t)box_x_px / 4, (cuuint32_t)box_y, 1};
    cuuint32_t estr[4] = {1, 1, 1, 1};
    CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, base, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                     CU_TENSOR_MAP_SWIZZLE_32B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return fail(c, HS_ERR_CUDA, "cuTensorMapEncodeTiled (flow plane) failed (CUresult %d)", (int)r);
    return HS_OK;
}

int env_int(const char* name, int dflt) {
    const char* s = getenv(name);
    return (s && *s) ? atoi(s) : dflt;
}

// Temporal-blocking depth k (sweeps fused per phase) of the fused kernel: pure host arithmetic, exported
// as hs_default_temporal_k so that the choice can be checked against the measured sweeps in profiles/.
//   `asked` > 0: the caller's k, only clamped to what a 128 x 48 staged tile can hold.
//   Large frames: measured on B200 (profiles/r02w_k_sweep*.jsonl, k = 1..12 for every window that has a fused
//   kernel; DRAM side in profiles/*traffic_table*): w=2 k=8, w=3 k=6, w=4 k=4, w=5 k=3, w=6..9 k=2.  (Until the
//   tile load lost its bank conflicts, w=3 wanted k=4 while the working set sat in L2; now a staged tile is
//   cheap enough to load that the deeper k wins everywhere.)
//   Small and medium frames (less than three tiles per SM) choose k from a cost model of a phase
//   (microseconds, fitted on B200 to profiles/r02al_k_sweep.jsonl and checked against it by
//   tests/test_abi_cpu.py::test_default_k_is_near_the_measured_best):
//     a tile costs          item(k)  = k * t_sweep + t_tile
//     chained launches      phase(k) = ceil(tiles / #SMs) * item + t_launch     (tiles < 1.25 #SMs)
//     one dataflow launch   phase(k) = max(tiles / #SMs * item, item + t_dep)
//   and the k with the lowest ceil(T / k) * phase(k) (+ t_coop for a dataflow launch) wins.  t_dep is the publish -> poll -> TMA chain from a finished
//   tile to its dependants: with few tiles per SM it, not the arithmetic, paces a phase, and fusing more
//   sweeps per phase amortises it.
inline int large_frame_k(int RL, int RR) {
    const int w = RL + RR + 1;
    return w <= 2 ? 8 : (w == 3 ? 6 : (w == 4 ? 4 : (w == 5 ? 3 : 2)));
}
int choose_temporal_k(int asked, int W, int rows, int row_parity, int B, int RL, int RR, int max_iterations, int num_sms,
                      bool may_dataflow, bool seam) {
    const int SY = TILE_R * TILE_NWARP, SX = 128;
    const int kmax = (SY - 3) / std::max(1, RL + RR);
    const int rad = std::max(RL, RR);
    auto tiles = [&](int k) -> size_t {
        const int vx = SX - round_up(RL * k, 4) - round_up(RR * k, 4);
        const int hyt = RL * k + ((row_parity + RL * k) & 1);
        const int vy = (SY - hyt - RR * k) & ~1;
        if (vx <= 0 || vy <= 0) return 0;
        return (size_t)((W + vx - 1) / vx) * ((rows + vy - 1) / vy) * B;
    };
    int k = asked;
    if (k <= 0) k = large_frame_k(RL, RR);
    k = std::min(k, kmax);
    // keep a useful centre: at least a quarter of the staged rows must be output rows
    while (k > 1 && SY - (RL + RR) * k < SY / 4) --k;
    if (asked <= 0 && !seam && tiles(k) < (size_t)3 * num_sms) {
        const double t_sweep = rad <= 1 ? 0.38 : 0.7, t_tile = 1.2, t_launch = 4.0, t_dep = 6.0, t_coop = 40.0;
        int kcap = std::min(kmax, rad <= 1 ? 12 : 7);
        if (max_iterations > 0) kcap = std::min(kcap, max_iterations);
        const int T = max_iterations > 0 ? max_iterations : 1000;
        double best = 1e300;
        for (int kk = 1; kk <= kcap; ++kk) {
            if (kk > 1 && SY - (RL + RR) * kk < SY / 4) break;
            const size_t n = tiles(kk);
            if (n == 0) break;
            const double item = kk * t_sweep + t_tile;
            const bool dataflow = may_dataflow && n * 4 >= (size_t)num_sms * 5;
            const double phase = dataflow ? std::max((double)n / num_sms * item, item + t_dep)
                                          : (double)((n + num_sms - 1) / num_sms) * item + t_launch;
            // whole solve: ceil(T / k) phases (the last one may be short but costs a full tile round), plus the
            // fixed cost of the cooperative launch (launch + counter memset + descriptor prefetch: ~40 us measured
            // on the bundled pair, profiles/r02y_traffic_table.txt: k=7 206 against k=4 243 Gpix-it/s at T=100)
            const double total = (double)((T + kk - 1) / kk) * phase + (dataflow ? t_coop : 0.0);
            if (total < best) { best = total; k = kk; }
        }
    }
    return k;
}

constexpr int EMU_MAXS = 4;     // row slabs of ONE device that a single (emulation) launch can hold

// tile-row geometry of `rows` produced rows whose first row has image-row parity `parity`
struct RowTiling {
    int hyt, vy, tiles_y, jt, jb;
};

template <int RL, int RR, bool TB = false>
struct Tile {
    using TS = hs::TileShape<RL, RR, TILE_R, TILE_NWARP>;
    template <int MAXS>
    static auto kernel() { return hs::k_jacobi_tile<RL, RR, TILE_R, TILE_NWARP, TB, MAXS>; }
    static cudaError_t configure() {
        cudaError_t e = cudaFuncSetAttribute(kernel<1>(), cudaFuncAttributeMaxDynamicSharedMemorySize, (int)TS::SMEM);
        if (e == cudaSuccess)
            e = cudaFuncSetAttribute(kernel<EMU_MAXS>(), cudaFuncAttributeMaxDynamicSharedMemorySize, (int)TS::SMEM);
        return e;
    }
    static int max_k() { return (TS::SY - 3) / std::max(1, RL + RR); }
    // every staged tile must start on an EVEN image row (the canonical column sums pair rows
    // (2i, 2i+1), and a thread's 4-row patch shares those pair sums): one more halo row above
    // when needed, and an even tile pitch
    static RowTiling tiling(int k, int rows, int parity) {
        RowTiling t;
        t.hyt = RL * k + ((parity + RL * k) & 1);
        t.vy = (TS::SY - t.hyt - RR * k) & ~1;
        if (t.vy <= 0) { t.tiles_y = 0; t.jt = t.jb = 0; return t; }
        t.tiles_y = (rows + t.vy - 1) / t.vy;
        // tile rows whose stored centre lies within max(RL, RR) * k rows of the slab's first / last row:
        // they read the neighbouring slab's rows and produce the rows it reads
        const int reach = std::max(RL, RR) * k;
        t.jt = std::min(t.tiles_y, (reach + t.vy - 1) / t.vy);
        t.jb = t.tiles_y - std::max(rows - reach, 0) / t.vy;
        return t;
    }
    static int vx_of(int k) { return TS::SX - round_up(RL * k, 4) - round_up(RR * k, 4); }
    // The multi-phase dataflow launch waits for the 3 x 3 adjacent tile counters only, so a staged
    // tile must not reach past its direct neighbours: both halos have to fit inside one tile pitch.
    // (w=5 from k=9, w=4 from k=10, w=3 from k=16 do not; those run as chained single-phase launches.)
    static bool dataflow_ok(int k, int row_parity) {
        const int hxl = round_up(RL * k, 4), hxr = round_up(RR * k, 4);
        const RowTiling t = tiling(k, 1 << 20, row_parity);
        const int vx = vx_of(k);
        return vx > 0 && t.vy > 0 && hxl <= vx && hxr <= vx && t.hyt <= t.vy && RR * k <= t.vy;
    }
    // Describe what context c contributes to a launch that produces its rows [row0, row1).
    static bool fill_slab(const hs_ctx* c, hs::SlabDesc& S, int k, int row0, int row1, int tile0, bool seams) {
        memset(&S, 0, sizeof S);
        for (int i = 0; i < 2; ++i) {
            S.tm_uv[i] = c->tm_uv[i];
            S.uv[i] = c->d_uv[i];
        }
        S.tm_cpk = c->tm_cpk; S.tm_inv = c->tm_inv;
        S.g = c->geom();
        S.g.oy0 = row0;                                        // rows this launch produces
        S.g.oy1 = row1;
        const RowTiling t = tiling(k, row1 - row0, row0 + c->grow0);
        const int vx = vx_of(k);
        if (vx <= 0 || t.vy <= 0) return false;
        S.hyt = t.hyt; S.vy = t.vy; S.tiles_y = t.tiles_y;
        const int tiles_x = (c->W + vx - 1) / vx;
        S.ntiles = tiles_x * t.tiles_y * c->B;
        S.tile0 = tile0;
        S.fd_per_img = hs::FastDiv::make((uint32_t)(tiles_x * t.tiles_y));
        S.jt = t.jt; S.jb = t.jb;
        if (seams) {
            if (t.jt > hs::SEAM_JMAX || t.jb > hs::SEAM_JMAX) return false;
            S.reverse = c->reverse;
            S.inbox = c->d_inbox;
            const hs_ctx::Seam* side[2] = {&c->up, &c->dn};
            for (int q = 0; q < 2; ++q) {
                const hs_ctx::Seam& L = *side[q];
                if (!L.on) continue;
                const RowTiling nt = tiling(k, L.nbr_rows, L.nbr_parity);
                if (nt.vy != t.vy || nt.jt > hs::SEAM_JMAX || nt.jb > hs::SEAM_JMAX) return false;   // tile columns AND pitch must agree
                if (q == 0) {
                    for (int i = 0; i < 2; ++i) S.up_uv[i] = L.uv[i];
                    S.up_dy = L.dy; S.push_up = RR * k; S.up_j = nt.jb;
                    S.out_up = L.inbox + (size_t)1 * hs::SEAM_JMAX * tiles_x;   // I am "the slab below" for it
                } else {
                    for (int i = 0; i < 2; ++i) S.dn_uv[i] = L.uv[i];
                    S.dn_dy = L.dy; S.push_dn = RL * k; S.dn_j = nt.jt;
                    S.out_dn = L.inbox;                                          // I am "the slab above" for it
                }
            }
        }
        return true;
    }
    // One launch advances `sweeps` sweeps in phases of k on n contexts of ONE device (n == 1 except for
    // the single-device emulation of row slabs).  phases > 1 - or any seam - needs every CTA resident
    // at once (they wait on each other's tiles) -> cooperative launch; a single phase is chained to
    // the previous launch with programmatic dependent launch instead.
    template <int MAXS>
    static cudaError_t launch_n(hs_ctx* const* cs, int n, int k, int sweeps, int row0, int row1, bool seams) {
        hs_ctx* c = cs[0];
        hs::LaunchDesc<MAXS> d;
        memset(&d, 0, sizeof d);
        int total = 0;
        for (int i = 0; i < n; ++i) {
            const int r0 = n == 1 ? row0 : cs[i]->oy0, r1 = n == 1 ? row1 : cs[i]->oy1;
            if (!fill_slab(cs[i], d.s[i], k, r0, r1, total, seams)) return cudaErrorInvalidValue;
            total += d.s[i].ntiles;
        }
        d.nslabs = n; d.k = k; d.sweeps = sweeps; d.cur = c->cur; d.phase_base = c->phase_count;
        d.hxl = round_up(RL * k, 4);
        d.vx = vx_of(k);
        d.tiles_x = (c->W + d.vx - 1) / d.vx;
        d.fd_tiles_x = hs::FastDiv::make((uint32_t)d.tiles_x);
        d.ntiles = total;
        d.done = c->d_done;
        d.kf = 1.0f / (float)(c->w * c->w);
        d.alpha2 = (float)(c->alpha * c->alpha);
        const int phases = (sweeps + k - 1) / k;
        bool any_seam = false;
        for (int i = 0; i < n; ++i) any_seam |= d.s[i].out_up || d.s[i].out_dn || d.s[i].up_j || d.s[i].dn_j;
        if (phases > 1 && !dataflow_ok(k, row0 + c->grow0)) return cudaErrorInvalidValue;
        if (any_seam && !dataflow_ok(k, row0 + c->grow0)) return cudaErrorInvalidValue;
        int grid = std::min(total, c->num_sms);                // persistent: one CTA per SM
        if (c->max_ctas > 0) grid = std::min(grid, c->max_ctas);
        if (phases > 1 || any_seam) {
            if ((size_t)total > c->done_cap) return cudaErrorInvalidValue;
            cudaError_t e = cudaMemsetAsync(c->d_done, 0, (size_t)total * sizeof(int), c->stream);
            if (e != cudaSuccess) return e;
        }
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = dim3(grid);
        cfg.blockDim = dim3(TS::THREADS);
        cfg.dynamicSmemBytes = TS::SMEM;
        cfg.stream = c->stream;
        cudaLaunchAttribute attr[1];
        if (phases > 1 || any_seam) {
            attr[0].id = cudaLaunchAttributeCooperative;
            attr[0].val.cooperative = 1;
        } else {
            attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;   // PDL: see pdl_wait() in the kernel
            attr[0].val.programmaticStreamSerializationAllowed = c->use_pdl ? 1 : 0;
        }
        cfg.attrs = attr;
        cfg.numAttrs = 1;
        return cudaLaunchKernelEx(&cfg, kernel<MAXS>(), d);
    }
    static cudaError_t launch(hs_ctx* c, int k, int sweeps, int row0, int row1, bool seams = false) {
        return launch_n<1>(&c, 1, k, sweeps, row0, row1, seams);
    }
    static cudaError_t launch_group(hs_ctx* const* cs, int n, int k, int sweeps) {
        if (n < 1 || n > EMU_MAXS) return cudaErrorInvalidValue;
        return launch_n<EMU_MAXS>(cs, n, k, sweeps, cs[0]->oy0, cs[0]->oy1, true);
    }
    static size_t tiles_for(const hs_ctx* c, int k) {          // tiles of one phase
        const int vx = vx_of(k);
        const RowTiling t = tiling(k, c->oy1 - c->oy0, c->oy0 + c->grow0);
        if (vx <= 0 || t.vy <= 0) return 0;
        return (size_t)((c->W + vx - 1) / vx) * t.tiles_y * c->B;
    }
    static size_t max_tiles(const hs_ctx* c) {                 // upper bound over all k (k = 1 tiles are the largest)
        size_t best = 0;
        for (int k = 1; k <= max_k(); ++k) {
            const int vx = vx_of(k), vy = (TS::SY - (RL + RR) * k - 1) & ~1;
            if (vx <= 0 || vy <= 0) break;
            const size_t n = (size_t)((c->W + vx - 1) / vx) * ((c->oy1 - c->oy0 + vy - 1) / vy) * c->B;
            best = std::max(best, n);
        }
        return best;
    }
};

// dispatch on (RL, RR) = (anchor, w - 1 - anchor); returns false when no fused kernel exists
template <typename F>
bool tile_dispatch(const hs_ctx* c, F&& f) {
    const int RL = c->RL, RR = c->RR;
    if (c->textbook) { f(Tile<1, 1, true>{}); return true; }    // weighted 3x3 average
    if (RL == 0 && RR == 1) { f(Tile<0, 1>{}); return true; }   // w = 2
    if (RL == 1 && RR == 1) { f(Tile<1, 1>{}); return true; }   // w = 3
    if (RL == 1 && RR == 2) { f(Tile<1, 2>{}); return true; }   // w = 4
    if (RL == 2 && RR == 2) { f(Tile<2, 2>{}); return true; }   // w = 5
    if (RL == 2 && RR == 3) { f(Tile<2, 3>{}); return true; }   // w = 6
    if (RL == 3 && RR == 3) { f(Tile<3, 3>{}); return true; }   // w = 7
    if (RL == 3 && RR == 4) { f(Tile<3, 4>{}); return true; }   // w = 8
    if (RL == 4 && RR == 4) { f(Tile<4, 4>{}); return true; }   // w = 9
    return false;
}


int ensure_out(hs_ctx* c, size_t bytes) {
    if (c->out_bytes >= bytes) return HS_OK;
    if (c->d_out) cudaFree(c->d_out);
    c->d_out = nullptr;
    c->out_bytes = 0;
    HS_CUDA(c, cudaMalloc(&c->d_out, bytes));
    c->out_bytes = bytes;
    return HS_OK;
}

float ev_ms(cudaEvent_t a, cudaEvent_t b) {
    float ms = 0.f;
    if (cudaEventElapsedTime(&ms, a, b) != cudaSuccess) ms = 0.f;
    return ms;
}

int do_upload(hs_ctx* c, const uint8_t* prev, size_t ps, size_t pis, const uint8_t* next, size_t ns,
              size_t nis) {
    if (!prev || !next) return fail(c, HS_ERR_INVALID_ARG, "null frame pointer");
    if (ps < (size_t)c->W * c->fes || ns < (size_t)c->W * c->fes)
        return fail(c, HS_ERR_INVALID_ARG, "row stride smaller than a row");
    if (c->B > 1 && (pis < ps * c->frows || nis < ns * c->frows))
        return fail(c, HS_ERR_INVALID_ARG, "image stride smaller than one image");
    if (c->d_ring[0]) { c->d_prev = c->d_ring[0]; c->d_next = c->d_ring[1]; }   // leave streaming mode
    cudaStream_t cs = c->xfer ? c->xfer : c->stream;
    const size_t rowb = (size_t)c->W * c->fes;
    // A pitched host->device copy whose rows are not whole 128-byte lines is row-bound, not byte-bound
    // (measured, profiles/r02q_memcpy_probe.txt: 60 us against 15 us for one 1242 x 375 frame).  Dense host
    // frames therefore cross PCIe as ONE flat copy into a staging buffer and are re-pitched on the device
    // (a device-to-device 2-D copy, a few microseconds).
    if (rowb != c->fpitch && ps == rowb && ns == rowb) {
        const size_t img = rowb * c->frows;
        if (!c->d_flat) HS_CUDA(c, cudaMalloc(&c->d_flat, 2 * img * c->B));
        const uint8_t* src[2] = {prev, next};
        const size_t is[2] = {pis, nis};
        uint8_t* dst[2] = {c->d_prev, c->d_next};
        for (int f = 0; f < 2; ++f) {
            uint8_t* flat = c->d_flat + (size_t)f * img * c->B;
            if (c->B == 1 || is[f] == img) {
                HS_CUDA(c, cudaMemcpyAsync(flat, src[f], img * c->B, cudaMemcpyHostToDevice, cs));
            } else {
                for (int b = 0; b < c->B; ++b)
                    HS_CUDA(c, cudaMemcpyAsync(flat + (size_t)b * img, src[f] + (size_t)b * is[f], img, cudaMemcpyHostToDevice, cs));
            }
            for (int b = 0; b < c->B; ++b)
                HS_CUDA(c, cudaMemcpy2DAsync(dst[f] + (size_t)b * c->fimg, c->fpitch, flat + (size_t)b * img, rowb, rowb, c->frows,
                                             cudaMemcpyDeviceToDevice, cs));
        }
        c->uploaded = true;
        c->prepared = false;
        return HS_OK;
    }
    for (int b = 0; b < c->B; ++b) {
        HS_CUDA(c, cudaMemcpy2DAsync(c->d_prev + (size_t)b * c->fimg, c->fpitch, prev + (size_t)b * pis, ps,
                                     (size_t)c->W * c->fes, c->frows, cudaMemcpyHostToDevice, cs));
        HS_CUDA(c, cudaMemcpy2DAsync(c->d_next + (size_t)b * c->fimg, c->fpitch, next + (size_t)b * nis, ns,
                                     (size_t)c->W * c->fes, c->frows, cudaMemcpyHostToDevice, cs));
    }
    c->uploaded = true;
    c->prepared = false;
    return HS_OK;
}

int f64_prepare(hs_ctx* c);
int f64_iterate(hs_ctx* c, int iters);
int f64_download(hs_ctx* c, void* u, size_t us, size_t uis, void* v, size_t vs, size_t vis, int dt);

int do_prepare(hs_ctx* c) {
    if (!c->uploaded) return fail(c, HS_ERR_STATE, "hs_prepare before frames were uploaded");
    if (c->f64) return f64_prepare(c);
    // u = v = 0 (hornSchunck.cpp:49-50): both plane pairs and the seam inbox are one allocation, one memset.
    // A linked row slab must not get here while a neighbour's previous hs_iterate is still running
    // (it stores into this arena): the group code orders that with events, separate processes barrier.
    HS_CUDA(c, cudaMemsetAsync(c->arena, 0, c->arena_bytes, c->stream));
    c->cur = 0;
    c->phase_count = 0;
    dim3 block(32, 8);
    dim3 grid((c->pitch / 4 + 31) / 32, (c->H + 7) / 8, c->B);
    const float a2 = (float)(c->alpha * c->alpha);
    if (c->textbook)
        hs::k_grad_coeff_tb<<<grid, block, 0, c->stream>>>(c->d_prev, c->d_next, c->fpitch, c->fimg, c->frows,
                                                           c->frow0, c->d_cpk, c->d_inv, c->geom());
    else
        hs::k_grad_coeff<<<grid, block, 0, c->stream>>>(c->d_prev, c->d_next, c->fpitch, c->fimg, c->frows,
                                                        c->frow0, c->d_cpk, c->d_inv, c->geom(), a2);
    HS_CUDA(c, cudaGetLastError());
    c->timing.launches += 1;
    c->prepared = true;
    return HS_OK;
}

int do_iterate(hs_ctx* c, int iters) {
    if (!c->prepared) return fail(c, HS_ERR_STATE, "hs_iterate before hs_prepare");
    if (iters < 0) return fail(c, HS_ERR_INVALID_ARG, "negative iteration count");
    if (c->f64) return f64_iterate(c, iters);
    int left = iters;
    while (left > 0) {
        int step = 1;
        cudaError_t e = cudaSuccess;
        if (c->kernel_id == 1) {
            // everything in one multi-phase launch, unless the caller must refresh halos between
            // fused launches (row slabs) or asked for per-launch behaviour
            // (a LINKED row slab exchanges its halos from inside the kernel: one launch as well)
            step = (c->multi_phase || c->linked) ? left : std::min(c->k, left);
            // (linked slabs keep the tile geometry of k even for a short call: seam flags are laid out by it)
            const int kk = c->linked ? c->k : std::min(c->k, step);
            tile_dispatch(c, [&](auto t) { e = decltype(t)::launch(c, kk, step, c->oy0, c->oy1, c->linked); });
            const int phases = (step + kk - 1) / kk;
            if (e == cudaSuccess) {
                if (phases & 1) c->cur ^= 1;
                c->phase_count += phases;
            }
        } else {
            dim3 block(32, 8);
            dim3 grid((c->W + 31) / 32, (c->oy1 - c->oy0 + 7) / 8, c->B);
            const float kf = 1.0f / (float)(c->w * c->w);
            if (c->textbook)
                hs::k_jacobi_generic_tb<<<grid, block, 0, c->stream>>>(c->d_uv[c->cur], c->d_uv[c->cur ^ 1], c->d_cpk,
                                                                       c->d_inv, c->geom(), (float)(c->alpha * c->alpha));
            else
                hs::k_jacobi_generic<<<grid, block, 0, c->stream>>>(c->d_uv[c->cur], c->d_uv[c->cur ^ 1], c->d_cpk, c->d_inv,
                                                                    c->geom(), c->w, c->a, kf);
            e = cudaGetLastError();
            c->cur ^= 1;
        }
        if (e != cudaSuccess) return fail(c, HS_ERR_CUDA, "sweep launch failed: %s", cudaGetErrorString(e));
        left -= step;
        c->timing.launches += 1;
    }
    return HS_OK;
}

int do_download(hs_ctx* c, void* u, size_t us, size_t uis, void* v, size_t vs, size_t vis, int dt) {
    if (!u || !v) return fail(c, HS_ERR_INVALID_ARG, "null output pointer");
    if (dt != HS_F32 && dt != HS_F64) return fail(c, HS_ERR_INVALID_ARG, "bad out_dtype");
    const size_t es = dt == HS_F64 ? 8 : 4;
    const int rows = c->oy1 - c->oy0;
    if (us < c->W * es || vs < c->W * es) return fail(c, HS_ERR_INVALID_ARG, "output row stride smaller than a row");
    if (c->f64) return f64_download(c, u, us, uis, v, vs, vis, dt);
    if (c->B > 1 && (uis < us * rows || vis < vs * rows))
        return fail(c, HS_ERR_INVALID_ARG, "output image stride smaller than one image");
    // the device keeps {u, v} interleaved; the caller gets two planes of out_dtype (K5 splits / widens)
    const long long n = c->plane * c->B;
    int rc = ensure_out(c, (size_t)n * 2 * es);
    if (rc) return rc;
    const char* su = static_cast<const char*>(c->d_out);
    const char* sv = su + (size_t)n * es;
    if (dt == HS_F64) {
        double* o = static_cast<double*>(c->d_out);
        hs::k_split<double><<<148 * 8, 256, 0, c->stream>>>(c->d_uv[c->cur], o, o + n, n);
    } else {
        float* o = static_cast<float*>(c->d_out);
        hs::k_split<float><<<148 * 8, 256, 0, c->stream>>>(c->d_uv[c->cur], o, o + n, n);
    }
    HS_CUDA(c, cudaGetLastError());
    c->timing.launches += 1;
    cudaStream_t cs = c->stream;
    if (c->xfer) {                 // hs_solve_async: the device->host copies leave on the transfer stream
        HS_CUDA(c, cudaEventRecord(c->ev_async[1], c->stream));
        HS_CUDA(c, cudaStreamWaitEvent(c->xfer, c->ev_async[1], 0));
        cs = c->xfer;
    }
    for (int b = 0; b < c->B; ++b) {
        const size_t off = ((size_t)b * c->plane + (size_t)c->oy0 * c->pitch) * es;
        HS_CUDA(c, cudaMemcpy2DAsync(static_cast<char*>(u) + (size_t)b * uis, us, su + off, (size_t)c->pitch * es,
                                     c->W * es, rows, cudaMemcpyDeviceToHost, cs));
        HS_CUDA(c, cudaMemcpy2DAsync(static_cast<char*>(v) + (size_t)b * vis, vs, sv + off, (size_t)c->pitch * es,
                                     c->W * es, rows, cudaMemcpyDeviceToHost, cs));
    }
    return HS_OK;
}

// ---- HS_PREC_F64 (hs_kernels_f64.cuh): reference arithmetic in fp64, frames of any depth -----------------
template <typename T>
void launch_grad_f64(hs_ctx* c) {
    dim3 block(32, 8), grid((c->W + 31) / 32, (c->H + 7) / 8, c->B);
    hs::k_grad_f64<T><<<grid, block, 0, c->stream>>>(reinterpret_cast<const T*>(c->d_prev), reinterpret_cast<const T*>(c->d_next),
                                                     c->fpitch, c->fimg, c->d64[0], c->d64[1], c->d64[2], c->d64[3], c->W, c->H,
                                                     c->pitch, c->plane, c->alpha * c->alpha);
}

int f64_prepare(hs_ctx* c) {
    const size_t npx = (size_t)c->plane * c->B;
    HS_CUDA(c, cudaMemsetAsync(c->d64[4], 0, npx * 4 * sizeof(double), c->stream));   // u = v = 0 (:49-50)
    c->cur = 0;
    switch (c->fdt) {
        case HS_FRAME_U8: launch_grad_f64<uint8_t>(c); break;
        case HS_FRAME_S8: launch_grad_f64<int8_t>(c); break;
        case HS_FRAME_U16: launch_grad_f64<uint16_t>(c); break;
        case HS_FRAME_S16: launch_grad_f64<int16_t>(c); break;
        case HS_FRAME_S32: launch_grad_f64<int32_t>(c); break;
        case HS_FRAME_F32: launch_grad_f64<float>(c); break;
        default: launch_grad_f64<double>(c); break;
    }
    HS_CUDA(c, cudaGetLastError());
    c->timing.launches += 1;
    c->prepared = true;
    return HS_OK;
}

int f64_iterate(hs_ctx* c, int iters) {
    dim3 block(32, 8), grid((c->W + 31) / 32, (c->H + 7) / 8, c->B);
    const double kf = 1.0 / (double)(c->w * c->w);                       // hornSchunck.cpp:53
    for (int i = 0; i < iters; ++i) {
        const int a = c->cur, b = c->cur ^ 1;
        hs::k_jacobi_f64<<<grid, block, 0, c->stream>>>(c->d64[4 + 2 * a], c->d64[5 + 2 * a], c->d64[4 + 2 * b], c->d64[5 + 2 * b],
                                                        c->d64[0], c->d64[1], c->d64[2], c->d64[3], c->W, c->H, c->pitch,
                                                        c->plane, c->w, c->a, kf);
        c->cur ^= 1;
        c->timing.launches += 1;
    }
    HS_CUDA(c, cudaGetLastError());
    return HS_OK;
}

int f64_copy_out(hs_ctx* c, const double* const* planes, int n, void* const* outs, size_t os, size_t ois, int dt) {
    const size_t es = dt == HS_F64 ? 8 : 4;
    const long long npx = c->plane * c->B;
    for (int i = 0; i < n; ++i) {
        const char* src = reinterpret_cast<const char*>(planes[i]);
        if (dt == HS_F32) {
            int rc = ensure_out(c, (size_t)npx * n * sizeof(float));
            if (rc) return rc;
            float* o = static_cast<float*>(c->d_out) + (size_t)i * npx;
            hs::k_narrow_f64<<<148 * 8, 256, 0, c->stream>>>(planes[i], o, npx);
            HS_CUDA(c, cudaGetLastError());
            c->timing.launches += 1;
            src = reinterpret_cast<const char*>(o);
        }
        for (int b = 0; b < c->B; ++b)
            HS_CUDA(c, cudaMemcpy2DAsync(static_cast<char*>(outs[i]) + (size_t)b * ois, os, src + (size_t)b * c->plane * es,
                                         (size_t)c->pitch * es, c->W * es, c->H, cudaMemcpyDeviceToHost, c->stream));
    }
    return HS_OK;
}

int f64_download(hs_ctx* c, void* u, size_t us, size_t uis, void* v, size_t vs, size_t vis, int dt) {
    if (us != vs || uis != vis) {
        const double* pu[1] = {c->d64[4 + 2 * c->cur]};
        const double* pv[1] = {c->d64[5 + 2 * c->cur]};
        void* ou[1] = {u}; void* ov[1] = {v};
        int rc = f64_copy_out(c, pu, 1, ou, us, uis, dt);
        if (rc) return rc;
        HS_CUDA(c, cudaStreamSynchronize(c->stream));                     // the narrow staging is reused
        return f64_copy_out(c, pv, 1, ov, vs, vis, dt);
    }
    const double* p[2] = {c->d64[4 + 2 * c->cur], c->d64[5 + 2 * c->cur]};
    void* o[2] = {u, v};
    return f64_copy_out(c, p, 2, o, us, uis, dt);
}

void group_destroy(hs_ctx* g);

void destroy_impl(hs_ctx* c) {
    if (!c) return;
    if (!c->kids.empty() || c->nccl) group_destroy(c);
    {
        DevGuard g(c->dev);
        if (c->stream) cudaStreamSynchronize(c->stream);
        cudaFree(c->d_ring[0] ? c->d_ring[0] : c->d_prev); cudaFree(c->d_ring[1] ? c->d_ring[1] : c->d_next);
        cudaFree(c->d_ring[2]); cudaFree(c->d_vout[0]); cudaFree(c->d_vout[1]);
        if (c->copy_stream) cudaStreamDestroy(c->copy_stream);
        if (c->ev_up) cudaEventDestroy(c->ev_up);
        for (auto& e : c->ev_k1) if (e) cudaEventDestroy(e);
        for (auto& e : c->ev_solved) if (e) cudaEventDestroy(e);
        for (auto& e : c->ev_async) if (e) cudaEventDestroy(e);
        if (c->ipc_up) cudaIpcCloseMemHandle(c->ipc_up);
        if (c->ipc_dn) cudaIpcCloseMemHandle(c->ipc_dn);
        cudaFree(c->arena);
        cudaFree(c->d_cpk); cudaFree(c->d_inv); cudaFree(c->d_done); cudaFree(c->d_bgr); cudaFree(c->d_flat); cudaFree(c->d_resid); cudaFree(c->d_out);
        for (auto& e : c->ev) if (e) cudaEventDestroy(e);
        if (c->own_stream && c->stream) cudaStreamDestroy(c->stream);
    }
    delete c;
}

}  // namespace

namespace {
int group_upload(hs_ctx* g, const uint8_t* prev, size_t ps, size_t pis, const uint8_t* next, size_t ns, size_t nis);
int group_prepare(hs_ctx* g);
int group_iterate(hs_ctx* g, int iters);
int group_download(hs_ctx* g, void* u, size_t us, size_t uis, void* v, size_t vs, size_t vis, int dt);
int group_sync(hs_ctx* g);
int group_solve_device(hs_ctx* g);
int group_solve(hs_ctx* g, const uint8_t* prev, size_t ps, size_t pis, const uint8_t* next, size_t ns, size_t nis,
                void* u, size_t us, size_t uis, void* v, size_t vs, size_t vis, int dt);
int create_single(const hs_config& cfg, hs_ctx** out);
int create_group(const hs_config& cfg, hs_ctx** out);
int create_slab_rank(const hs_config& cfg, hs_ctx** out);
}  // namespace

extern "C" {

int hs_version(void) { return HS_VERSION; }

int hs_default_temporal_k(const hs_config* cfg_in, int32_t num_sms) {
    if (!cfg_in || cfg_in->struct_size == 0 || cfg_in->struct_size > sizeof(hs_config) || num_sms < 1) return 0;
    hs_config cfg{};
    memcpy(&cfg, cfg_in, cfg_in->struct_size);
    if (cfg.width < 1 || cfg.height < 1 || cfg.window_size < 2 || cfg.window_size > 9) return 0;   // 0: no fused kernel
    const int a = cfg.window_size - cfg.window_size / 2 - 1, RL = a, RR = cfg.window_size - 1 - a;
    const int B = cfg.batch > 0 ? cfg.batch : 1;
    const bool seam = cfg.flags & (HS_FLAG_TOP_IS_SEAM | HS_FLAG_BOTTOM_IS_SEAM);
    int r0 = cfg.out_row_begin, r1 = cfg.out_row_end;
    if (r0 == 0 && r1 == 0) r1 = cfg.height;
    return choose_temporal_k(cfg.temporal_k, cfg.width, r1 - r0, r0 + cfg.global_row0, B, RL, RR, cfg.max_iterations, num_sms,
                             !(cfg.flags & HS_FLAG_SINGLE_PHASE), seam);
}

const char* hs_last_error(const hs_ctx* ctx) { return ctx ? ctx->err.c_str() : g_create_err.c_str(); }

int hs_create(const hs_config* cfg_in, hs_ctx** out) {
    if (out) *out = nullptr;
    if (!cfg_in || !out) return fail(nullptr, HS_ERR_INVALID_ARG, "null argument");
    if (cfg_in->struct_size == 0 || cfg_in->struct_size > sizeof(hs_config))
        return fail(nullptr, HS_ERR_INVALID_ARG, "hs_config.struct_size %u not understood (library knows %zu)",
                    cfg_in->struct_size, sizeof(hs_config));
    hs_config cfg{};
    memcpy(&cfg, cfg_in, cfg_in->struct_size);
    if (cfg.num_devices > 1) return create_group(cfg, out);
    if (cfg.slab_world > 1) return create_slab_rank(cfg, out);
    return create_single(cfg, out);
}

}  // extern "C"

namespace {

int create_single(const hs_config& cfg_in, hs_ctx** out) {
    hs_config cfg = cfg_in;
    if (cfg.width < 1 || cfg.height < 1) return fail(nullptr, HS_ERR_INVALID_ARG, "width and height must be >= 1");
    if (cfg.window_size < 1) return fail(nullptr, HS_ERR_INVALID_ARG, "window_size must be >= 1");
    if (cfg.max_iterations < 0) return fail(nullptr, HS_ERR_INVALID_ARG, "max_iterations must be >= 0");
    if (cfg.batch < 0) return fail(nullptr, HS_ERR_INVALID_ARG, "batch must be >= 1");
    if (cfg.temporal_k < 0) return fail(nullptr, HS_ERR_INVALID_ARG, "temporal_k must be >= 0");
    if (!(cfg.alpha == cfg.alpha)) return fail(nullptr, HS_ERR_INVALID_ARG, "alpha is NaN");

    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev == 0)
        return fail(nullptr, HS_ERR_CUDA, "no usable CUDA device (%s); this library has no CPU fallback",
                    e == cudaSuccess ? "device count 0" : cudaGetErrorString(e));
    int dev = cfg.device;
    if (dev < 0) {
        if (cudaGetDevice(&dev) != cudaSuccess) dev = 0;
    }
    if (dev >= ndev) return fail(nullptr, HS_ERR_INVALID_ARG, "device %d out of range (%d devices)", dev, ndev);

    hs_ctx* c = new (std::nothrow) hs_ctx();
    if (!c) return fail(nullptr, HS_ERR_OOM, "out of host memory");
    c->cfg = cfg;
    c->dev = dev;
    c->W = cfg.width; c->H = cfg.height; c->B = cfg.batch > 0 ? cfg.batch : 1;
    c->w = cfg.window_size; c->T = cfg.max_iterations; c->alpha = cfg.alpha;
    c->a = c->w - c->w / 2 - 1;               // hornSchunck.cpp:54
    c->RL = c->a; c->RR = c->w - 1 - c->a;
    c->oy0 = cfg.out_row_begin; c->oy1 = cfg.out_row_end;
    if (c->oy0 == 0 && c->oy1 == 0) c->oy1 = c->H;
    c->grow0 = cfg.global_row0;
    c->textbook = (cfg.flags & HS_FLAG_TEXTBOOK) != 0;
    c->top_seam = cfg.flags & HS_FLAG_TOP_IS_SEAM;
    c->bot_seam = cfg.flags & HS_FLAG_BOTTOM_IS_SEAM;
    c->pitch = round_up(c->W, 32);
    c->plane = (long long)c->pitch * c->H;
    c->frow0 = c->top_seam ? 1 : 0;
    c->frows = c->H + (c->top_seam ? 1 : 0) + (c->bot_seam ? 1 : 0);
    c->fpitch = (size_t)round_up(c->W, 128);
    c->fimg = c->fpitch * c->frows;

    auto bail = [&](int code) { g_create_err = c->err; destroy_impl(c); return code; };
    // precision / frame depth (hornSchunck.cpp:23-24 accepts any depth)
    static const int frame_bytes[7] = {1, 1, 2, 2, 4, 4, 8};
    if (cfg.frame_dtype < 0 || cfg.frame_dtype > HS_FRAME_F64) return bail(fail(c, HS_ERR_INVALID_ARG, "unknown frame_dtype %d", cfg.frame_dtype));
    if (cfg.precision != HS_PREC_F32 && cfg.precision != HS_PREC_F64) return bail(fail(c, HS_ERR_INVALID_ARG, "unknown precision %d", cfg.precision));
    if (cfg.precision == HS_PREC_F32 && cfg.frame_dtype != HS_FRAME_U8)
        return bail(fail(c, HS_ERR_UNSUPPORTED, "the fp32 path packs exact 8-bit gradients: frames of another depth need "
                                                "hs_config.precision = HS_PREC_F64 (the reference's own fp64 arithmetic)"));
    c->f64 = cfg.precision == HS_PREC_F64;
    c->fdt = cfg.frame_dtype;
    c->fes = frame_bytes[cfg.frame_dtype];
    if (c->f64) {
        if (c->top_seam || c->bot_seam || c->textbook || c->oy0 != 0 || c->oy1 != c->H)
            return bail(fail(c, HS_ERR_UNSUPPORTED, "HS_PREC_F64 contexts are whole-image, reference-arithmetic contexts"));
        c->fpitch = (size_t)round_up(c->W * c->fes, 128);
        c->fimg = c->fpitch * c->frows;
    }
    if (c->textbook && c->w != 3)
        return bail(fail(c, HS_ERR_INVALID_ARG, "HS_FLAG_TEXTBOOK is a 3x3 weighted average: window_size must be 3"));
    if (c->oy0 < 0 || c->oy1 > c->H || c->oy0 >= c->oy1)
        return bail(fail(c, HS_ERR_INVALID_ARG, "out rows [%d,%d) not inside [0,%d)", c->oy0, c->oy1, c->H));

    DevGuard guard(dev);
    if (!guard.ok) return bail(fail(c, HS_ERR_CUDA, "cudaSetDevice(%d) failed", dev));
    cudaDeviceProp prop;
    if ((e = cudaGetDeviceProperties(&prop, dev)) != cudaSuccess)
        return bail(fail(c, HS_ERR_CUDA, "cudaGetDeviceProperties: %s", cudaGetErrorString(e)));
    c->num_sms = prop.multiProcessorCount;
    c->use_pdl = env_int("HS_NO_PDL", 0) == 0;
    if (prop.major < 10)
        return bail(fail(c, HS_ERR_UNSUPPORTED, "device %d is sm_%d%d; this library is built for sm_100a (B200) only",
                         dev, prop.major, prop.minor));

#define HS_CREATE_CUDA(call)                                                                         \
    do {                                                                                             \
        cudaError_t e__ = (call);                                                                    \
        if (e__ != cudaSuccess)                                                                      \
            return bail(fail(c, e__ == cudaErrorMemoryAllocation ? HS_ERR_OOM : HS_ERR_CUDA,         \
                             "%s failed: %s", #call, cudaGetErrorString(e__)));                      \
    } while (0)

    if (cfg.stream) {
        c->stream = static_cast<cudaStream_t>(cfg.stream);
    } else {
        HS_CREATE_CUDA(cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking));
        c->own_stream = true;
    }
    for (auto& ev : c->ev) HS_CREATE_CUDA(cudaEventCreate(&ev));

    const size_t npx = (size_t)c->plane * c->B;
    HS_CREATE_CUDA(cudaMalloc(&c->d_prev, c->fimg * c->B));
    HS_CREATE_CUDA(cudaMalloc(&c->d_next, c->fimg * c->B));
    if (c->f64) {   // gx, gy, gt, den, u0, v0, u1, v1 in fp64: one allocation (arena), no fused kernel
        c->arena_bytes = npx * 8 * sizeof(double);
        HS_CREATE_CUDA(cudaMalloc(&c->arena, c->arena_bytes));
        for (int i = 0; i < 8; ++i) c->d64[i] = static_cast<double*>(c->arena) + (size_t)i * npx;
        c->k = 1; c->kernel_id = 2; c->multi_phase = false;
        c->timing.temporal_k = 1; c->timing.kernel_id = 2;
        *out = c;
        return HS_OK;
    }
    {   // the four flow planes and the seam inbox share ONE allocation: a single memset zeroes the state
        // (:49-50), and a single IPC handle exposes everything a neighbouring slab writes to
        const size_t pbytes = (npx * sizeof(float2) + 255) / 256 * 256;
        const size_t ibytes = (size_t)2 * hs::SEAM_JMAX * (c->W / 32 + 2) * sizeof(int);
        c->inbox_off = 2 * pbytes;
        c->arena_bytes = c->inbox_off + (ibytes + 255) / 256 * 256;
        HS_CREATE_CUDA(cudaMalloc(&c->arena, c->arena_bytes));
        char* base = static_cast<char*>(c->arena);
        for (int i = 0; i < 2 && !c->f64; ++i) {
            c->plane_off[i] = (size_t)i * pbytes;
            c->d_uv[i] = reinterpret_cast<float2*>(base + c->plane_off[i]);
        }
        c->d_inbox = reinterpret_cast<int*>(base + c->inbox_off);
        c->max_ctas = env_int("HS_MAX_CTAS", 0);
    }
    HS_CREATE_CUDA(cudaMalloc(&c->d_cpk, npx * sizeof(uint32_t)));
    HS_CREATE_CUDA(cudaMalloc(&c->d_inv, npx * sizeof(float)));

    // kernel selection: fused tile kernel for w in {2,3,4,5}, generic sweep otherwise
    const bool force_generic = (cfg.flags & HS_FLAG_FORCE_GENERIC) || env_int("HS_FORCE_GENERIC", 0);
    bool have_tile = false;
    int kmax = 1;
    if (!force_generic)
        have_tile = tile_dispatch(c, [&](auto t) {
            using TT = decltype(t);
            kmax = TT::max_k();
            e = TT::configure();
        });
    if (have_tile && e != cudaSuccess)
        return bail(fail(c, HS_ERR_CUDA, "cudaFuncSetAttribute(max dynamic smem): %s", cudaGetErrorString(e)));
    if (have_tile) {
        using TS0 = hs::TileShape<1, 1, TILE_R, TILE_NWARP>;
        const bool may_dataflow = !(cfg.flags & HS_FLAG_SINGLE_PHASE) && env_int("HS_SINGLE_PHASE", 0) == 0;
        const int k = choose_temporal_k(cfg.temporal_k > 0 ? cfg.temporal_k : env_int("HS_K", 0), c->W, c->oy1 - c->oy0,
                                        c->oy0 + c->grow0, c->B, c->RL, c->RR, cfg.max_iterations, c->num_sms,
                                        may_dataflow, c->top_seam || c->bot_seam);
        (void)kmax;
        c->k = k;
        c->kernel_id = 1;
        int rc;
        for (int i = 0; i < 2; ++i) {
            // {u, v} interleaved: a float plane 2*W wide, box 2*SX floats (256 = the TMA box limit)
            if ((rc = make_map_uv(c, &c->tm_uv[i], c->d_uv[i], TS0::SX, TS0::SY))) return bail(rc);
        }
        if ((rc = make_map(c, &c->tm_cpk, c->d_cpk, CU_TENSOR_MAP_DATA_TYPE_UINT32, 4, TS0::SX, TS0::SY))) return bail(rc);
        if ((rc = make_map(c, &c->tm_inv, c->d_inv, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, TS0::SX, TS0::SY))) return bail(rc);
    } else {
        c->k = 1;
        c->kernel_id = 0;
    }
    // a seam needs k sweeps' worth of halo rows between the buffer edge and the output rows
    if ((c->top_seam && c->oy0 < c->RL * c->k) || (c->bot_seam && c->H - c->oy1 < c->RR * c->k))
        return bail(fail(c, HS_ERR_INVALID_ARG,
                         "row slab needs %d halo rows above and %d below its output rows for k=%d (have %d / %d)",
                         c->RL * c->k, c->RR * c->k, c->k, c->oy0, c->H - c->oy1));
    if (c->kernel_id == 1) {
        size_t cap = 0;
        tile_dispatch(c, [&](auto t) { cap = decltype(t)::max_tiles(c); });
        HS_CREATE_CUDA(cudaMalloc(&c->d_done, std::max<size_t>(cap, 1) * sizeof(int)));
        c->done_cap = cap;
        int coop = 0;
        cudaDeviceGetAttribute(&coop, cudaDevAttrCooperativeLaunch, dev);
        c->multi_phase = coop && !c->top_seam && !c->bot_seam && !(cfg.flags & HS_FLAG_SINGLE_PHASE) &&
                         env_int("HS_SINGLE_PHASE", 0) == 0;
        // The dataflow launch pays off once a phase is more than about one round of tiles per SM: a
        // tile's inputs are then published well before its turn.  With one round or less every
        // phase is a dependent round and the publish/poll latency is exposed; plain launches
        // chained by PDL are faster there (1242x375, 121 tiles: 377 vs 275 Gpix-it/s; 1280x720,
        // 198..276 tiles: 354..506 vs 366..543, profiles/r01j_multi_phase_720p.txt).
        size_t tiles_k = 0;
        tile_dispatch(c, [&](auto t) { tiles_k = decltype(t)::tiles_for(c, c->k); });
        if (tiles_k * 4 < (size_t)c->num_sms * 5 && env_int("HS_MULTI_PHASE", 0) == 0) c->multi_phase = false;
        // halos wider than one tile pitch (an explicit, very deep temporal_k): the 3 x 3 dependency
        // rule of the dataflow launch would not cover them -> chained single-phase launches
        bool df_ok = true;
        tile_dispatch(c, [&](auto t) { df_ok = decltype(t)::dataflow_ok(c->k, c->oy0 + c->grow0); });
        if (!df_ok) c->multi_phase = false;
    }
    c->timing.temporal_k = c->k;
    c->timing.kernel_id = c->kernel_id;
    *out = c;
    return HS_OK;
#undef HS_CREATE_CUDA
}

}  // namespace

#include "hs_multi.inl"

extern "C" {

void hs_destroy(hs_ctx* ctx) { destroy_impl(ctx); }

#define HS_NO_F64(c, name)                                                                                \
    if ((c)->f64) return fail((c), HS_ERR_UNSUPPORTED, name " is not available on an HS_PREC_F64 context")

// an hs_solve_async call is in flight: its buffers and planes must not be touched until hs_solve_wait
#define HS_NO_ASYNC(c, name)                                                                              \
    if ((c)->async_pending)                                                                               \
        return fail((c), HS_ERR_STATE, name ": an hs_solve_async call is in flight on this context (hs_solve_wait first)")

#define HS_NO_GROUP(c, name)                                                                              \
    if (!(c)->kids.empty())                                                                               \
        return fail((c), HS_ERR_UNSUPPORTED, name " is not available on a multi-device context (use hs_solve, "   \
                                             "or hs_upload / hs_solve_device / hs_download / hs_sync)")

int hs_upload(hs_ctx* c, const uint8_t* prev, size_t ps, size_t pis, const uint8_t* next, size_t ns, size_t nis) {
    if (!c) return HS_ERR_INVALID_ARG;
    HS_NO_ASYNC(c, "hs_upload");
    if (!c->kids.empty()) return group_upload(c, prev, ps, pis, next, ns, nis);
    DevGuard g(c->dev);
    return do_upload(c, prev, ps, pis, next, ns, nis);
}

int hs_prepare(hs_ctx* c) {
    if (!c) return HS_ERR_INVALID_ARG;
    HS_NO_ASYNC(c, "hs_prepare");
    if (!c->kids.empty()) return group_prepare(c);
    DevGuard g(c->dev);
    return do_prepare(c);
}

int hs_iterate(hs_ctx* c, int iterations) {
    if (!c) return HS_ERR_INVALID_ARG;
    HS_NO_ASYNC(c, "hs_iterate");
    if (!c->kids.empty()) {
        if (!c->prepared) return fail(c, HS_ERR_STATE, "hs_iterate before hs_prepare");
        return group_iterate(c, iterations);
    }
    if ((c->top_seam || c->bot_seam) && !c->linked && iterations > c->k)
        return fail(c, HS_ERR_UNSUPPORTED, "row-slab context that is not linked to its neighbours (hs_slab_connect): at most "
                                           "temporal_k=%d sweeps between halo refreshes", c->k);
    DevGuard g(c->dev);
    return do_iterate(c, iterations);
}

int hs_iterate_rows(hs_ctx* c, int sweeps, int row_begin, int row_end, int flip) {
    if (!c) return HS_ERR_INVALID_ARG;
    HS_NO_GROUP(c, "hs_iterate_rows");
    HS_NO_F64(c, "hs_iterate_rows");
    if (c->linked) return fail(c, HS_ERR_STATE, "hs_iterate_rows on a connected row slab: hs_iterate exchanges the halos itself");
    if (!c->prepared) return fail(c, HS_ERR_STATE, "hs_iterate_rows before hs_prepare");
    if (c->kernel_id != 1) return fail(c, HS_ERR_UNSUPPORTED, "hs_iterate_rows needs the fused kernel (window 2..5)");
    if (sweeps < 1 || sweeps > c->k) return fail(c, HS_ERR_INVALID_ARG, "sweeps must be in [1, temporal_k=%d]", c->k);
    // any rows of the buffer may be produced: a deep-halo row slab also advances (part of) its halo
    if (row_begin < 0 || row_end > c->H || row_begin >= row_end)
        return fail(c, HS_ERR_INVALID_ARG, "rows [%d,%d) not inside the buffer rows [0,%d)", row_begin, row_end, c->H);
    DevGuard g(c->dev);
    cudaError_t e = cudaSuccess;
    tile_dispatch(c, [&](auto t) { e = decltype(t)::launch(c, sweeps, sweeps, row_begin, row_end); });
    if (e != cudaSuccess) return fail(c, HS_ERR_CUDA, "sweep launch failed: %s", cudaGetErrorString(e));
    c->timing.launches += 1;
    if (flip) c->cur ^= 1;
    return HS_OK;
}

int hs_iterate_until(hs_ctx* c, int max_sweeps, double tolerance, int check_every, int* sweeps_done, double* residual) {
    if (!c) return HS_ERR_INVALID_ARG;
    HS_NO_GROUP(c, "hs_iterate_until");
    HS_NO_F64(c, "hs_iterate_until");
    if (!c->prepared) return fail(c, HS_ERR_STATE, "hs_iterate_until before hs_prepare");
    if (max_sweeps < 0 || check_every < 1 || !(tolerance >= 0)) return fail(c, HS_ERR_INVALID_ARG, "bad early-exit arguments");
    if (c->top_seam || c->bot_seam) return fail(c, HS_ERR_UNSUPPORTED, "hs_iterate_until needs a whole-image context");
    DevGuard g(c->dev);
    if (!c->d_resid) HS_CUDA(c, cudaMalloc(&c->d_resid, sizeof(unsigned int)));
    int done = 0, rc;
    float r = __builtin_inff();
    while (done < max_sweeps) {
        const int chunk = std::min(check_every, max_sweeps - done);
        // the last sweep of a chunk runs alone, so that the two plane pairs hold consecutive iterates
        if (chunk > 1 && (rc = do_iterate(c, chunk - 1))) return rc;
        if ((rc = do_iterate(c, 1))) return rc;
        done += chunk;
        HS_CUDA(c, cudaMemsetAsync(c->d_resid, 0, sizeof(unsigned int), c->stream));
        hs::k_max_abs_diff<<<148 * 4, 256, 0, c->stream>>>(c->d_uv[c->cur ^ 1], c->d_uv[c->cur], c->geom(), c->B, c->d_resid);
        HS_CUDA(c, cudaGetLastError());
        c->timing.launches += 1;
        unsigned int bits = 0;
        HS_CUDA(c, cudaMemcpyAsync(&bits, c->d_resid, sizeof bits, cudaMemcpyDeviceToHost, c->stream));
        HS_CUDA(c, cudaStreamSynchronize(c->stream));
        memcpy(&r, &bits, sizeof r);
        if (r <= (float)tolerance) break;
    }
    if (sweeps_done) *sweeps_done = done;
    if (residual) *residual = (double)r;
    return HS_OK;
}

int hs_solve_device(hs_ctx* c) {
    if (!c) return HS_ERR_INVALID_ARG;
    HS_NO_ASYNC(c, "hs_solve_device");
    if (!c->kids.empty()) return group_solve_device(c);
    if (c->top_seam || c->bot_seam)
        return fail(c, HS_ERR_UNSUPPORTED, "hs_solve_device on a row-slab context: halos must be refreshed every "
                                           "temporal_k sweeps (use hs_iterate / hs_iterate_rows or hs_slab_*)");
    DevGuard g(c->dev);
    c->timing.launches = 0;
    HS_CUDA(c, cudaEventRecord(c->ev[1], c->stream));
    int rc = do_prepare(c);
    if (rc) return rc;
    HS_CUDA(c, cudaEventRecord(c->ev[2], c->stream));
    if ((rc = do_iterate(c, c->T))) return rc;
    HS_CUDA(c, cudaEventRecord(c->ev[3], c->stream));
    c->timing_pending = true;   // resolved by the next hs_sync
    return HS_OK;
}

int hs_download(hs_ctx* c, void* u, size_t us, size_t uis, void* v, size_t vs, size_t vis, int dt) {
    if (!c) return HS_ERR_INVALID_ARG;
    HS_NO_ASYNC(c, "hs_download");
    if (!c->kids.empty()) {
        int rc = group_download(c, u, us, uis, v, vs, vis, dt);
        return rc ? rc : group_sync(c);
    }
    DevGuard g(c->dev);
    int rc = do_download(c, u, us, uis, v, vs, vis, dt);
    if (rc) return rc;
    HS_CUDA(c, cudaStreamSynchronize(c->stream));
    return HS_OK;
}

int hs_sync(hs_ctx* c) {
    if (!c) return HS_ERR_INVALID_ARG;
    if (!c->kids.empty()) return group_sync(c);
    DevGuard g(c->dev);
    HS_CUDA(c, cudaStreamSynchronize(c->stream));
    if (c->timing_pending) {
        c->timing_pending = false;
        c->timing.h2d_ms = c->timing.d2h_ms = 0.f;
        c->timing.prepare_ms = ev_ms(c->ev[1], c->ev[2]);
        c->timing.iterate_ms = ev_ms(c->ev[2], c->ev[3]);
        c->timing.total_ms = ev_ms(c->ev[1], c->ev[3]);
    }
    return HS_OK;
}

int hs_solve(hs_ctx* c, const uint8_t* prev, size_t ps, size_t pis, const uint8_t* next, size_t ns, size_t nis,
             void* u, size_t us, size_t uis, void* v, size_t vs, size_t vis, int dt) {
    if (!c) return HS_ERR_INVALID_ARG;
    HS_NO_ASYNC(c, "hs_solve");
    if (!c->kids.empty()) return group_solve(c, prev, ps, pis, next, ns, nis, u, us, uis, v, vs, vis, dt);
    if (c->top_seam || c->bot_seam)
        return fail(c, HS_ERR_UNSUPPORTED, "hs_solve on a row-slab context: halos must be refreshed every temporal_k "
                                           "sweeps (use hs_iterate / hs_iterate_rows or hs_slab_*)");
    DevGuard g(c->dev);
    c->timing.launches = 0;
    int rc;
    HS_CUDA(c, cudaEventRecord(c->ev[0], c->stream));
    if ((rc = do_upload(c, prev, ps, pis, next, ns, nis))) return rc;
    HS_CUDA(c, cudaEventRecord(c->ev[1], c->stream));
    if ((rc = do_prepare(c))) return rc;
    HS_CUDA(c, cudaEventRecord(c->ev[2], c->stream));
    if ((rc = do_iterate(c, c->T))) return rc;
    HS_CUDA(c, cudaEventRecord(c->ev[3], c->stream));
    if ((rc = do_download(c, u, us, uis, v, vs, vis, dt))) return rc;
    HS_CUDA(c, cudaEventRecord(c->ev[4], c->stream));
    HS_CUDA(c, cudaStreamSynchronize(c->stream));
    c->timing.h2d_ms = ev_ms(c->ev[0], c->ev[1]);
    c->timing.prepare_ms = ev_ms(c->ev[1], c->ev[2]);
    c->timing.iterate_ms = ev_ms(c->ev[2], c->ev[3]);
    c->timing.d2h_ms = ev_ms(c->ev[3], c->ev[4]);
    c->timing.total_ms = ev_ms(c->ev[0], c->ev[4]);
    return HS_OK;
}

// hs_solve without the wait.  Host<->device copies run on a transfer stream of the context, the kernels on its
// compute stream (hs_config.stream), ordered by events: contexts that share ONE compute stream therefore run their
// kernels back to back in call order while the copies of one overlap the sweeps of the other.
int hs_solve_async(hs_ctx* c, const uint8_t* prev, size_t ps, size_t pis, const uint8_t* next, size_t ns, size_t nis,
                   void* u, size_t us, size_t uis, void* v, size_t vs, size_t vis, int dt) {
    if (!c) return HS_ERR_INVALID_ARG;
    HS_NO_GROUP(c, "hs_solve_async");
    HS_NO_F64(c, "hs_solve_async");
    if (c->top_seam || c->bot_seam) return fail(c, HS_ERR_UNSUPPORTED, "hs_solve_async needs a whole-image context");
    if (c->async_pending) return fail(c, HS_ERR_STATE, "hs_solve_async: the previous call was not waited for (hs_solve_wait)");
    DevGuard g(c->dev);
    if (!c->copy_stream) HS_CUDA(c, cudaStreamCreateWithFlags(&c->copy_stream, cudaStreamNonBlocking));
    for (auto& e : c->ev_async)
        if (!e) HS_CUDA(c, cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
    c->timing.launches = 0;
    // the transfer stream must not overwrite the frames / the output staging while an earlier solve of this
    // context still uses them: hs_solve_wait (or hs_sync) has returned, so there is none
    c->xfer = c->copy_stream;
    int rc = do_upload(c, prev, ps, pis, next, ns, nis);
    if (!rc) {
        cudaError_t e = cudaEventRecord(c->ev_async[0], c->xfer);
        if (e == cudaSuccess) e = cudaStreamWaitEvent(c->stream, c->ev_async[0], 0);
        if (e == cudaSuccess) e = cudaEventRecord(c->ev[1], c->stream);
        if (e != cudaSuccess) rc = fail(c, HS_ERR_CUDA, "hs_solve_async: %s", cudaGetErrorString(e));
    }
    if (!rc) rc = do_prepare(c);
    if (!rc && cudaEventRecord(c->ev[2], c->stream) != cudaSuccess) rc = fail(c, HS_ERR_CUDA, "cudaEventRecord failed");
    if (!rc) rc = do_iterate(c, c->T);
    if (!rc && cudaEventRecord(c->ev[3], c->stream) != cudaSuccess) rc = fail(c, HS_ERR_CUDA, "cudaEventRecord failed");
    if (!rc) rc = do_download(c, u, us, uis, v, vs, vis, dt);
    c->xfer = nullptr;
    if (rc) return rc;
    c->async_pending = true;
    return HS_OK;
}

int hs_solve_wait(hs_ctx* c) {
    if (!c) return HS_ERR_INVALID_ARG;
    if (!c->async_pending) return HS_OK;
    DevGuard g(c->dev);
    HS_CUDA(c, cudaStreamSynchronize(c->copy_stream));    // the last copy waits for everything before it
    c->async_pending = false;
    c->timing.h2d_ms = c->timing.d2h_ms = 0.f;
    c->timing.prepare_ms = ev_ms(c->ev[1], c->ev[2]);
    c->timing.iterate_ms = ev_ms(c->ev[2], c->ev[3]);
    c->timing.total_ms = ev_ms(c->ev[1], c->ev[3]);
    return HS_OK;
}

int hs_solve_bgr(hs_ctx* c, const uint8_t* prev, size_t ps, const uint8_t* next, size_t ns,
                 void* u, size_t us, void* v, size_t vs, int dt) {
    if (!c) return HS_ERR_INVALID_ARG;
    HS_NO_ASYNC(c, "hs_solve_bgr");
    HS_NO_GROUP(c, "hs_solve_bgr");
    HS_NO_F64(c, "hs_solve_bgr");
    if (c->B != 1 || c->top_seam || c->bot_seam)
        return fail(c, HS_ERR_UNSUPPORTED, "hs_solve_bgr needs a whole-image, batch == 1 context");
    if (!prev || !next) return fail(c, HS_ERR_INVALID_ARG, "null frame pointer");
    if (ps < (size_t)c->W * 3 || ns < (size_t)c->W * 3) return fail(c, HS_ERR_INVALID_ARG, "row stride smaller than 3*width");
    DevGuard g(c->dev);
    c->timing.launches = 0;
    if (!c->d_bgr) {
        c->bgr_pitch = (size_t)round_up(c->W * 3, 128);
        HS_CUDA(c, cudaMalloc(&c->d_bgr, c->bgr_pitch * c->H * 2));
    }
    int rc;
    HS_CUDA(c, cudaEventRecord(c->ev[0], c->stream));
    uint8_t* d0 = c->d_bgr;
    uint8_t* d1 = c->d_bgr + c->bgr_pitch * c->H;
    HS_CUDA(c, cudaMemcpy2DAsync(d0, c->bgr_pitch, prev, ps, (size_t)c->W * 3, c->H, cudaMemcpyHostToDevice, c->stream));
    HS_CUDA(c, cudaMemcpy2DAsync(d1, c->bgr_pitch, next, ns, (size_t)c->W * 3, c->H, cudaMemcpyHostToDevice, c->stream));
    HS_CUDA(c, cudaEventRecord(c->ev[1], c->stream));
    dim3 grid(((c->W + 3) / 4 + 255) / 256, c->H);
    hs::k_bgr2gray<<<grid, 256, 0, c->stream>>>(d0, c->bgr_pitch, c->d_prev, c->fpitch, c->W, c->H);   // main.cpp:11-26
    hs::k_bgr2gray<<<grid, 256, 0, c->stream>>>(d1, c->bgr_pitch, c->d_next, c->fpitch, c->W, c->H);
    HS_CUDA(c, cudaGetLastError());
    c->timing.launches += 2;
    c->uploaded = true;
    if ((rc = do_prepare(c))) return rc;
    HS_CUDA(c, cudaEventRecord(c->ev[2], c->stream));
    if ((rc = do_iterate(c, c->T))) return rc;
    HS_CUDA(c, cudaEventRecord(c->ev[3], c->stream));
    if ((rc = do_download(c, u, us, 0, v, vs, 0, dt))) return rc;
    HS_CUDA(c, cudaEventRecord(c->ev[4], c->stream));
    HS_CUDA(c, cudaStreamSynchronize(c->stream));
    c->timing.h2d_ms = ev_ms(c->ev[0], c->ev[1]);
    c->timing.prepare_ms = ev_ms(c->ev[1], c->ev[2]);
    c->timing.iterate_ms = ev_ms(c->ev[2], c->ev[3]);
    c->timing.d2h_ms = ev_ms(c->ev[3], c->ev[4]);
    c->timing.total_ms = ev_ms(c->ev[0], c->ev[4]);
    return HS_OK;
}

int hs_gradients(hs_ctx* c, const uint8_t* prev, size_t ps, const uint8_t* next, size_t ns, void* gx, void* gy,
                 void* gt, size_t os, int dt) {
    if (!c) return HS_ERR_INVALID_ARG;
    HS_NO_ASYNC(c, "hs_gradients");
    HS_NO_GROUP(c, "hs_gradients");
    if (c->B != 1) return fail(c, HS_ERR_UNSUPPORTED, "hs_gradients needs a batch == 1 context");
    if (!gx || !gy || !gt) return fail(c, HS_ERR_INVALID_ARG, "null output pointer");
    if (dt != HS_F32 && dt != HS_F64) return fail(c, HS_ERR_INVALID_ARG, "bad out_dtype");
    const size_t es = dt == HS_F64 ? 8 : 4;
    if (os < c->W * es) return fail(c, HS_ERR_INVALID_ARG, "output row stride smaller than a row");
    DevGuard g(c->dev);
    c->timing.launches = 0;
    int rc;
    if ((rc = do_upload(c, prev, ps, 0, next, ns, 0))) return rc;
    if ((rc = do_prepare(c))) return rc;
    if (c->f64) {
        const double* p[3] = {c->d64[0], c->d64[1], c->d64[2]};
        void* o3[3] = {gx, gy, gt};
        if ((rc = f64_copy_out(c, p, 3, o3, os, 0, dt))) return rc;
        HS_CUDA(c, cudaStreamSynchronize(c->stream));
        return HS_OK;
    }
    const long long n = c->plane;
    if ((rc = ensure_out(c, (size_t)n * 3 * es))) return rc;
    char* o = static_cast<char*>(c->d_out);
    if (dt == HS_F64) {
        double* d = reinterpret_cast<double*>(o);
        if (c->textbook) hs::k_unpack_grad_tb<double><<<148 * 8, 256, 0, c->stream>>>(c->d_cpk, c->d_inv, d, d + n, d + 2 * n, n);
        else hs::k_unpack_grad<double><<<148 * 8, 256, 0, c->stream>>>(c->d_cpk, d, d + n, d + 2 * n, n);
    } else {
        float* d = reinterpret_cast<float*>(o);
        if (c->textbook) hs::k_unpack_grad_tb<float><<<148 * 8, 256, 0, c->stream>>>(c->d_cpk, c->d_inv, d, d + n, d + 2 * n, n);
        else hs::k_unpack_grad<float><<<148 * 8, 256, 0, c->stream>>>(c->d_cpk, d, d + n, d + 2 * n, n);
    }
    HS_CUDA(c, cudaGetLastError());
    c->timing.launches += 1;
    void* outs[3] = {gx, gy, gt};
    for (int i = 0; i < 3; ++i)
        HS_CUDA(c, cudaMemcpy2DAsync(outs[i], os, o + ((size_t)i * n + (size_t)c->oy0 * c->pitch) * es,
                                     (size_t)c->pitch * es, c->W * es, c->oy1 - c->oy0, cudaMemcpyDeviceToHost,
                                     c->stream));
    HS_CUDA(c, cudaStreamSynchronize(c->stream));
    return HS_OK;
}

int hs_sample_grid(hs_ctx* c, int delta, double* u, double* v, int* ny_out, int* nx_out) {
    if (!c || delta < 1) return HS_ERR_INVALID_ARG;
    HS_NO_GROUP(c, "hs_sample_grid");
    HS_NO_F64(c, "hs_sample_grid");
    if (c->B != 1) return fail(c, HS_ERR_UNSUPPORTED, "hs_sample_grid needs a batch == 1 context");
    const int rows = c->oy1 - c->oy0;
    const int ny = (rows + delta - 1) / delta, nx = (c->W + delta - 1) / delta;
    if (ny_out) *ny_out = ny;
    if (nx_out) *nx_out = nx;
    if (!u || !v) return HS_OK;                           // size query
    DevGuard g(c->dev);
    const size_t n = (size_t)ny * nx;
    int rc = ensure_out(c, n * 2 * sizeof(double));
    if (rc) return rc;
    double* o = static_cast<double*>(c->d_out);
    hs::k_sample_grid<<<(unsigned)((n + 255) / 256), 256, 0, c->stream>>>(c->d_uv[c->cur], o, o + n,
                                                                       c->pitch, c->oy0, delta, ny, nx);
    HS_CUDA(c, cudaGetLastError());
    HS_CUDA(c, cudaMemcpyAsync(u, o, n * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
    HS_CUDA(c, cudaMemcpyAsync(v, o + n, n * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
    HS_CUDA(c, cudaStreamSynchronize(c->stream));
    return HS_OK;
}

int hs_get_device_view(hs_ctx* c, hs_device_view* o) {
    if (!c || !o) return HS_ERR_INVALID_ARG;
    HS_NO_GROUP(c, "hs_get_device_view");
    HS_NO_F64(c, "hs_get_device_view");
    o->prev = c->d_prev; o->next = c->d_next;
    o->frame_pitch = c->fpitch; o->frame_pair_stride = c->fimg;
    o->frame_rows = c->frows; o->frame_row0 = c->frow0;
    o->uv = reinterpret_cast<float*>(c->d_uv[c->cur]); o->reserved = nullptr;
    o->flow_pitch = (size_t)c->pitch * sizeof(float2);
    o->flow_pair_stride = (size_t)c->plane * sizeof(float2);
    o->width = c->W; o->height = c->H; o->batch = c->B;
    o->halo_rows_top = c->RL * c->k;
    o->halo_rows_bottom = c->RR * c->k;
    return HS_OK;
}

int hs_get_timing(const hs_ctx* c, hs_timing* o) {
    if (!c || !o) return HS_ERR_INVALID_ARG;
    *o = c->timing;
    return HS_OK;
}

#ifdef HS_TILE_PROFILE
// debug build only: read (and clear) the per-CTA tile phase counters
int hs_debug_tile_profile(long long* out, int clear) {
    cudaError_t e = cudaMemcpyFromSymbol(out, hs::g_tile_prof, sizeof(long long) * 256 * 8);
    if (e == cudaSuccess && clear) {
        static long long zeros[256 * 8];
        e = cudaMemcpyToSymbol(hs::g_tile_prof, zeros, sizeof zeros);
    }
    return e == cudaSuccess ? HS_OK : HS_ERR_CUDA;
}
#endif

int hs_host_alloc(void** ptr, size_t bytes) {
    if (!ptr) return HS_ERR_INVALID_ARG;
    *ptr = nullptr;
    cudaError_t e = cudaMallocHost(ptr, bytes);
    if (e != cudaSuccess) { g_create_err = cudaGetErrorString(e); return e == cudaErrorMemoryAllocation ? HS_ERR_OOM : HS_ERR_CUDA; }
    return HS_OK;
}

int hs_host_free(void* ptr) {
    cudaError_t e = cudaFreeHost(ptr);
    if (e != cudaSuccess) { g_create_err = cudaGetErrorString(e); return HS_ERR_CUDA; }
    return HS_OK;
}

}  // extern "C"

// ---- streaming front-end -----------------------------------------------------------------------
namespace {

int video_setup(hs_ctx* c, int dt) {
    if (c->B != 1 || c->top_seam || c->bot_seam)
        return fail(c, HS_ERR_UNSUPPORTED, "hs_video_* needs a whole-image, batch == 1 context");
    if (dt != HS_F32 && dt != HS_F64) return fail(c, HS_ERR_INVALID_ARG, "bad out_dtype");
    if (!c->ev_up) {
        if (!c->copy_stream) HS_CUDA(c, cudaStreamCreateWithFlags(&c->copy_stream, cudaStreamNonBlocking));
        HS_CUDA(c, cudaEventCreateWithFlags(&c->ev_up, cudaEventDisableTiming));
        for (auto& e : c->ev_k1) HS_CUDA(c, cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
        for (auto& e : c->ev_solved) HS_CUDA(c, cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
        c->d_ring[0] = c->d_prev;
        c->d_ring[1] = c->d_next;
        HS_CUDA(c, cudaMalloc(&c->d_ring[2], c->fimg));
        for (auto& o : c->d_vout) HS_CUDA(c, cudaMalloc(&o, (size_t)c->plane * 2 * sizeof(double)));
    }
    if (c->vid_dtype >= 0 && c->vid_dtype != dt && (c->vid_pending >= 0))
        return fail(c, HS_ERR_INVALID_ARG, "out_dtype changed while a pair is pending");
    c->vid_dtype = dt;
    return HS_OK;
}

// copy the flow of pair p from its staging slot to the host and wait for exactly that copy
int video_fetch(hs_ctx* c, int p, void* u, size_t us, void* v, size_t vs) {
    if (!u || !v) return fail(c, HS_ERR_INVALID_ARG, "null output pointer");
    const size_t es = c->vid_dtype == HS_F64 ? 8 : 4;
    if (us < c->W * es || vs < c->W * es) return fail(c, HS_ERR_INVALID_ARG, "output row stride smaller than a row");
    const char* src = static_cast<const char*>(c->d_vout[p & 1]);
    HS_CUDA(c, cudaStreamWaitEvent(c->copy_stream, c->ev_solved[p & 1], 0));
    HS_CUDA(c, cudaMemcpy2DAsync(u, us, src, (size_t)c->pitch * es, c->W * es, c->H, cudaMemcpyDeviceToHost, c->copy_stream));
    HS_CUDA(c, cudaMemcpy2DAsync(v, vs, src + (size_t)c->plane * es, (size_t)c->pitch * es, c->W * es, c->H,
                                 cudaMemcpyDeviceToHost, c->copy_stream));
    HS_CUDA(c, cudaStreamSynchronize(c->copy_stream));
    return HS_OK;
}

}  // namespace

extern "C" int hs_video_reset(hs_ctx* c) {
    if (!c) return HS_ERR_INVALID_ARG;
    DevGuard g(c->dev);
    HS_CUDA(c, cudaStreamSynchronize(c->stream));
    if (c->copy_stream) HS_CUDA(c, cudaStreamSynchronize(c->copy_stream));
    c->vid_frames = 0;
    c->vid_pending = -1;
    c->vid_dtype = -1;
    return HS_OK;
}

extern "C" int hs_video_push(hs_ctx* c, const uint8_t* frame, size_t stride, void* u, size_t us, void* v, size_t vs,
                             int dt, int* pair_index) {
    if (!c || !pair_index) return HS_ERR_INVALID_ARG;
    *pair_index = -1;
    HS_NO_ASYNC(c, "hs_video_push");
    HS_NO_GROUP(c, "hs_video_push");
    HS_NO_F64(c, "hs_video_push");
    if (!frame) return fail(c, HS_ERR_INVALID_ARG, "null frame pointer");
    if (stride < (size_t)c->W) return fail(c, HS_ERR_INVALID_ARG, "row stride smaller than width");
    DevGuard g(c->dev);
    int rc = video_setup(c, dt);
    if (rc) return rc;
    const int n = c->vid_frames;
    // a flow is due whenever a pair is pending: check its destination BEFORE anything is queued, so a
    // bad argument cannot cost the caller the pending pair
    if (n >= 1 && c->vid_pending >= 0) {
        const size_t es = dt == HS_F64 ? 8 : 4;
        if (!u || !v) return fail(c, HS_ERR_INVALID_ARG, "null output pointer (the flow of pair %d is due)", c->vid_pending);
        if (us < c->W * es || vs < c->W * es) return fail(c, HS_ERR_INVALID_ARG, "output row stride smaller than a row");
    }
    // 1. upload frame n into ring slot n % 3 (last read by the gradient kernel of pair n-3)
    if (n >= 3) HS_CUDA(c, cudaStreamWaitEvent(c->copy_stream, c->ev_k1[(n - 3) % 3], 0));
    HS_CUDA(c, cudaMemcpy2DAsync(c->d_ring[n % 3], c->fpitch, frame, stride, c->W, c->H, cudaMemcpyHostToDevice,
                                 c->copy_stream));
    HS_CUDA(c, cudaEventRecord(c->ev_up, c->copy_stream));
    if (n == 0) { c->vid_frames = 1; return HS_OK; }
    // 2. queue the solve of pair p = (n-1, n) on the compute stream
    const int p = n - 1;
    const int prev_pending = c->vid_pending;
    HS_CUDA(c, cudaStreamWaitEvent(c->stream, c->ev_up, 0));
    c->d_prev = c->d_ring[(n - 1) % 3];
    c->d_next = c->d_ring[n % 3];
    c->uploaded = true;
    c->timing.launches = 0;
    if ((rc = do_prepare(c))) return rc;
    HS_CUDA(c, cudaEventRecord(c->ev_k1[p % 3], c->stream));
    if ((rc = do_iterate(c, c->T))) return rc;
    const long long npx = c->plane;
    if (dt == HS_F64) {
        double* o = static_cast<double*>(c->d_vout[p & 1]);
        hs::k_split<double><<<148 * 8, 256, 0, c->stream>>>(c->d_uv[c->cur], o, o + npx, npx);
    } else {
        float* o = static_cast<float*>(c->d_vout[p & 1]);
        hs::k_split<float><<<148 * 8, 256, 0, c->stream>>>(c->d_uv[c->cur], o, o + npx, npx);
    }
    HS_CUDA(c, cudaGetLastError());
    c->timing.launches += 1;
    HS_CUDA(c, cudaEventRecord(c->ev_solved[p & 1], c->stream));
    c->vid_frames = n + 1;            // committed only now: a failed step above leaves the sequence where it was
    c->vid_pending = p;
    // 3. hand out the previous pair while this one is being solved
    if (prev_pending >= 0) {
        if ((rc = video_fetch(c, prev_pending, u, us, v, vs))) return rc;
        *pair_index = prev_pending;
    }
    return HS_OK;
}

extern "C" int hs_video_flush(hs_ctx* c, void* u, size_t us, void* v, size_t vs, int* pair_index) {
    if (!c || !pair_index) return HS_ERR_INVALID_ARG;
    *pair_index = -1;
    if (c->vid_pending < 0) return HS_OK;
    DevGuard g(c->dev);
    int rc = video_fetch(c, c->vid_pending, u, us, v, vs);
    if (rc) return rc;
    *pair_index = c->vid_pending;
    c->vid_pending = -1;
    return HS_OK;
}
