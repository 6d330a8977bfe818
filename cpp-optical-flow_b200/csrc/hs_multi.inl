// Several GPUs behind the C ABI (included by hs_api.cu): row-slab planning, seam wiring (peer access
// inside one process, CUDA IPC between processes), and group contexts that drive N devices from the
// calling host thread.  Reference call site: HornSchunckOF/main.cpp:97-98 - one getFlow call; the halo
// geometry follows the anchor of hornSchunck.cpp:54 (a rows above, w/2 rows below per sweep).
#include <unistd.h>

namespace {

// ------------------------------------------------------------------------------------------------
// planning
// ------------------------------------------------------------------------------------------------
struct SlabPlan {
    int y0, y1;   // rows the slab produces (image rows)
    int b0, b1;   // rows its flow planes hold: own rows + halo
    int f0, f1;   // frame rows it needs: buffer rows + one Sobel row per seam
    bool top, bot;
};

// Balanced contiguous bands with EVEN first rows: every slab then has the same tile-row parity (the
// fused kernel pairs image rows (2i, 2i+1)), hence the same tile pitch, and tile columns/rows of
// neighbouring slabs line up.
bool plan_slab(int H, int world, int rank, int RL, int RR, int k, SlabPlan* p, std::string* why) {
    const int units = (H + 1) / 2, base = units / world, extra = units % world;
    auto first = [&](int r) { return std::min(H, 2 * (r * base + std::min(r, extra))); };
    p->y0 = first(rank);
    p->y1 = first(rank + 1);
    const int halo_top = RL * k + ((RL * k) & 1);      // the odd extra row keeps staged tiles on even image rows
    const int halo_bot = RR * k;
    for (int r = 0; r < world; ++r) {
        const int rows = first(r + 1) - first(r);
        if (rows < std::max(std::max(halo_top, halo_bot), 1)) {
            char buf[200];
            snprintf(buf, sizeof buf, "row slab %d of %d has %d rows, fewer than the %d-row halo (k=%d): use fewer devices or a smaller k",
                     r, world, rows, std::max(halo_top, halo_bot), k);
            *why = buf;
            return false;
        }
    }
    p->top = rank > 0;
    p->bot = rank < world - 1;
    p->b0 = p->top ? p->y0 - halo_top : p->y0;
    p->b1 = p->bot ? p->y1 + halo_bot : p->y1;
    p->f0 = p->b0 - (p->top ? 1 : 0);
    p->f1 = p->b1 + (p->bot ? 1 : 0);
    return true;
}

// default temporal-blocking depth of a slab (same rule as create_single for large frames; the slabs of
// one image must all use the same k, so it is chosen once here)
int pick_slab_k(const hs_config& cfg, int RL, int RR) {
    int k = cfg.temporal_k > 0 ? cfg.temporal_k : env_int("HS_K", 0);
    if (k <= 0) k = large_frame_k(RL, RR);
    const int SY = TILE_R * TILE_NWARP;
    k = std::min(k, (SY - 3) / std::max(1, RL + RR));
    while (k > 1 && SY - (RL + RR) * k < SY / 4) --k;
    // seams use the dataflow launch: both halos must fit inside one tile pitch
    auto ok = [&](int kk) {
        const int hxl = round_up(RL * kk, 4), hxr = round_up(RR * kk, 4), vx = 128 - hxl - hxr;
        const int hyt = RL * kk + ((RL * kk) & 1), vy = (SY - hyt - RR * kk) & ~1;
        return vx > 0 && vy > 0 && hxl <= vx && hxr <= vx && hyt <= vy && RR * kk <= vy;
    };
    while (k > 1 && !ok(k)) --k;
    return k;
}

hs_config child_config(const hs_config& cfg, const SlabPlan& p, int k, int device, void* stream) {
    hs_config cc = cfg;
    cc.struct_size = sizeof(hs_config);
    cc.height = p.b1 - p.b0;
    cc.batch = 1;
    cc.device = device;
    cc.temporal_k = k;
    cc.flags = (cfg.flags & ~(uint32_t)(HS_FLAG_TOP_IS_SEAM | HS_FLAG_BOTTOM_IS_SEAM)) |
               (p.top ? HS_FLAG_TOP_IS_SEAM : 0) | (p.bot ? HS_FLAG_BOTTOM_IS_SEAM : 0);
    cc.out_row_begin = p.y0 - p.b0;
    cc.out_row_end = p.y1 - p.b0;
    cc.global_row0 = p.b0;
    cc.stream = stream;
    cc.num_devices = 0; cc.device_ids = nullptr; cc.slab_world = 0; cc.slab_rank = 0;
    return cc;
}

void note_plan(hs_ctx* c, const SlabPlan& p, int rank, int world, int img_h) {
    c->slab_rank = rank; c->slab_world = world; c->img_h = img_h;
    c->sl_y0 = p.y0; c->sl_y1 = p.y1; c->sl_b0 = p.b0; c->sl_b1 = p.b1; c->sl_f0 = p.f0; c->sl_f1 = p.f1;
    c->reverse = rank & 1;
}

// ------------------------------------------------------------------------------------------------
// seam wiring
// ------------------------------------------------------------------------------------------------
struct HandleBody {                 // what hs_slab_handle carries (<= 256 bytes, plain data)
    uint32_t magic, version;
    int32_t pid, device;
    uint64_t arena_ptr;             // valid inside the exporting process only
    uint64_t arena_bytes, inbox_off, plane_off[2];
    int32_t rank, world, width, window, k;
    int32_t b0, y0, y1;
    cudaIpcMemHandle_t ipc;
};
static_assert(sizeof(HandleBody) <= sizeof(hs_slab_handle), "hs_slab_handle too small");
constexpr uint32_t HANDLE_MAGIC = 0x48535342u;   // "HSSB"

void wire_seam(hs_ctx::Seam& L, const hs_ctx* me, const HandleBody& n, char* nbr_arena) {
    L.on = true;
    L.uv[0] = reinterpret_cast<float2*>(nbr_arena + n.plane_off[0]);
    L.uv[1] = reinterpret_cast<float2*>(nbr_arena + n.plane_off[1]);
    L.inbox = reinterpret_cast<int*>(nbr_arena + n.inbox_off);
    L.dy = me->sl_b0 - n.b0;
    L.nbr_rows = n.y1 - n.y0;
    L.nbr_parity = n.y0 & 1;
}

int export_handle(hs_ctx* c, HandleBody* h, bool want_ipc) {
    memset(h, 0, sizeof *h);
    h->magic = HANDLE_MAGIC; h->version = HS_VERSION;
    h->pid = (int32_t)getpid(); h->device = c->dev;
    h->arena_ptr = (uint64_t)(uintptr_t)c->arena;
    h->arena_bytes = c->arena_bytes; h->inbox_off = c->inbox_off;
    for (int i = 0; i < 2; ++i) h->plane_off[i] = c->plane_off[i];
    h->rank = c->slab_rank; h->world = c->slab_world; h->width = c->W; h->window = c->w; h->k = c->k;
    h->b0 = c->sl_b0; h->y0 = c->sl_y0; h->y1 = c->sl_y1;
    if (want_ipc) {
        DevGuard g(c->dev);
        HS_CUDA(c, cudaIpcGetMemHandle(&h->ipc, c->arena));
    }
    return HS_OK;
}

int connect_side(hs_ctx* c, hs_ctx::Seam& L, void** ipc_slot, const HandleBody& n, int want_rank) {
    if (n.magic != HANDLE_MAGIC || n.version != HS_VERSION) return fail(c, HS_ERR_INVALID_ARG, "not an hs_slab_handle of this library version");
    if (n.rank != want_rank || n.world != c->slab_world || n.width != c->W || n.window != c->w || n.k != c->k)
        return fail(c, HS_ERR_INVALID_ARG, "neighbour handle does not match: rank %d/%d width %d window %d k %d, expected rank %d/%d width %d window %d k %d",
                    n.rank, n.world, n.width, n.window, n.k, want_rank, c->slab_world, c->W, c->w, c->k);
    char* base = nullptr;
    if (n.pid == (int32_t)getpid()) {                       // same process: the pointer itself, plus peer access
        base = reinterpret_cast<char*>((uintptr_t)n.arena_ptr);
        if (n.device != c->dev) {
            int can = 0;
            HS_CUDA(c, cudaDeviceCanAccessPeer(&can, c->dev, n.device));
            if (!can) return fail(c, HS_ERR_UNSUPPORTED, "device %d cannot access device %d's memory (no peer access)", c->dev, n.device);
            cudaError_t e = cudaDeviceEnablePeerAccess(n.device, 0);
            if (e == cudaErrorPeerAccessAlreadyEnabled) cudaGetLastError();
            else if (e != cudaSuccess) return fail(c, HS_ERR_CUDA, "cudaDeviceEnablePeerAccess(%d): %s", n.device, cudaGetErrorString(e));
        }
    } else {                                                // another process: CUDA IPC
        void* p = nullptr;
        HS_CUDA(c, cudaIpcOpenMemHandle(&p, n.ipc, cudaIpcMemLazyEnablePeerAccess));
        *ipc_slot = p;
        base = static_cast<char*>(p);
    }
    wire_seam(L, c, n, base);
    return HS_OK;
}

int slab_connect(hs_ctx* c, const HandleBody* up, const HandleBody* dn) {
    if (c->kernel_id != 1) return fail(c, HS_ERR_UNSUPPORTED, "row slabs need the fused kernel (window 2..9)");
    if ((c->top_seam && !up) || (c->bot_seam && !dn)) return fail(c, HS_ERR_INVALID_ARG, "missing neighbour handle for a seam of this slab");
    DevGuard g(c->dev);
    int rc;
    if (c->top_seam && (rc = connect_side(c, c->up, &c->ipc_up, *up, c->slab_rank - 1))) return rc;
    if (c->bot_seam && (rc = connect_side(c, c->dn, &c->ipc_dn, *dn, c->slab_rank + 1))) return rc;
    bool ok = true;
    tile_dispatch(c, [&](auto t) { ok = decltype(t)::dataflow_ok(c->k, c->oy0 + c->grow0); });
    if (!ok) return fail(c, HS_ERR_UNSUPPORTED, "temporal_k=%d: halo wider than one tile pitch, seams cannot use the dataflow launch", c->k);
    int coop = 0;
    cudaDeviceGetAttribute(&coop, cudaDevAttrCooperativeLaunch, c->dev);
    if (!coop) return fail(c, HS_ERR_UNSUPPORTED, "device lacks cooperative launch");
    c->linked = true;
    return HS_OK;
}

int create_slab_rank(const hs_config& cfg, hs_ctx** out) {
    if (cfg.precision != HS_PREC_F32) return fail(nullptr, HS_ERR_UNSUPPORTED, "row slabs run the fp32 fused kernel");
    if (cfg.batch > 1) return fail(nullptr, HS_ERR_UNSUPPORTED, "row slabs need batch == 1");
    if (cfg.window_size < 1 || cfg.width < 1 || cfg.height < 1) return fail(nullptr, HS_ERR_INVALID_ARG, "bad geometry");
    if (cfg.slab_rank < 0 || cfg.slab_rank >= cfg.slab_world) return fail(nullptr, HS_ERR_INVALID_ARG, "slab_rank %d not in [0, %d)", cfg.slab_rank, cfg.slab_world);
    const int a = cfg.window_size - cfg.window_size / 2 - 1, RL = a, RR = cfg.window_size - 1 - a;
    SlabPlan p0;
    std::string why;
    const int k = pick_slab_k(cfg, RL, RR);
    if (!plan_slab(cfg.height, cfg.slab_world, cfg.slab_rank, RL, RR, k, &p0, &why)) return fail(nullptr, HS_ERR_INVALID_ARG, "%s", why.c_str());
    hs_config cc = child_config(cfg, p0, k, cfg.device, cfg.stream);
    int rc = create_single(cc, out);
    if (rc) return rc;
    note_plan(*out, p0, cfg.slab_rank, cfg.slab_world, cfg.height);
    if ((*out)->k != k || (*out)->kernel_id != 1) {
        destroy_impl(*out); *out = nullptr;
        return fail(nullptr, HS_ERR_UNSUPPORTED, "row slabs need the fused kernel (window 2..9) at the planned k");
    }
    return HS_OK;
}

// ------------------------------------------------------------------------------------------------
// NCCL exchange (A/B variant, single process): libnccl is dlopen'ed so the library has no link-time
// dependency on it
// ------------------------------------------------------------------------------------------------
}  // namespace
#include <dlfcn.h>
struct NcclState {
    void* lib = nullptr;
    std::vector<void*> comms;
    int (*CommInitAll)(void**, int, const int*) = nullptr;
    int (*CommDestroy)(void*) = nullptr;
    int (*GroupStart)() = nullptr;
    int (*GroupEnd)() = nullptr;
    int (*Send)(const void*, size_t, int, int, void*, cudaStream_t) = nullptr;
    int (*Recv)(void*, size_t, int, int, void*, cudaStream_t) = nullptr;
    const char* (*GetErrorString)(int) = nullptr;
};
namespace {

int nccl_open(hs_ctx* g, const std::vector<int>& devs) {
    NcclState* st = new NcclState();
    const char* names[] = {"libnccl.so.2", "libnccl.so"};
    for (const char* n : names)
        if ((st->lib = dlopen(n, RTLD_NOW | RTLD_GLOBAL))) break;
    if (!st->lib) { delete st; return fail(g, HS_ERR_NCCL, "HS_EXCHANGE_NCCL: libnccl.so.2 not found (%s)", dlerror()); }
#define HS_NCCL_SYM(field, name)                                                        \
    *reinterpret_cast<void**>(&st->field) = dlsym(st->lib, name);                       \
    if (!st->field) { delete st; return fail(g, HS_ERR_NCCL, "libnccl lacks %s", name); }
    HS_NCCL_SYM(CommInitAll, "ncclCommInitAll")
    HS_NCCL_SYM(CommDestroy, "ncclCommDestroy")
    HS_NCCL_SYM(GroupStart, "ncclGroupStart")
    HS_NCCL_SYM(GroupEnd, "ncclGroupEnd")
    HS_NCCL_SYM(Send, "ncclSend")
    HS_NCCL_SYM(Recv, "ncclRecv")
    HS_NCCL_SYM(GetErrorString, "ncclGetErrorString")
#undef HS_NCCL_SYM
    st->comms.assign(devs.size(), nullptr);
    int r = st->CommInitAll(st->comms.data(), (int)devs.size(), devs.data());
    if (r != 0) {
        std::string msg = st->GetErrorString(r);
        delete st;
        return fail(g, HS_ERR_NCCL, "ncclCommInitAll failed: %s", msg.c_str());
    }
    g->nccl = st;
    return HS_OK;
}

void nccl_close(hs_ctx* g) {
    if (!g->nccl) return;
    for (void* cm : g->nccl->comms) if (cm) g->nccl->CommDestroy(cm);
    delete g->nccl;
    g->nccl = nullptr;
}

// halo rows of the CURRENT planes of every child across every seam, one grouped send/recv
int nccl_exchange(hs_ctx* g) {
    NcclState* st = g->nccl;
    const int n = (int)g->kids.size();
    int r = st->GroupStart();
    for (int i = 0; i < n && r == 0; ++i) {
        hs_ctx* c = g->kids[i];
        const int up = c->RL * c->k, dn = c->RR * c->k;          // rows needed from above / below
        const size_t rowf = (size_t)c->pitch * 2;                // floats per row of the interleaved {u, v} plane
        float* planes[1] = {reinterpret_cast<float*>(c->d_uv[c->cur])};
        for (float* P : planes) {
            if (i > 0) {                                          // seam above
                if (dn && r == 0) r = st->Send(P + (size_t)c->oy0 * rowf, (size_t)dn * rowf, 7 /* ncclFloat */, i - 1, st->comms[i], c->stream);
                if (up && r == 0) r = st->Recv(P + (size_t)(c->oy0 - up) * rowf, (size_t)up * rowf, 7, i - 1, st->comms[i], c->stream);
            }
            if (i < n - 1) {                                      // seam below
                if (up && r == 0) r = st->Send(P + (size_t)(c->oy1 - up) * rowf, (size_t)up * rowf, 7, i + 1, st->comms[i], c->stream);
                if (dn && r == 0) r = st->Recv(P + (size_t)c->oy1 * rowf, (size_t)dn * rowf, 7, i + 1, st->comms[i], c->stream);
            }
        }
    }
    const int r2 = st->GroupEnd();
    if (r == 0) r = r2;
    if (r != 0) return fail(g, HS_ERR_NCCL, "NCCL halo exchange failed: %s", st->GetErrorString(r));
    return HS_OK;
}

// ------------------------------------------------------------------------------------------------
// group contexts: N devices behind one hs_ctx
// ------------------------------------------------------------------------------------------------
int adopt_error(hs_ctx* g, hs_ctx* kid, int rc) {
    if (rc) g->err = kid->err;
    return rc;
}

int create_group(const hs_config& cfg, hs_ctx** out) {
    const int n = cfg.num_devices;
    if (cfg.precision != HS_PREC_F32 && cfg.decomposition == HS_DECOMP_ROW_SLAB)
        return fail(nullptr, HS_ERR_UNSUPPORTED, "row slabs run the fp32 fused kernel (HS_PREC_F64 is a single-device diagnostic path)");
    if (cfg.width < 1 || cfg.height < 1 || cfg.window_size < 1) return fail(nullptr, HS_ERR_INVALID_ARG, "bad geometry");
    if (cfg.decomposition != HS_DECOMP_BATCH && cfg.decomposition != HS_DECOMP_ROW_SLAB)
        return fail(nullptr, HS_ERR_INVALID_ARG, "unknown decomposition %d", cfg.decomposition);
    if (cfg.stream) return fail(nullptr, HS_ERR_INVALID_ARG, "a multi-device context creates its own streams (hs_config.stream must be NULL)");
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev == 0)
        return fail(nullptr, HS_ERR_CUDA, "no usable CUDA device (%s); this library has no CPU fallback",
                    e == cudaSuccess ? "device count 0" : cudaGetErrorString(e));
    std::vector<int> devs(n);
    for (int i = 0; i < n; ++i) {
        devs[i] = cfg.device_ids ? cfg.device_ids[i] : i;
        if (devs[i] < 0 || devs[i] >= ndev) return fail(nullptr, HS_ERR_INVALID_ARG, "device_ids[%d] = %d out of range (%d devices)", i, devs[i], ndev);
    }
    bool all_same = true, all_distinct = true;
    for (int i = 0; i < n; ++i)
        for (int j = i + 1; j < n; ++j) {
            if (devs[i] == devs[j]) all_distinct = false; else all_same = false;
        }
    hs_ctx* g = new (std::nothrow) hs_ctx();
    if (!g) return fail(nullptr, HS_ERR_OOM, "out of host memory");
    g->cfg = cfg; g->dev = devs[0];
    g->W = cfg.width; g->H = cfg.height; g->B = cfg.batch > 0 ? cfg.batch : 1;
    g->w = cfg.window_size; g->T = cfg.max_iterations; g->alpha = cfg.alpha;
    g->a = g->w - g->w / 2 - 1; g->RL = g->a; g->RR = g->w - 1 - g->a;
    g->oy0 = 0; g->oy1 = g->H;
    g->decomp = cfg.decomposition; g->exchange = cfg.exchange;
    auto bail = [&](int code) { g_create_err = g->err; destroy_impl(g); return code; };

    if (g->decomp == HS_DECOMP_BATCH) {
        if (!all_distinct) return bail(fail(g, HS_ERR_INVALID_ARG, "HS_DECOMP_BATCH needs distinct devices"));
        if (g->B < n) return bail(fail(g, HS_ERR_INVALID_ARG, "batch %d smaller than num_devices %d", g->B, n));
        for (int i = 0; i < n; ++i) {
            hs_config cc = cfg;
            cc.struct_size = sizeof(hs_config);
            cc.num_devices = 0; cc.device_ids = nullptr;
            cc.device = devs[i];
            cc.batch = (g->B - i + n - 1) / n;                // pairs i, i+n, i+2n, ...
            hs_ctx* kid = nullptr;
            int rc = create_single(cc, &kid);
            if (rc) { g->err = g_create_err; return bail(rc); }
            g->kids.push_back(kid);
        }
    } else {
        if (g->B != 1) return bail(fail(g, HS_ERR_UNSUPPORTED, "HS_DECOMP_ROW_SLAB splits ONE image: batch must be 1"));
        if (!all_distinct && !all_same) return bail(fail(g, HS_ERR_INVALID_ARG, "device_ids must be all distinct, or all the same (single-GPU emulation)"));
        g->emulate = all_same;
        if (g->emulate && n > EMU_MAXS) return bail(fail(g, HS_ERR_UNSUPPORTED, "at most %d row slabs on one device", EMU_MAXS));
        if (g->emulate && g->exchange == HS_EXCHANGE_NCCL) return bail(fail(g, HS_ERR_UNSUPPORTED, "HS_EXCHANGE_NCCL needs distinct devices"));
        const int k = pick_slab_k(cfg, g->RL, g->RR);
        void* shared_stream = nullptr;
        for (int i = 0; i < n; ++i) {
            SlabPlan p;
            std::string why;
            if (!plan_slab(cfg.height, n, i, g->RL, g->RR, k, &p, &why)) return bail(fail(g, HS_ERR_INVALID_ARG, "%s", why.c_str()));
            hs_config cc = child_config(cfg, p, k, devs[i], g->emulate ? shared_stream : nullptr);
            hs_ctx* kid = nullptr;
            int rc = create_single(cc, &kid);
            if (rc) { g->err = g_create_err; return bail(rc); }
            g->kids.push_back(kid);
            note_plan(kid, p, i, n, cfg.height);
            if (kid->k != k || kid->kernel_id != 1)
                return bail(fail(g, HS_ERR_UNSUPPORTED, "row slabs need the fused kernel (window 2..9) at the planned k=%d", k));
            if (g->emulate && i == 0) shared_stream = kid->stream;   // one stream: the slabs share one launch
        }
        g->k = k; g->kernel_id = 1;
        if (g->exchange == HS_EXCHANGE_PEER) {
            std::vector<HandleBody> hb(n);
            for (int i = 0; i < n; ++i) { int rc = export_handle(g->kids[i], &hb[i], false); if (rc) return bail(adopt_error(g, g->kids[i], rc)); }
            for (int i = 0; i < n; ++i) {
                int rc = slab_connect(g->kids[i], i > 0 ? &hb[i - 1] : nullptr, i < n - 1 ? &hb[i + 1] : nullptr);
                if (rc) return bail(adopt_error(g, g->kids[i], rc));
            }
            if (g->emulate) {      // one launch holds the tiles of every slab: its done[] lives in child 0
                hs_ctx* k0 = g->kids[0];
                size_t cap = 0;
                for (hs_ctx* kid : g->kids) cap += kid->done_cap;
                DevGuard dg(k0->dev);
                cudaFree(k0->d_done); k0->d_done = nullptr; k0->done_cap = 0;
                if (cudaMalloc(&k0->d_done, std::max<size_t>(cap, 1) * sizeof(int)) != cudaSuccess)
                    return bail(fail(g, HS_ERR_OOM, "out of device memory"));
                k0->done_cap = cap;
            }
        } else if (g->exchange == HS_EXCHANGE_NCCL) {
            int rc = nccl_open(g, devs);
            if (rc) return bail(rc);
        } else {
            return bail(fail(g, HS_ERR_INVALID_ARG, "unknown exchange %d", g->exchange));
        }
        for (hs_ctx* kid : g->kids) {
            DevGuard dg(kid->dev);
            if (cudaEventCreateWithFlags(&kid->ev_x, cudaEventDisableTiming) != cudaSuccess)
                return bail(fail(g, HS_ERR_CUDA, "cudaEventCreate failed"));
        }
    }
    g->timing.temporal_k = g->kids[0]->k;
    g->timing.kernel_id = g->kids[0]->kernel_id;
    *out = g;
    return HS_OK;
}

// every child's stream waits for the neighbours' work recorded so far (device-side, no host sync)
int group_fence(hs_ctx* g) {
    const int n = (int)g->kids.size();
    if (g->emulate) return HS_OK;                                // one stream
    for (hs_ctx* kid : g->kids) {
        DevGuard dg(kid->dev);
        HS_CUDA(g, cudaEventRecord(kid->ev_x, kid->stream));
    }
    for (int i = 0; i < n; ++i) {
        hs_ctx* kid = g->kids[i];
        DevGuard dg(kid->dev);
        if (i > 0) HS_CUDA(g, cudaStreamWaitEvent(kid->stream, g->kids[i - 1]->ev_x, 0));
        if (i < n - 1) HS_CUDA(g, cudaStreamWaitEvent(kid->stream, g->kids[i + 1]->ev_x, 0));
    }
    return HS_OK;
}

int group_upload(hs_ctx* g, const uint8_t* prev, size_t ps, size_t pis, const uint8_t* next, size_t ns, size_t nis) {
    if (!prev || !next) return fail(g, HS_ERR_INVALID_ARG, "null frame pointer");
    const int n = (int)g->kids.size();
    for (int i = 0; i < n; ++i) {
        hs_ctx* kid = g->kids[i];
        DevGuard dg(kid->dev);
        int rc;
        if (g->decomp == HS_DECOMP_BATCH)
            rc = do_upload(kid, prev + (size_t)i * pis, ps, pis * n, next + (size_t)i * nis, ns, nis * n);
        else
            rc = do_upload(kid, prev + (size_t)kid->sl_f0 * ps, ps, 0, next + (size_t)kid->sl_f0 * ns, ns, 0);
        if (rc) return adopt_error(g, kid, rc);
    }
    g->uploaded = true;
    return HS_OK;
}

int group_prepare(hs_ctx* g) {
    int rc;
    // a child's prepare zeroes its arena: the neighbours' previous sweeps (which store into it) must be done
    if (g->decomp == HS_DECOMP_ROW_SLAB && (rc = group_fence(g))) return rc;
    for (hs_ctx* kid : g->kids) {
        DevGuard dg(kid->dev);
        HS_CUDA(g, cudaEventRecord(kid->ev[1], kid->stream));
        if ((rc = do_prepare(kid))) return adopt_error(g, kid, rc);
        HS_CUDA(g, cudaEventRecord(kid->ev[2], kid->stream));
    }
    g->prepared = true;
    return HS_OK;
}

int group_iterate(hs_ctx* g, int iters) {
    int rc;
    const int n = (int)g->kids.size();
    if (g->decomp == HS_DECOMP_BATCH) {
        for (hs_ctx* kid : g->kids) {
            DevGuard dg(kid->dev);
            if ((rc = do_iterate(kid, iters))) return adopt_error(g, kid, rc);
        }
        return HS_OK;
    }
    if ((rc = group_fence(g))) return rc;                        // every slab is prepared / has its halos
    if (g->exchange == HS_EXCHANGE_NCCL) {
        int left = iters;
        while (left > 0) {
            const int kk = std::min(g->k, left);
            for (hs_ctx* kid : g->kids) {
                DevGuard dg(kid->dev);
                if ((rc = do_iterate(kid, kk))) return adopt_error(g, kid, rc);
            }
            left -= kk;
            if (left > 0 && (rc = nccl_exchange(g))) return rc;
        }
        return HS_OK;
    }
    if (!g->emulate) {
        for (hs_ctx* kid : g->kids) {
            DevGuard dg(kid->dev);
            if ((rc = do_iterate(kid, iters))) return adopt_error(g, kid, rc);
        }
        return HS_OK;
    }
    // single-GPU emulation: all slabs in ONE cooperative launch (kernels that wait on one another must
    // never be separate launches on one device)
    if (iters <= 0) return HS_OK;
    hs_ctx* k0 = g->kids[0];
    DevGuard dg(k0->dev);
    cudaError_t e = cudaSuccess;
    tile_dispatch(k0, [&](auto t) { e = decltype(t)::launch_group(g->kids.data(), n, g->k, iters); });
    if (e != cudaSuccess) return fail(g, HS_ERR_CUDA, "row-slab group launch failed: %s", cudaGetErrorString(e));
    const int phases = (iters + g->k - 1) / g->k;
    for (hs_ctx* kid : g->kids) {
        if (phases & 1) kid->cur ^= 1;
        kid->phase_count += phases;
    }
    k0->timing.launches += 1;
    return HS_OK;
}

int group_download(hs_ctx* g, void* u, size_t us, size_t uis, void* v, size_t vs, size_t vis, int dt) {
    if (!u || !v) return fail(g, HS_ERR_INVALID_ARG, "null output pointer");
    const int n = (int)g->kids.size();
    for (int i = 0; i < n; ++i) {
        hs_ctx* kid = g->kids[i];
        DevGuard dg(kid->dev);
        int rc;
        if (g->decomp == HS_DECOMP_BATCH)
            rc = do_download(kid, static_cast<char*>(u) + (size_t)i * uis, us, uis * n, static_cast<char*>(v) + (size_t)i * vis, vs, vis * n, dt);
        else
            rc = do_download(kid, static_cast<char*>(u) + (size_t)kid->sl_y0 * us, us, 0, static_cast<char*>(v) + (size_t)kid->sl_y0 * vs, vs, 0, dt);
        if (rc) return adopt_error(g, kid, rc);
    }
    return HS_OK;
}

int group_sync(hs_ctx* g) {
    for (hs_ctx* kid : g->kids) {
        DevGuard dg(kid->dev);
        HS_CUDA(g, cudaStreamSynchronize(kid->stream));
    }
    if (g->timing_pending) {      // device time of the last hs_solve_device: the slowest child
        g->timing_pending = false;
        float prep = 0.f, iter = 0.f, tot = 0.f;
        int launches = 0;
        for (hs_ctx* kid : g->kids) {
            DevGuard dg(kid->dev);
            prep = std::max(prep, ev_ms(kid->ev[1], kid->ev[2]));
            iter = std::max(iter, ev_ms(kid->ev[2], kid->ev[3]));
            tot = std::max(tot, ev_ms(kid->ev[1], kid->ev[3]));
            launches += kid->timing.launches;
        }
        g->timing.h2d_ms = g->timing.d2h_ms = 0.f;
        g->timing.prepare_ms = prep; g->timing.iterate_ms = iter; g->timing.total_ms = tot;
        g->timing.launches = launches;
    }
    return HS_OK;
}

int group_solve_device(hs_ctx* g) {
    if (!g->uploaded) return fail(g, HS_ERR_STATE, "hs_solve_device before frames were uploaded");
    int rc;
    for (hs_ctx* kid : g->kids) kid->timing.launches = 0;
    if ((rc = group_prepare(g))) return rc;
    if ((rc = group_iterate(g, g->T))) return rc;
    for (hs_ctx* kid : g->kids) {
        DevGuard dg(kid->dev);
        HS_CUDA(g, cudaEventRecord(kid->ev[3], kid->stream));
    }
    g->timing_pending = true;
    return HS_OK;
}

int group_solve(hs_ctx* g, const uint8_t* prev, size_t ps, size_t pis, const uint8_t* next, size_t ns, size_t nis,
                void* u, size_t us, size_t uis, void* v, size_t vs, size_t vis, int dt) {
    int rc;
    if ((rc = group_upload(g, prev, ps, pis, next, ns, nis))) return rc;
    if ((rc = group_solve_device(g))) return rc;
    if ((rc = group_download(g, u, us, uis, v, vs, vis, dt))) return rc;
    return group_sync(g);
}

void group_destroy(hs_ctx* g) {
    for (hs_ctx* kid : g->kids) {
        if (kid->stream) { DevGuard dg(kid->dev); cudaStreamSynchronize(kid->stream); }
    }
    nccl_close(g);
    bool first = true;
    for (hs_ctx* kid : g->kids) {
        if (g->emulate && !first) { kid->own_stream = false; }   // the shared stream belongs to child 0
        first = false;
    }
    // children 1.. of an emulated group share child 0's stream: destroy them first
    for (size_t i = g->kids.size(); i-- > 0;) {
        hs_ctx* kid = g->kids[i];
        if (kid->ev_x) { DevGuard dg(kid->dev); cudaEventDestroy(kid->ev_x); kid->ev_x = nullptr; }
        destroy_impl(kid);
    }
    g->kids.clear();
}

}  // namespace

extern "C" int hs_get_slab_info(hs_ctx* c, hs_slab_info* o) {
    if (!c || !o) return HS_ERR_INVALID_ARG;
    if (c->slab_world <= 1) return fail(c, HS_ERR_STATE, "not a row-slab context");
    o->rank = c->slab_rank; o->world = c->slab_world;
    o->own_begin = c->sl_y0; o->own_end = c->sl_y1;
    o->buf_begin = c->sl_b0; o->buf_end = c->sl_b1;
    o->frame_begin = c->sl_f0; o->frame_end = c->sl_f1;
    o->temporal_k = c->k;
    o->halo_top = c->top_seam ? c->RL * c->k : 0;
    o->halo_bottom = c->bot_seam ? c->RR * c->k : 0;
    return HS_OK;
}

// pure host arithmetic (no device needed): where slab `rank` of `world` lies for a given k
extern "C" int hs_plan_slab(int32_t height, int32_t world, int32_t rank, int32_t window_size, int32_t temporal_k,
                            hs_slab_info* o) {
    if (!o || height < 1 || world < 1 || rank < 0 || rank >= world || window_size < 1 || temporal_k < 1) return HS_ERR_INVALID_ARG;
    const int a = window_size - window_size / 2 - 1, RL = a, RR = window_size - 1 - a;
    SlabPlan p;
    std::string why;
    if (!plan_slab(height, world, rank, RL, RR, temporal_k, &p, &why)) return fail(nullptr, HS_ERR_INVALID_ARG, "%s", why.c_str());
    o->rank = rank; o->world = world;
    o->own_begin = p.y0; o->own_end = p.y1; o->buf_begin = p.b0; o->buf_end = p.b1;
    o->frame_begin = p.f0; o->frame_end = p.f1;
    o->temporal_k = temporal_k;
    o->halo_top = p.top ? RL * temporal_k : 0;
    o->halo_bottom = p.bot ? RR * temporal_k : 0;
    return HS_OK;
}

extern "C" int hs_slab_export(hs_ctx* c, hs_slab_handle* out) {
    if (!c || !out) return HS_ERR_INVALID_ARG;
    if (c->slab_world <= 1 || !c->kids.empty()) return fail(c, HS_ERR_STATE, "hs_slab_export needs a slab_world > 1 context");
    memset(out, 0, sizeof *out);
    HandleBody h;
    int rc = export_handle(c, &h, true);
    if (rc) return rc;
    memcpy(out->bytes, &h, sizeof h);
    return HS_OK;
}

extern "C" int hs_slab_connect(hs_ctx* c, const hs_slab_handle* up, const hs_slab_handle* down) {
    if (!c) return HS_ERR_INVALID_ARG;
    if (c->slab_world <= 1 || !c->kids.empty()) return fail(c, HS_ERR_STATE, "hs_slab_connect needs a slab_world > 1 context");
    if (c->linked) return fail(c, HS_ERR_STATE, "already connected");
    HandleBody hu, hd;
    if (up) memcpy(&hu, up->bytes, sizeof hu);
    if (down) memcpy(&hd, down->bytes, sizeof hd);
    return slab_connect(c, up ? &hu : nullptr, down ? &hd : nullptr);
}
