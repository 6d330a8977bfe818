"""Row-slab decomposition of one large frame pair over several GPUs (BASELINE config 5).

One process per GPU (torch.distributed over NCCL/NVLink).  Rank r owns a contiguous band of
rows; per fused launch of k sweeps its kernel reads a*k rows above and (w/2)*k rows below the
band (a = anchor = w - w/2 - 1, hornSchunck.cpp:54), so after every launch each seam exchanges
exactly those rows of u and of v with its two neighbours - rows are contiguous in the pitched
planes, so a halo is one contiguous send, no packing kernel.  Coefficients are never exchanged:
every rank gets the frame rows of its band + halo (+1 seam row for the Sobel taps) and recomputes
them.  The result is bit-identical to the single-GPU solve (Jacobi has no ordering freedom).

Deep halos: with `depth` d the slab keeps d*a*k / d*(w/2)*k halo rows and exchanges them only
every d fused launches; launch j of such a cycle also advances the d-1-j halo layers that later
launches of the cycle still read (a fraction of a percent of redundant rows).  That divides the
number of exchanges - and of host round trips through Python/NCCL, which is what bounds 8 GPUs,
where a launch is only ~270 us of device time - by d.

The exchange logic is written against torch tensors only, so the same code runs on CPU tensors
with the gloo backend in tests/test_slab_gloo.py.
"""
from __future__ import annotations

import json
import os
import time
from dataclasses import dataclass

import numpy as np


def radii(window: int):
    """(taps above/left, taps below/right) of the box window: anchor a and w-1-a."""
    a = window - window // 2 - 1
    return a, window - 1 - a


def partition_rows(height: int, world: int):
    """Balanced contiguous row bands; the first (height % world) bands get one extra row."""
    base, extra = divmod(height, world)
    bounds, y = [], 0
    for r in range(world):
        n = base + (1 if r < extra else 0)
        bounds.append((y, y + n))
        y += n
    return bounds


@dataclass
class SlabGeometry:
    rank: int
    world: int
    height: int          # full image
    width: int
    window: int
    k: int               # sweeps fused per launch
    depth: int           # fused launches between two halo exchanges (halo = depth * radius * k rows)
    y0: int              # owned rows [y0, y1) in image coordinates
    y1: int
    b0: int              # buffer rows [b0, b1) = owned rows + halos, clipped to the image
    b1: int
    f0: int              # frame rows [f0, f1) = buffer rows + one Sobel row at each seam
    f1: int
    top_seam: bool
    bottom_seam: bool

    @property
    def halo_top(self):
        return self.y0 - self.b0

    @property
    def halo_bottom(self):
        return self.b1 - self.y1

    @property
    def out_rows(self):
        """Owned rows in buffer-local coordinates."""
        return (self.y0 - self.b0, self.y1 - self.b0)

    @property
    def rows(self):
        return self.b1 - self.b0


def plan(height: int, width: int, world: int, rank: int, window: int, k: int, depth: int = 1) -> SlabGeometry:
    rl, rr = radii(window)
    bounds = partition_rows(height, world)
    y0, y1 = bounds[rank]
    need = max(rl, rr) * k * depth
    if world > 1 and min(b - a for a, b in bounds) < need:
        raise ValueError(f"row slabs of {min(b - a for a, b in bounds)} rows are thinner than the {need}-row halo "
                         f"(window {window}, k {k}, depth {depth}); use fewer ranks, a smaller k or depth")
    b0 = max(0, y0 - rl * k * depth) if rank > 0 else y0
    b1 = min(height, y1 + rr * k * depth) if rank < world - 1 else y1
    top_seam, bottom_seam = b0 > 0, b1 < height
    return SlabGeometry(rank, world, height, width, window, k, depth, y0, y1, b0, b1,
                        b0 - (1 if top_seam else 0), b1 + (1 if bottom_seam else 0), top_seam, bottom_seam)


def cycle_rows(geom: SlabGeometry, j: int, launches: int):
    """Buffer rows launch j (0-based) of an exchange cycle of `launches` fused launches has to
    produce: the owned rows plus the halo layers the remaining launches of the cycle still read."""
    rl, rr = radii(geom.window)
    ext = launches - 1 - j
    o0, o1 = geom.out_rows
    r0 = max(0, o0 - ext * rl * geom.k) if geom.rank > 0 else o0
    r1 = min(geom.rows, o1 + ext * rr * geom.k) if geom.rank < geom.world - 1 else o1
    return r0, r1


def exchange_halos(geom: SlabGeometry, planes, group=None):
    """Refresh the halo rows of `planes` (2-D torch tensors [buffer rows, pitch], the CURRENT u and
    v of this rank) from the neighbours' owned rows.  Neighbour-only send/recv, one batch."""
    import torch.distributed as dist
    rl, rr = radii(geom.window)
    up, dn = rl * geom.k * geom.depth, rr * geom.k * geom.depth   # rows needed from above / below
    o0, o1 = geom.out_rows
    ops = []
    for t in planes:
        if geom.rank > 0:                     # seam above: my top halo <- neighbour's last rows
            if dn:
                ops.append(dist.P2POp(dist.isend, t[o0:o0 + dn], geom.rank - 1, group))
            if up:
                ops.append(dist.P2POp(dist.irecv, t[o0 - up:o0], geom.rank - 1, group))
        if geom.rank < geom.world - 1:        # seam below
            if up:
                ops.append(dist.P2POp(dist.isend, t[o1 - up:o1], geom.rank + 1, group))
            if dn:
                ops.append(dist.P2POp(dist.irecv, t[o1:o1 + dn], geom.rank + 1, group))
    if not ops:
        return
    for req in dist.batch_isend_irecv(ops):
        req.wait()


class _DevMem:
    """Expose a raw device pointer to torch through __cuda_array_interface__ (no copy)."""

    def __init__(self, ptr, shape, typestr="<f4"):
        self.__cuda_array_interface__ = {"shape": tuple(shape), "typestr": typestr, "data": (int(ptr), False),
                                         "version": 3, "strides": None}


class DeviceSlab:
    """One rank's slab on its GPU: an hs_ctx sized for band + halo, plus torch views of u, v."""

    def __init__(self, geom: SlabGeometry, iterations: int, alpha: float, device: int):
        import torch
        from . import hs_ctypes as H
        from .horn_schunck import Solver
        self.geom, self.iterations = geom, iterations
        self._torch = torch
        self._device = torch.device("cuda", device)
        # One explicit stream carries the kernels AND is torch's current stream around the NCCL
        # send/recv, so launches and halo exchanges are ordered on the device without host syncs.
        # (A NULL stream in hs_config means "private stream", so the default stream cannot be used.)
        self.stream = torch.cuda.Stream(device=self._device)
        self.comm_stream = torch.cuda.Stream(device=self._device)   # halo exchange, overlapped with interior rows
        flags = (H.FLAG_TOP_IS_SEAM if geom.top_seam else 0) | (H.FLAG_BOTTOM_IS_SEAM if geom.bottom_seam else 0)
        self.solver = Solver(geom.width, geom.rows, geom.window, iterations, alpha, device=device,
                             temporal_k=geom.k, flags=flags, out_rows=geom.out_rows, stream=self.stream.cuda_stream,
                             global_row0=geom.b0)
        self.k = self.solver.timing().temporal_k
        if self.k != geom.k:
            raise ValueError(f"library chose k={self.k}, slab plan was made for k={geom.k}")
        self._views = {}

    def close(self):
        self.solver.close()

    def upload(self, prev_rows, next_rows):
        """prev_rows/next_rows: uint8 [f1 - f0, width] = frame rows geom.f0 .. geom.f1 of the image."""
        self.solver.upload(prev_rows, next_rows)

    def planes(self):
        """[torch view [rows, 2 * pitch]] of the current flow plane ({u, v} interleaved per pixel; the two
        planes ping-pong).  A halo is still a contiguous block of rows."""
        dv = self.solver.device_view()
        ptr = dv.uv
        if ptr not in self._views:
            self._views[ptr] = self._torch.as_tensor(_DevMem(ptr, (dv.height, dv.flow_pitch // 4)), device=self._device)
        return [self._views[ptr]]

    def run(self, exchange=exchange_halos, group=None, overlap=True):
        """prepare + `iterations` sweeps; halos are exchanged after every `depth` fused launches.
        Asynchronous: everything is queued on self.stream (and self.comm_stream).

        overlap: on the last launch of a cycle the rows the neighbours read are produced first (two
        thin strip launches), their exchange runs on a second stream, and the interior rows - the
        bulk of the work - are computed meanwhile; the next launch waits for both."""
        torch = self._torch
        g = self.geom
        rl, rr = radii(g.window)
        o0, o1 = g.out_rows
        strip = max(rl, rr) * self.k * g.depth              # rows a neighbour reads from us
        overlap = overlap and g.world > 1 and (o1 - o0) >= 4 * strip and strip > 0
        launches = 0
        with torch.cuda.stream(self.stream):
            self.solver.prepare()
            left = self.iterations
            while left > 0:
                cyc = min(g.depth, (left + self.k - 1) // self.k)       # fused launches of this cycle
                for j in range(cyc):
                    kk = min(self.k, left)
                    left -= kk
                    r0, r1 = cycle_rows(g, j, cyc)
                    last = j == cyc - 1
                    if g.world == 1 or not (last and left > 0):
                        self.solver.iterate_rows(kk, r0, r1, True)     # mid-cycle, or nothing follows
                        launches += 1
                    elif not overlap:
                        self.solver.iterate_rows(kk, r0, r1, True)
                        launches += 1
                        exchange(g, self.planes(), group)
                    else:                                               # r0, r1 == o0, o1 here
                        top, bot = g.top_seam, g.bottom_seam
                        a = o0 + (strip if top else 0)
                        b = o1 - (strip if bot else 0)
                        if top:
                            self.solver.iterate_rows(kk, o0, a, False)
                        if bot:
                            self.solver.iterate_rows(kk, b, o1, False)
                        strips_done = torch.cuda.Event()
                        strips_done.record(self.stream)
                        self.solver.iterate_rows(kk, a, b, True)        # interior; flips the planes
                        launches += 1 + int(top) + int(bot)
                        with torch.cuda.stream(self.comm_stream):
                            self.comm_stream.wait_event(strips_done)
                            exchange(g, self.planes(), group)           # planes() = the NEW planes now
                            exchanged = torch.cuda.Event()
                            exchanged.record(self.comm_stream)
                        self.stream.wait_event(exchanged)
        return launches

    def download(self, dtype=np.float32):
        return self.solver.download(dtype)


def solve_slabs_single_process(prev, nxt, window, iterations, alpha, nslab, temporal_k=0, device=0, depth=1):
    """N slab contexts on ONE GPU, halos copied device-to-device by the host between exchange
    cycles.  Exercises exactly the kernels / row ranges / seam handling of the multi-GPU path (used
    by the GPU tests; never run waiting kernels of several ranks concurrently on one GPU)."""
    import torch
    from .horn_schunck import Solver
    H, W = prev.shape
    probe = Solver(W, max(H // nslab, 1), window, iterations, alpha, device=device, temporal_k=temporal_k)
    k = probe.timing().temporal_k
    probe.close()
    # the library sizes its default k for a whole small frame; a slab must find k * depth sweeps'
    # worth of halo rows inside its neighbour
    while k > 1 and max(radii(window)) * k * depth > max(H // nslab, 1):
        k -= 1
    geoms = [plan(H, W, nslab, r, window, k, depth) for r in range(nslab)]
    slabs = [DeviceSlab(g, iterations, alpha, device) for g in geoms]
    rl, rr = radii(window)
    up, dn = rl * k * depth, rr * k * depth
    try:
        for s in slabs:
            g = s.geom
            s.upload(prev[g.f0:g.f1], nxt[g.f0:g.f1])
            s.solver.prepare()
        left = iterations
        while left > 0:
            cyc = min(depth, (left + k - 1) // k)
            for j in range(cyc):
                kk = min(k, left)
                left -= kk
                for s in slabs:
                    r0, r1 = cycle_rows(s.geom, j, cyc)
                    s.solver.iterate_rows(kk, r0, r1, True)
            if left > 0:
                for s in slabs:
                    s.solver.sync()
                views = [s.planes() for s in slabs]
                for r in range(nslab - 1):               # seam between slab r (above) and r + 1 (below)
                    (a0, a1), (c0, c1) = geoms[r].out_rows, geoms[r + 1].out_rows
                    for f in range(len(views[r])):
                        if up:
                            views[r + 1][f][c0 - up:c0].copy_(views[r][f][a1 - up:a1])
                        if dn:
                            views[r][f][a1:a1 + dn].copy_(views[r + 1][f][c0:c0 + dn])
                torch.cuda.synchronize()
        parts = [s.download(np.float32) for s in slabs]
    finally:
        for s in slabs:
            s.close()
    return np.concatenate([p[0] for p in parts]), np.concatenate([p[1] for p in parts])


class PeerSlab:
    """One rank's row slab with the IN-KERNEL halo exchange (include/hs.h: slab_world / slab_rank,
    hs_slab_export / hs_slab_connect).  This class is only a binding: the plan, the seam wiring (CUDA
    IPC) and the exchange itself (peer stores + per-tile flags inside the fused kernel) live in
    libhs_b200.so.  `all_gather(bytes) -> [bytes per rank]` is the only communication it needs."""

    def __init__(self, height, width, window, iterations, alpha, rank, world, device, temporal_k=0, stream=None):
        from .horn_schunck import Solver
        self.rank, self.world, self.height, self.width = rank, world, height, width
        self.solver = Solver(width, height, window, iterations, alpha, device=device, temporal_k=temporal_k,
                             stream=stream, slab=(rank, world) if world > 1 else None)
        if world > 1:
            info = self.solver.slab_info()
            self.own = (info.own_begin, info.own_end)
            self.frame_rows = (info.frame_begin, info.frame_end)
            self.halo = (info.halo_top, info.halo_bottom)
        else:
            self.own, self.frame_rows, self.halo = (0, height), (0, height), (0, 0)
        self.k = self.solver.timing().temporal_k

    def connect(self, all_gather):
        if self.world == 1:
            return
        handles = all_gather(self.solver.slab_export())
        self.solver.slab_connect(handles[self.rank - 1] if self.rank > 0 else None,
                                 handles[self.rank + 1] if self.rank < self.world - 1 else None)

    def close(self):
        self.solver.close()


def torch_all_gather_bytes(device):
    """all_gather of a fixed-size bytes object through torch.distributed (NCCL needs device tensors)."""
    import torch
    import torch.distributed as dist

    def gather(blob: bytes):
        mine = torch.frombuffer(bytearray(blob), dtype=torch.uint8).to(device)
        parts = [torch.empty_like(mine) for _ in range(dist.get_world_size())]
        dist.all_gather(parts, mine)
        return [bytes(p.cpu().numpy().tobytes()) for p in parts]
    return gather


def peer_slab_solve(prev_rows_fn, height, width, window, iterations, alpha, rank, world, local_rank, temporal_k=0):
    """Solve one image over `world` ranks with the in-kernel exchange; returns this rank's (own rows, u, v)
    as float32.  prev_rows_fn(f0, f1) -> (prev, next) uint8 rows [f0, f1) of the image."""
    import torch
    import torch.distributed as dist
    dev = torch.device("cuda", local_rank)
    s = PeerSlab(height, width, window, iterations, alpha, rank, world, local_rank, temporal_k)
    try:
        s.connect(torch_all_gather_bytes(dev))
        a, b = prev_rows_fn(*s.frame_rows)
        s.solver.upload(a, b)
        s.solver.prepare()
        s.solver.sync()
        if world > 1:
            dist.barrier()                    # every neighbour's planes and seam flags are zeroed
        s.solver.iterate(iterations)
        u, v = s.solver.download(np.float32)
        if world > 1:
            dist.barrier()
        return s.own, u, v, s.k
    finally:
        s.close()


def slab_bit_identity_check(rank, world, local_rank, size=2048, iterations=61, window=3):
    """The multi-GPU result must equal the single-GPU result BIT FOR BIT: every rank solves its slab
    of a size x size pair with the in-kernel exchange, rank 0 also solves the whole image alone."""
    import torch
    import torch.distributed as dist
    from . import synth
    from .horn_schunck import Solver
    own, u, v, k = peer_slab_solve(lambda f0, f1: synth.frame_pair(f1 - f0, size, y0=f0), size, size, window,
                                   iterations, 1.0, rank, world, local_rank)
    dev = torch.device("cuda", local_rank)
    ok = None
    if world == 1:
        return {"size": size, "iterations": iterations, "window": window, "k": k, "bit_identical": True, "note": "one GPU"}
    if rank == 0:
        full_u = np.empty((size, size), np.float32); full_v = np.empty((size, size), np.float32)
        full_u[own[0]:own[1]] = u; full_v[own[0]:own[1]] = v
        y = own[1]
        for r in range(1, world):
            hdr = torch.empty(2, dtype=torch.int64, device=dev)
            dist.recv(hdr, r)
            y0, y1 = int(hdr[0]), int(hdr[1])
            bu = torch.empty((y1 - y0, size), dtype=torch.float32, device=dev); bv = torch.empty_like(bu)
            dist.recv(bu, r); dist.recv(bv, r)
            full_u[y0:y1] = bu.cpu().numpy(); full_v[y0:y1] = bv.cpu().numpy()
            assert y0 == y
            y = y1
        a, b = synth.frame_pair(size, size)
        with Solver(size, size, window, iterations, 1.0, device=local_rank, temporal_k=k) as one:
            ou, ov = one.solve(a, b, np.float32)
        ok = bool(np.array_equal(ou, full_u) and np.array_equal(ov, full_v))
        res = {"size": size, "iterations": iterations, "window": window, "k": k, "bit_identical": ok,
               "max_abs_diff": float(max(np.abs(ou - full_u).max(), np.abs(ov - full_v).max()))}
    else:
        dist.send(torch.tensor(list(own), dtype=torch.int64, device=dev), 0)
        dist.send(torch.from_numpy(u).to(dev), 0); dist.send(torch.from_numpy(v).to(dev), 0)
        res = None
    dist.barrier()
    return res


def bench_slab_record(args, rank, local_rank, world, algo_bytes, peak_fn, size=None, iterations=None, steps=2,
                      exchange="peer", clock_sampler=None):
    """BASELINE configs[4]: ONE 16384 x 16384 pair, 5000 sweeps, row slabs over the ranks (strong scaling).
    Returns the record rank 0 prints (None on other ranks).  exchange = "peer": halos move inside the
    fused kernel (stores into the neighbour's halo rows over NVLink + per-tile flags, one launch per
    GPU for all sweeps); "nccl": one launch per k sweeps and torch.distributed send/recv in between
    (the A/B the north star names)."""
    import torch
    import torch.distributed as dist
    from . import synth
    H = W = int(size or os.environ.get("HS_SLAB_SIZE", 16384))
    T = int(iterations or 5000)
    window = args.window
    rl, rr = radii(window)
    dev = torch.device("cuda", local_rank)
    stream = torch.cuda.Stream(device=dev)

    def barrier():
        if world > 1:
            dist.barrier()

    if exchange == "nccl" and world > 1:
        k = args.k or max(1, 6 // max(1, max(rl, rr)))
        depth = int(os.environ.get("HS_SLAB_DEPTH", 0)) or (3 if world >= 8 else (2 if world >= 4 else 1))
        geom = plan(H, W, world, rank, window, k, depth)
        prev, nxt = synth.frame_pair(geom.f1 - geom.f0, W, y0=geom.f0)
        old = DeviceSlab(geom, T, 1.0, local_rank)
        old.upload(prev, nxt); old.solver.sync()
        run_stream = old.stream

        def run(iters):
            old.iterations = iters
            return old.run() + 1
        halo_rows = (rl * k * depth, rr * k * depth)
        exchanges = (T + k * depth - 1) // (k * depth) - 1
        closer = old.close
        how = f"NCCL send/recv between launches (torch.distributed), halo depth {depth} launches"
    else:
        ps = PeerSlab(H, W, window, T, 1.0, rank, world, local_rank, args.k, stream=stream.cuda_stream)
        ps.connect(torch_all_gather_bytes(dev))
        prev, nxt = synth.frame_pair(ps.frame_rows[1] - ps.frame_rows[0], W, y0=ps.frame_rows[0])
        ps.solver.upload(prev, nxt); ps.solver.sync()
        k = ps.k
        run_stream = stream

        def run(iters):
            ps.solver.prepare()
            if world > 1:
                ps.solver.sync(); dist.barrier()      # neighbours' planes and seam flags are zeroed (inside the timed region)
            ps.solver.iterate(iters)
            return 2
        halo_rows = (rl * k, rr * k)
        exchanges = (T + k - 1) // k
        closer = ps.close
        how = ("in-kernel: seam tiles store their rows into the neighbour's halo rows over NVLink (peer memory, CUDA IPC) "
               "and publish per-tile flags; one launch per GPU for all sweeps, no host work between sweeps")
    del prev, nxt
    flush = torch.empty(512 << 20, dtype=torch.uint8, device=dev)
    for _ in range(max(1, min(args.warmup, 2))):
        run(min(T, 8 * k)); torch.cuda.synchronize(); barrier()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
    launches = 0
    if clock_sampler is not None:
        clock_sampler.mark()
    barrier(); torch.cuda.synchronize()
    for s, e in ev:
        with torch.cuda.stream(run_stream):
            flush.zero_()
        torch.cuda.synchronize(); barrier()
        s.record(run_stream)
        launches += run(T)
        e.record(run_stream)
        torch.cuda.synchronize(); barrier()
    total = torch.tensor([sum(s.elapsed_time(e) for s, e in ev)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(total, op=dist.ReduceOp.MAX)
    total_s = float(total.item()) / 1e3
    closer()
    del flush
    torch.cuda.empty_cache()
    if rank != 0:
        return None
    value = float(H) * W * T * steps / total_s / 1e6
    peak, src = peak_fn()
    achieved = algo_bytes * H * W * T * steps / total_s / 1e9 / world     # per GPU
    return {"value": value, "unit": "Mpixel-iter/s", "ms_per_step": total_s / steps * 1e3, "steps": steps,
            "scaling": "strong", "n_gpus": world,
            "workload": f"configs[4]: one {W}x{H} pair, {T} sweeps, {world} row slab(s)", "window": window,
            "iterations": T, "temporal_k": k, "exchange": how,
            "halo_rows_per_exchange": {"from_above": halo_rows[0], "from_below": halo_rows[1]},
            "halo_bytes_per_exchange_per_seam": int((halo_rows[0] + halo_rows[1]) * W * 4 * 2),
            "exchanges_per_step": exchanges if world > 1 else 0,
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "peak_source": src, "note": "per GPU, 32 algorithmic bytes per pixel-iteration, halo exchange included"},
            "gpu_launches": launches,
            "l2": "planes (GBs per GPU) exceed L2; 512 MiB memset before every step anyway"}


