"""Multi-GPU row-slab correctness (run under torchrun, one rank per GPU): every rank solves its slab;
rank 0 also solves the whole image on its GPU and the stitched result must be BIT-IDENTICAL.
Default: the in-kernel exchange behind the C ABI (hs_slab_export / hs_slab_connect); HS_SLAB_NCCL=1:
the older host-driven path with NCCL send/recv between launches.  Prints one JSON line per case.
   torchrun --nproc-per-node N tools/slab_check.py [H W T window k] ..."""
import json, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import torch.distributed as dist
import cpp_optical_flow_b200 as P
from cpp_optical_flow_b200 import slab, synth

rank, local, world = int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
dev = torch.device("cuda", local)
vals = [int(x) for x in sys.argv[1:]] or [4096, 4096, 51, 3, 6]
cases = [vals[i:i + 5] for i in range(0, len(vals), 5)]
for H, W, T, w, k in cases:
    if os.environ.get("HS_SLAB_NCCL"):
        depth = int(os.environ.get("HS_SLAB_DEPTH", 1))
        g = slab.plan(H, W, world, rank, w, k, depth)
        prev, nxt = synth.frame_pair(g.f1 - g.f0, W, y0=g.f0)
        s = slab.DeviceSlab(g, T, 1.0, local)
        s.upload(prev, nxt); s.run(); torch.cuda.synchronize()
        u, v = s.download(np.float32); s.close()
        own, kk, how = (g.y0, g.y1), k, f"nccl depth {depth}"
    else:
        own, u, v, kk = slab.peer_slab_solve(lambda f0, f1: synth.frame_pair(f1 - f0, W, y0=f0), H, W, w, T, 1.0,
                                             rank, world, local, temporal_k=k)
        how = "in-kernel peer exchange"
    if rank == 0:
        full_u = np.empty((H, W), np.float32); full_v = np.empty((H, W), np.float32)
        full_u[own[0]:own[1]] = u; full_v[own[0]:own[1]] = v
        for r in range(1, world):
            hdr = torch.empty(2, dtype=torch.int64, device=dev); dist.recv(hdr, r)
            y0, y1 = int(hdr[0]), int(hdr[1])
            bu = torch.empty((y1 - y0, W), dtype=torch.float32, device=dev); bv = torch.empty_like(bu)
            dist.recv(bu, r); dist.recv(bv, r)
            full_u[y0:y1] = bu.cpu().numpy(); full_v[y0:y1] = bv.cpu().numpy()
        a, b = synth.frame_pair(H, W)
        with P.Solver(W, H, w, T, 1.0, device=local, temporal_k=kk) as one:
            ou, ov = one.solve(a, b, np.float32)
        print(json.dumps({"world": world, "H": H, "W": W, "T": T, "window": w, "k": kk, "exchange": how,
                          "bit_identical": bool(np.array_equal(ou, full_u) and np.array_equal(ov, full_v)),
                          "max_abs_diff": float(max(np.abs(ou - full_u).max(), np.abs(ov - full_v).max())),
                          "median_u": float(np.median(full_u))}), flush=True)
    else:
        dist.send(torch.tensor(list(own), dtype=torch.int64, device=dev), 0)
        dist.send(torch.from_numpy(u).to(dev), 0); dist.send(torch.from_numpy(v).to(dev), 0)
    dist.barrier()
dist.destroy_process_group()
