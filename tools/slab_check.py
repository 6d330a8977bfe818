"""Multi-GPU row-slab correctness (run under torchrun, one rank per GPU): every rank solves its
slab with NCCL halo exchange; rank 0 also solves the whole image on its GPU and the stitched
result must be BIT-IDENTICAL.  Prints one JSON line.
   torchrun --nproc-per-node N tools/slab_check.py [H W T window k]"""
import json, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import torch.distributed as dist
import cpp_optical_flow_b200 as P
from cpp_optical_flow_b200 import slab, synth

H, W, T, w, k = [int(x) for x in (sys.argv[1:6] + ["4096", "4096", "51", "3", "6"][len(sys.argv) - 1:])]
depth = int(os.environ.get("HS_SLAB_DEPTH", 1))
rank, local, world = int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
g = slab.plan(H, W, world, rank, w, k, depth)
prev, nxt = synth.frame_pair(g.f1 - g.f0, W, y0=g.f0)
s = slab.DeviceSlab(g, T, 1.0, local)
s.upload(prev, nxt)
s.run()
torch.cuda.synchronize()
u, v = s.download(np.float32)
s.close()
tu = torch.from_numpy(u).cuda(); tv = torch.from_numpy(v).cuda()
if rank == 0:
    full_u = np.empty((H, W), np.float32); full_v = np.empty((H, W), np.float32)
    full_u[g.y0:g.y1] = u; full_v[g.y0:g.y1] = v
    for r in range(1, world):
        gr = slab.plan(H, W, world, r, w, k, depth)
        bu = torch.empty((gr.y1 - gr.y0, W), dtype=torch.float32, device="cuda"); bv = torch.empty_like(bu)
        dist.recv(bu, r); dist.recv(bv, r)
        full_u[gr.y0:gr.y1] = bu.cpu().numpy(); full_v[gr.y0:gr.y1] = bv.cpu().numpy()
    a, b = synth.frame_pair(H, W)
    with P.Solver(W, H, w, T, 1.0, device=local, temporal_k=k) as one:
        ou, ov = one.solve(a, b, np.float32)
    print(json.dumps({"world": world, "H": H, "W": W, "T": T, "window": w, "k": k, "depth": depth,
                      "bit_identical": bool(np.array_equal(ou, full_u) and np.array_equal(ov, full_v)),
                      "max_abs_diff": float(max(np.abs(ou - full_u).max(), np.abs(ov - full_v).max())),
                      "median_u": float(np.median(full_u))}), flush=True)
else:
    dist.send(tu, 0); dist.send(tv, 0)
dist.barrier()
dist.destroy_process_group()
