"""CPU tests of the multi-GPU host logic: the row-slab plan and the halo exchange run for real over
torch.distributed (gloo, world_size 2 and 3, 127.0.0.1) with a NumPy stand-in for the device
kernels, and the stitched result must equal the whole-image oracle bit for bit."""
import os
import sys
import tempfile

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_partition_and_plan_cover_the_image():
    from cpp_optical_flow_b200 import slab
    for H, world in ((16384, 8), (1080, 7), (100, 3), (9, 2)):
        b = slab.partition_rows(H, world)
        assert b[0][0] == 0 and b[-1][1] == H and all(b[i][1] == b[i + 1][0] for i in range(world - 1))
        assert max(y1 - y0 for y0, y1 in b) - min(y1 - y0 for y0, y1 in b) <= 1
    gd = slab.plan(1000, 64, 4, 1, 3, 4, depth=2)        # deep halo: 2 launches of k=4 between exchanges
    assert (gd.b0, gd.b1) == (242, 508) and slab.cycle_rows(gd, 0, 2) == (4, 262) and slab.cycle_rows(gd, 1, 2) == (8, 258)
    g = slab.plan(1000, 64, 4, 1, 5, 3)                  # w=5: 2 taps each side, k=3 -> 6 halo rows
    assert (g.y0, g.y1) == (250, 500) and (g.b0, g.b1) == (244, 506) and (g.f0, g.f1) == (243, 507)
    assert g.top_seam and g.bottom_seam and g.out_rows == (6, 256)
    g0 = slab.plan(1000, 64, 4, 0, 4, 2)                 # w=4: anchor 1, 2 taps below
    assert not g0.top_seam and g0.bottom_seam and (g0.b0, g0.b1) == (0, 254) and g0.f0 == 0 and g0.f1 == 255
    g3 = slab.plan(1000, 64, 4, 3, 4, 2)
    assert (g3.b0, g3.b1, g3.f0, g3.f1) == (748, 1000, 747, 1000)
    with pytest.raises(ValueError):
        slab.plan(40, 64, 8, 0, 5, 4)                    # 5-row slabs cannot feed an 8-row halo


def _worker(rank, world, port, w, k, iters, shape, outdir, depth):
    import torch
    import torch.distributed as dist
    sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import hs_oracle as O
    from cpp_optical_flow_b200 import slab
    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    rng = np.random.default_rng(11)
    prev = rng.integers(0, 256, shape, dtype=np.uint8)
    nxt = rng.integers(0, 256, shape, dtype=np.uint8)
    g = slab.plan(shape[0], shape[1], world, rank, w, k, depth)
    # coefficient stage on this rank's frame rows only (seam row included), cropped to the buffer
    gx, gy, gt = O.np_gradients(prev[g.f0:g.f1], nxt[g.f0:g.f1])
    c = slice(g.b0 - g.f0, g.b0 - g.f0 + g.rows)
    gx, gy, gt = gx[c], gy[c], gt[c]
    den = 1.0 + gx * gx + gy * gy
    u = torch.zeros((g.rows, shape[1]), dtype=torch.float64)
    v = torch.zeros_like(u)
    o0, o1 = g.out_rows
    left = iters
    while left > 0:
        cyc = min(depth, (left + k - 1) // k)
        for j in range(cyc):
            kk = min(k, left)
            r0, r1 = slab.cycle_rows(g, j, cyc)
            un, vn = u.numpy().copy(), v.numpy().copy()
            for _ in range(kk):                           # kk sweeps on the whole buffer, zero outside
                ua, va = O.np_box(un, w), O.np_box(vn, w)
                cc = (gx * ua + gy * va + gt) / den
                un, vn = ua - gx * cc, va - gy * cc
            u[r0:r1] = torch.from_numpy(un[r0:r1])        # only rows r0..r1 are exact (like the kernel's launch)
            v[r0:r1] = torch.from_numpy(vn[r0:r1])
            left -= kk
        if left > 0:
            slab.exchange_halos(g, [u, v])
    np.save(os.path.join(outdir, f"u{rank}.npy"), u[o0:o1].numpy())
    np.save(os.path.join(outdir, f"v{rank}.npy"), v[o0:o1].numpy())
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("world,w,k,iters,depth", [(2, 3, 4, 11, 1), (2, 5, 2, 7, 1), (3, 4, 2, 6, 1), (2, 2, 3, 7, 1),
                                                   (2, 3, 2, 11, 2), (3, 5, 1, 9, 3)])
def test_slab_exchange_over_gloo_equals_whole_image(world, w, k, iters, depth):
    import torch.multiprocessing as mp
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import hs_oracle as O
    shape = (61, 40)
    port = 29500 + (os.getpid() + world * 7 + w + 13 * depth) % 2000
    with tempfile.TemporaryDirectory() as d:
        mp.spawn(_worker, args=(world, port, w, k, iters, shape, d, depth), nprocs=world, join=True)
        u = np.concatenate([np.load(os.path.join(d, f"u{r}.npy")) for r in range(world)])
        v = np.concatenate([np.load(os.path.join(d, f"v{r}.npy")) for r in range(world)])
    rng = np.random.default_rng(11)
    prev = rng.integers(0, 256, shape, dtype=np.uint8)
    nxt = rng.integers(0, 256, shape, dtype=np.uint8)
    *_, ou, ov = O.np_flow(prev, nxt, w, iters, 1.0)
    assert np.array_equal(u, ou) and np.array_equal(v, ov)
