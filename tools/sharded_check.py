#!/usr/bin/env python3
"""Real-NCCL check of the sharded paths (run under torchrun, one rank per GPU):

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 \
        tools/sharded_check.py [Z,H,W]

Every rank reconstructs its z-slab (occupancy path: staged, fused, fused+graph; SDF path incl. the EDT all-to-all
transpose); rank 0 gathers the stitched mesh and compares it bit for bit with its own single-GPU run on the whole stack.
Prints one line per check and exits non-zero on a mismatch."""
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402


def main():
    rank, local_rank, world = int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"]), int(os.environ["WORLD_SIZE"])
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    os.environ.setdefault("NCCL_DEBUG", "WARN")
    dist.init_process_group("nccl", device_id=dev)
    from tomography_3d_reconstructor_b200 import pipeline, sharded
    Z, H, W = (int(v) for v in (sys.argv[1] if len(sys.argv) > 1 else "96,256,320").split(","))
    sides = bench.side_counts(Z)
    phys = (bench.PHYS["total_depth_mm"], bench.PHYS["x_length_mm"], bench.PHYS["y_length_mm"])
    z0, z1 = sharded.slab_range(Z, rank, world)
    masks = bench.make_phantom_u8(Z, H, W, z0, z1, dev)
    masks[:, H // 2:H // 2 + 3, W // 3:W // 3 + 5] = 0          # a tunnel through every slab boundary
    ok_all = True

    def compare(name, out, ref):
        nonlocal ok_all
        m = sharded.gather_mesh(out, 0)
        ok = True
        if rank == 0:
            v, f = m
            rv, rf = ref["mesh"].verts, ref["mesh"].faces
            ok = v.shape == rv.shape and f.shape == rf.shape and bool(torch.equal(v.view(torch.int32), rv.view(torch.int32))) \
                and bool(torch.equal(f, rf)) and out["voxel_volume_mm3"] == ref["voxel_volume_mm3"] \
                and out["processed_voxel_volume_mm3"] == ref["processed_voxel_volume_mm3"] and out["bbox_index"] == ref["bbox_index"] \
                and abs(out["mesh_volume_mm3"] - ref["mesh_volume_mm3"]) <= 1e-9 * ref["mesh_volume_mm3"] and out["stitch_consistent"]
            print("%-28s %s  V=%d F=%d" % (name, "OK" if ok else "MISMATCH", v.shape[0], f.shape[0]), flush=True)
        ok_all = ok_all and ok

    ref = ref_sdf = None
    if rank == 0:
        full = bench.make_phantom_u8(Z, H, W, 0, Z, dev)
        full[:, H // 2:H // 2 + 3, W // 3:W // 3 + 5] = 0
        ref = pipeline.reconstruct(full, bench.THRESHOLD, sides, *phys)
        ref_sdf = pipeline.reconstruct_sdf(full, bench.THRESHOLD, sides, *phys)
    compare("occupancy staged", sharded.reconstruct(masks, Z, z0, bench.THRESHOLD, sides, *phys), ref)
    for rep in range(3):
        compare("occupancy fused #%d" % rep, sharded.reconstruct_fused(masks, Z, z0, bench.THRESHOLD, sides, *phys, use_graph=False), ref)
    sharded._slab_plans.clear()
    for rep in range(3):
        compare("occupancy fused+graph #%d" % rep, sharded.reconstruct_fused(masks, Z, z0, bench.THRESHOLD, sides, *phys, use_graph=True),
                ref)
    out = sharded.reconstruct_sdf(masks, Z, z0, bench.THRESHOLD, sides, *phys)
    compare("sdf (sharded EDT)", out, ref_sdf)
    if rank == 0:
        same = bool(torch.equal(out["sdf"], ref_sdf["sdf"][z0:z1]))
        print("%-28s %s" % ("sdf field of rank 0", "OK" if same else "MISMATCH"), flush=True)
        ok_all = ok_all and same
    flag = torch.tensor([0 if ok_all else 1], device=dev)
    dist.all_reduce(flag)
    sharded._slab_plans.clear()
    import gc
    gc.collect()
    torch.cuda.synchronize()
    dist.barrier()
    dist.destroy_process_group()
    sys.exit(1 if int(flag.item()) else 0)


if __name__ == "__main__":
    main()
