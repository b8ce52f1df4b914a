"""Multi-GPU parity check, one process per GPU (launch with torch.distributed.run / torchrun):

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 tests/multi_gpu_check.py

Rank g processes LiDAR g's sweep (its extrinsic set with mloam_set_extrinsic) against the replicated submap; the packed
normal equations are all-reduced over NCCL inside the library every LM evaluation.  Every rank must end with the same
pose, equal (<= 1e-4 m / rad) to the oracle run on the concatenation of the per-LiDAR down-sampled features."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests"))
sys.path.insert(0, ROOT)


def main():
    import torch
    import torch.distributed as dist

    import oracle_lib as orc
    import synthetic as syn
    from bench import lidar_extrinsic, load_mloam

    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("gloo")
    m = load_mloam()
    p = m.default_params()
    p.n_scans, p.max_outer, p.max_inner, p.map_cell = 16, 5, 1, 0.5
    ctx = m.Context(local, p)
    uid = [m.Context.comm_unique_id() if rank == 0 else None]
    dist.broadcast_object_list(uid, src=0)
    ctx.comm_init(world, rank, uid[0])
    mode = os.environ.get("MLOAM_EXCHANGE", "p2p")
    if mode == "p2p":  # peer-memory exchange inside the k_linearize tail; MLOAM_EXCHANGE=nccl checks the all-reduce path
        handles = [None] * world
        dist.all_gather_object(handles, ctx.comm_p2p_export())
        ctx.comm_p2p_init(world, rank, handles)

    scene = syn.make_scene()
    truth = syn.trajectory(3)[2]
    surf_map, corner_map = syn.make_submap(scene, 60000)
    init = syn.perturb_pose(truth, np.random.Generator(np.random.PCG64(77)))
    exts = [lidar_extrinsic(syn, r, max(world, 2)) for r in range(world)]
    clouds = [syn.make_sweep(scene, truth, 16, 1024, seed=50, lidar_id=r, ext=exts[r]) for r in range(world)]
    ctx.set_extrinsic(exts[rank])
    cloud, ss, se = clouds[rank]
    pose, st = ctx.frame(cloud, ss, se, surf_map, corner_map, init)

    poses = [None] * world
    dist.all_gather_object(poses, pose.tolist())
    ok = True
    if rank == 0:
        for r in range(1, world):
            if not np.array_equal(np.array(poses[r]), np.array(poses[0])):
                print(f"FAIL: rank {r} pose differs from rank 0")
                ok = False
        # oracle: per-LiDAR extraction -> base frame -> per-LiDAR down-sampling -> merged features -> scan2map
        cs_all, sf_all = [], []
        for r in range(world):
            f = orc.extract_cloud(*clouds[r])
            less_sharp, less_flat = f["corner_points_less_sharp"], f["surf_points_less_flat"]
            if exts[r] is not None:
                less_sharp, less_flat = orc.associate(less_sharp, exts[r]), orc.associate(less_flat, exts[r])
            cs_all.append(orc.voxel_grid(less_sharp, 0.2, True)[0])
            sf_all.append(orc.voxel_grid(less_flat, 0.4, True)[0])
        o = orc.default_opts()
        o[orc.O_MAX_OUTER], o[orc.O_MAX_INNER] = 5, 1
        ref, rst = orc.scan2map(surf_map, corner_map, np.concatenate(sf_all), np.concatenate(cs_all), init, o)
        dt, dr = syn.pose_err(pose, ref)
        print(f"multi-gpu x{world} [{mode}]: pose vs merged-feature oracle dt={dt:.3e} m dr={dr:.3e} rad; "
              f"matches {st['n_surf']}+{st['n_corner']} (this rank's reduced count) vs oracle {int(rst['n_surf'])}+{int(rst['n_corner'])}")
        if not (dt <= 1e-4 and dr <= 1e-4):
            print("FAIL: pose parity")
            ok = False
        if st["n_surf"] != int(rst["n_surf"]) or st["n_corner"] != int(rst["n_corner"]):
            print("FAIL: reduced match counts differ from the oracle")
            ok = False
    flag = [ok]
    dist.broadcast_object_list(flag, src=0)
    ctx.close()
    dist.destroy_process_group()
    if rank == 0:
        print("MULTI_GPU_CHECK", "OK" if flag[0] else "FAILED")
    sys.exit(0 if flag[0] else 1)


if __name__ == "__main__":
    main()
