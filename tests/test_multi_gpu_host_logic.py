"""N > 1 path on CPU: world_size-2 gloo run of the keyframe sharding and the
final gather (rpg_open_remode_b200/multi_gpu.py).  The per-rank maps are
produced by the CPU oracle here (the product has no CPU path); on GPUs the same
functions run over NCCL in bench.py."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _keyframe_maps(keyframe, n_frames=3, w=96, h=72):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import oracle_binding as ob
    from rpg_open_remode_b200 import multi_gpu, synth
    seq = synth.SyntheticSequence(w, h, seed=multi_gpu.keyframe_seed(keyframe))
    f0 = seq.frame(0)
    o = ob.OracleSeeds(w, h, *seq.camera, patch=5)
    o.set_reference(f0.image, f0.T_cam_world, float(f0.depth.min()), float(f0.depth.max()))
    for k in range(1, n_frames + 1):
        f = seq.frame(k, want_depth=False)
        o.update(f.image, f.T_cam_world)
    return o.mu.copy(), o.convergence.copy()


def _worker(rank, world, port, n_keyframes, out_dir):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    sys.path.insert(0, ROOT)
    from rpg_open_remode_b200 import multi_gpu
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        mine = multi_gpu.shard_keyframes(n_keyframes, rank, world)
        shards = [multi_gpu.shard_keyframes(n_keyframes, r, world) for r in range(world)]
        results = {}
        rounds = max(len(s) for s in shards)
        for i in range(rounds):
            kf = mine[i] if i < len(mine) else mine[-1]  # every rank takes part in every gather
            mu, conv = _keyframe_maps(kf)
            got = multi_gpu.gather_maps(torch.from_numpy(mu), torch.from_numpy(conv), dst=0)
            if rank == 0:
                depths, convs = got
                round_kfs = [[s[i]] if i < len(s) else [] for s in shards]
                for k, v in multi_gpu.assemble(round_kfs, depths).items():
                    results[k] = (v.numpy(), multi_gpu.assemble(round_kfs, convs)[k].numpy())
            else:
                assert got is None
        # the single-collective gatherer of the bench loop gives the same maps
        mu, conv = _keyframe_maps(mine[0])
        mg = multi_gpu.MapGatherer(mu.shape[0], mu.shape[1])
        for _ in range(2):   # reusable
            mg.depth.copy_(torch.from_numpy(mu))
            mg.convergence.copy_(torch.from_numpy(conv))
            got = mg.gather()
            if rank == 0:
                for r in range(world):
                    mu_r, conv_r = _keyframe_maps(shards[r][0])
                    assert np.array_equal(got[0][r].numpy(), mu_r) and np.array_equal(got[1][r].numpy(), conv_r)
            else:
                assert got is None
        t = multi_gpu.max_over_ranks(float(rank + 1))
        assert t == float(world)
        if rank == 0:
            np.savez(os.path.join(out_dir, "gathered.npz"),
                     **{"mu%d" % k: v[0] for k, v in results.items()},
                     **{"conv%d" % k: v[1] for k, v in results.items()})
    finally:
        dist.destroy_process_group()


def test_shard_keyframes_partition():
    from rpg_open_remode_b200 import multi_gpu
    for world in (1, 2, 4, 8):
        for n in (1, 7, 8, 9):
            shards = [multi_gpu.shard_keyframes(n, r, world) for r in range(world)]
            flat = sorted(k for s in shards for k in s)
            assert flat == list(range(n))
            assert max(len(s) for s in shards) - min(len(s) for s in shards) <= 1
    with pytest.raises(ValueError):
        multi_gpu.shard_keyframes(4, 2, 2)
    assert len({multi_gpu.keyframe_seed(k) for k in range(8)}) == 8


def test_gather_world_size_2_gloo(tmp_path):
    n_keyframes, world = 3, 2
    port = _free_port()
    mp.start_processes(_worker, args=(world, port, n_keyframes, str(tmp_path)), nprocs=world, join=True,
                       start_method="spawn")
    data = np.load(os.path.join(str(tmp_path), "gathered.npz"))
    for kf in range(n_keyframes):
        mu, conv = _keyframe_maps(kf)
        assert np.array_equal(data["mu%d" % kf], mu), f"keyframe {kf}: depth map differs after the gather"
        assert np.array_equal(data["conv%d" % kf], conv)


def test_gather_without_process_group():
    from rpg_open_remode_b200 import multi_gpu
    d, c = torch.zeros(4, 5), torch.zeros(4, 5, dtype=torch.int32)
    depths, convs = multi_gpu.gather_maps(d, c)
    assert depths[0] is d and convs[0] is c
    assert multi_gpu.max_over_ranks(1.5) == 1.5
    mg = multi_gpu.MapGatherer(4, 5)
    mg.depth.fill_(1.25)
    mg.convergence.fill_(3)
    depths, convs = mg.gather()
    assert depths[0].dtype == torch.float32 and float(depths[0][2, 2]) == 1.25 and int(convs[0][3, 4]) == 3
    assert depths[0].data_ptr() == mg.packed.data_ptr()      # views of the one buffer: no copy
