"""World-size-2 (and 3) gloo tests of the N>1 host logic: the shard layout and the ordered merge of
per-shard best splits that every rank performs after the per-level all-gather.  The CUDA path runs
the same `merge_shard_bests` function inside k_select_global."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import ydf_b200


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _scores(seed, nodes, features):
    rng = np.random.default_rng(seed)
    sc = rng.random((nodes, features)).astype(np.float32)
    sc[rng.random((nodes, features)) < 0.3] = 0.0          # no valid split for that (node, feature)
    sc[:, features // 2] = sc[:, 1]                          # exact ties across shards
    return sc


def _local_best(sc, begin, end):
    nodes = sc.shape[0]
    out = np.zeros(nodes, dtype=ydf_b200.SHARD_BEST_DTYPE)
    out["feature"] = -1
    for j in range(nodes):
        best = np.float32(0)
        for f in range(begin, end):
            if sc[j, f] > best:
                best = sc[j, f]
                cat = f % 3 == 0  # every third feature is categorical: the mask must survive the exchange
                out[j] = (sc[j, f], f, 0 if cat else 7 + f, 100 + f, int(cat), f & 1,
                          tuple((j * 2654435761 + f * 40503 + w) & 0xFFFFFFFF if cat else 0 for w in range(8)))
    return out


def _worker(rank, world, port, nodes, features, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    sc = _scores(123, nodes, features)
    b, e = ydf_b200.feature_shard(features, rank, world)
    mine = _local_best(sc, b, e)
    send = torch.from_numpy(mine.view(np.uint8).copy())
    recv = torch.empty(world * send.numel(), dtype=torch.uint8)
    dist.all_gather_into_tensor(recv, send)
    records = recv.numpy().view(ydf_b200.SHARD_BEST_DTYPE).reshape(world, nodes)
    merged = ydf_b200.merge_shard_best(records)
    q.put((rank, b, e, merged.tobytes()))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_sharded_merge_equals_single_rank(world):
    nodes, features = 37, 11
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, nodes, features, q)) for r in range(world)]
    for p in procs:
        p.start()
    results = [q.get(timeout=120) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    # shard layout: contiguous, ascending, covering [0, F)
    bounds = sorted((b, e) for _, b, e, _ in results)
    assert bounds[0][0] == 0 and bounds[-1][1] == features
    assert all(bounds[i][1] == bounds[i + 1][0] for i in range(world - 1))
    # every rank computes the same merge, equal to the single-rank fold over all features
    want = _local_best(_scores(123, nodes, features), 0, features)
    for _, _, _, blob in results:
        got = np.frombuffer(blob, dtype=ydf_b200.SHARD_BEST_DTYPE)
        np.testing.assert_array_equal(got["feature"], want["feature"])
        np.testing.assert_array_equal(got["score"], want["score"])
        np.testing.assert_array_equal(got["threshold_bin"], want["threshold_bin"])


def test_shard_argument_errors():
    with pytest.raises(ydf_b200.YggError):
        ydf_b200.feature_shard(3, 0, 8)   # fewer features than ranks
    with pytest.raises(ydf_b200.YggError):
        ydf_b200.feature_shard(10, 4, 4)
    assert ydf_b200.feature_shard(200, 7, 8) == (175, 200)
