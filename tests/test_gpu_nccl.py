"""Multi-GPU parity through the native NCCL communicator (include/ygg_b200_comm.h): one process per
GPU, world 2, both sharding modes; trees and predictions must be bit-identical to a single-GPU run.
Needs two GPUs (skipped otherwise); the host-side logic is covered on CPU by test_multi_rank_cpu.py and
the collectives' data path on one GPU by test_gpu_sharding.py."""
import os
import socket

import numpy as np
import pytest

import ydf_b200
from tests.util import synth_mixed

pytestmark = pytest.mark.gpu


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _data():
    return synth_mixed(60000, 6, [5, 70], seed=41)


def _worker(rank, world, port, mode, uid_path, q):
    import torch
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("gloo", rank=rank, world_size=world)  # bootstrap only; the data path is NCCL from C++
    try:
        comm = ydf_b200.Comm.from_torch_distributed(rank)
        bins, nb, na, ft, y = _data()
        n, f, iters = bins.shape[1], bins.shape[0], 4
        r0, r1 = ((n * rank) // world, (n * (rank + 1)) // world) if mode.startswith("rows") else (0, n)
        ds = ydf_b200.Dataset(bins[:, r0:r1], nb, na, device=rank, feature_types=ft)
        gbt = ydf_b200.Gbt(ds, ydf_b200.default_config(num_trees=iters, max_depth=6))
        gbt.set_labels(y[r0:r1])
        ratio = float((y == 2).mean())
        init = float(np.float32(np.log(ratio / (1 - ratio))))
        if mode == "rows":
            gbt.set_row_shard(rank, world, n, init, comm)
        elif mode.startswith("rows_scatter"):
            gbt.set_row_shard_scatter(rank, world, n, init, comm)
        else:
            b, e = ydf_b200.feature_shard(f, rank, world)
            gbt.set_feature_shard(b, e, rank, world, comm)
        if mode.endswith("_p2p"):   # best splits exchanged by k_select_global over NVLink peer memory, no all-gather
            gbt.use_peer_windows(comm)
        gbt.train(iters)
        q.put((rank, r0, r1, [gbt.get_tree(i).tobytes() for i in range(iters)], gbt.get_predictions()))
        dist.barrier()
        gbt.close()
        comm.close()
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("mode", ["rows", "rows_scatter", "features", "rows_scatter_p2p", "features_p2p"])
def test_two_gpus_match_one(mode):
    if ydf_b200.device_count() < 2:
        pytest.skip("needs two GPUs")
    import torch.multiprocessing as mp
    bins, nb, na, ft, y = _data()
    ds = ydf_b200.Dataset(bins, nb, na, feature_types=ft)
    gbt = ydf_b200.Gbt(ds, ydf_b200.default_config(num_trees=4, max_depth=6))
    gbt.set_labels(y)
    gbt.train(4)
    want_trees = [gbt.get_tree(i).tobytes() for i in range(4)]
    want_pred = gbt.get_predictions()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, mode, None, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=240) for _ in range(2)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, r0, r1, trees, pred in res:
        assert trees == want_trees, f"rank {rank}: trees differ from the single-GPU run"
        np.testing.assert_array_equal(pred, want_pred[r0:r1])
