"""Multi-rank decompositions exercised on ONE GPU with an in-process fake communicator (two engine
handles driven by two threads; the collectives are emulated with torch ops on the same device) —
the same idea as the reference's MULTI_THREAD distribute backend for testing without a cluster
(utils/distribute/implementations/multi_thread).  Both decompositions must reproduce the
single-rank trees bit for bit:
  * rows     : integer level histograms all-reduced (exact, order independent);
  * features : best-split records all-gathered and merged in rank order."""
import threading

import numpy as np
import pytest
import torch

import ydf_b200
from tests.util import synth, synth_mixed

pytestmark = pytest.mark.gpu


class _Buf:
    def __init__(self, p, n, typestr):
        self.__cuda_array_interface__ = {"shape": (n,), "typestr": typestr, "data": (p, False), "version": 2}


class FakeComm:
    def __init__(self, world):
        self.world = world
        self.barrier = threading.Barrier(world, timeout=60)  # a failed rank must not hang the others
        self.slots = [None] * world

    def _rendezvous(self, rank, item, stream):
        torch.cuda.ExternalStream(stream).synchronize()
        self.slots[rank] = item
        self.barrier.wait()

    def allreduce(self, rank):
        def fn(buf, count, dtype, op, stream):
            try:
                return body(buf, count, dtype, op, stream)
            except threading.BrokenBarrierError:
                return 1

        def body(buf, count, dtype, op, stream):
            typestr = {0: "<i4", 1: "<i8", 2: "<f8"}[dtype]
            self._rendezvous(rank, (buf, count, typestr), stream)
            if rank == 0:
                ts = [torch.as_tensor(_Buf(*s), device="cuda:0") for s in self.slots]
                acc = ts[0].clone()
                for t in ts[1:]:
                    acc = acc + t if op == 0 else torch.maximum(acc, t)
                for t in ts:
                    t.copy_(acc)
                torch.cuda.synchronize()
            self.barrier.wait()
            return 0
        return fn

    def reducescatter(self, rank):
        """In place: world * count elements per rank, rank r keeps the sum of everybody's chunk r."""
        def fn(buf, count, dtype, op, stream):
            try:
                typestr = {0: "<i4", 1: "<i8", 2: "<f8"}[dtype]
                self._rendezvous(rank, (buf, count * self.world, typestr), stream)
                if rank == 0:
                    ts = [torch.as_tensor(_Buf(*s), device="cuda:0") for s in self.slots]
                    acc = ts[0].clone()
                    for t in ts[1:]:
                        acc = acc + t if op == 0 else torch.maximum(acc, t)
                    for r, t in enumerate(ts):
                        t[r * count:(r + 1) * count].copy_(acc[r * count:(r + 1) * count])
                        # the other chunks are left as garbage on purpose: nothing may read them
                        mask = torch.ones_like(t, dtype=torch.bool)
                        mask[r * count:(r + 1) * count] = False
                        t[mask] = -1 if typestr != "<f8" else float("nan")
                    torch.cuda.synchronize()
                self.barrier.wait()
                return 0
            except threading.BrokenBarrierError:
                return 1
        return fn

    def allgather(self, rank):
        def fn(send, recv, nbytes, stream):
            try:
                return body(send, recv, nbytes, stream)
            except threading.BrokenBarrierError:
                return 1

        def body(send, recv, nbytes, stream):
            self._rendezvous(rank, (send, recv, nbytes), stream)
            if rank == 0:
                parts = [torch.as_tensor(_Buf(s[0], s[2], "|u1"), device="cuda:0").clone() for s in self.slots]
                for s in self.slots:
                    out = torch.as_tensor(_Buf(s[1], s[2] * self.world, "|u1"), device="cuda:0")
                    for r, p in enumerate(parts):
                        out[r * s[2]:(r + 1) * s[2]].copy_(p)
                torch.cuda.synchronize()
            self.barrier.wait()
            return 0
        return fn


def _single(bins, nb, na, y, iters, ft=None, **kw):
    ds = ydf_b200.Dataset(bins, nb, na, feature_types=ft)
    gbt = ydf_b200.Gbt(ds, ydf_b200.default_config(num_trees=iters, **kw))
    gbt.set_labels(y)
    gbt.train(iters)
    return ([gbt.get_tree(i).tobytes() for i in range(iters)], [gbt.train_loss(i) for i in range(iters)],
            gbt.get_predictions(), gbt.initial_prediction())


def _run_ranks(world, target, comm=None):
    out, errs = [None] * world, []

    def wrap(r):
        try:
            out[r] = target(r)
        except Exception as e:  # surface worker failures in the main thread
            errs.append(e)
            if comm is not None:
                comm.barrier.abort()

    ts = [threading.Thread(target=wrap, args=(r,), daemon=True) for r in range(world)]
    for t in ts:
        t.start()
    for t in ts:
        t.join(timeout=300)
    assert not errs, errs
    return out


@pytest.mark.parametrize("loss,hess,world", [(0, 0, 2), (0, 1, 2), (1, 0, 3)])
def test_row_shard_matches_single_rank(loss, hess, world):
    n, f, iters = 50000, 9, 6
    bins, nb, na, y = synth(n, f, seed=5, task="binary" if loss == 0 else "regression", bins=64)
    kw = dict(loss=loss, use_hessian_gain=hess, max_depth=6)
    want_trees, want_loss, want_pred, init = _single(bins, nb, na, y, iters, **kw)
    comm = FakeComm(world)

    def rank_main(r):
        r0, r1 = (n * r) // world, (n * (r + 1)) // world
        ds = ydf_b200.Dataset(bins[:, r0:r1], nb, na)   # row slice taken in place (column stride = n)
        gbt = ydf_b200.Gbt(ds, ydf_b200.default_config(num_trees=iters, **kw))
        gbt.set_labels(y[r0:r1])
        gbt.set_row_shard(r, world, n, init, comm.allreduce(r))
        gbt.train(iters)
        return ([gbt.get_tree(i).tobytes() for i in range(iters)], [gbt.train_loss(i) for i in range(iters)],
                gbt.get_predictions())

    res = _run_ranks(world, rank_main, comm)
    for r, (trees, losses, pred) in enumerate(res):
        assert trees == want_trees, f"rank {r}: trees differ from the single-rank run"
        r0, r1 = (n * r) // world, (n * (r + 1)) // world
        np.testing.assert_array_equal(pred, want_pred[r0:r1])
        for (l, s), (wl, ws) in zip(losses, want_loss):
            assert abs(l - wl) <= 1e-6 * abs(wl) and abs(s - ws) <= 1e-6


def test_feature_shard_matches_single_rank():
    n, f, iters, world = 40000, 10, 5, 2
    bins, nb, na, y = synth(n, f, seed=6, bins=64)
    want_trees, want_loss, want_pred, _ = _single(bins, nb, na, y, iters, max_depth=6)
    comm = FakeComm(world)

    def rank_main(r):
        ds = ydf_b200.Dataset(bins, nb, na)
        gbt = ydf_b200.Gbt(ds, ydf_b200.default_config(num_trees=iters, max_depth=6))
        gbt.set_labels(y)
        b, e = ydf_b200.feature_shard(f, r, world)
        gbt.set_feature_shard(b, e, r, world, comm.allgather(r))
        gbt.train(iters)
        return [gbt.get_tree(i).tobytes() for i in range(iters)], gbt.get_predictions()

    for trees, pred in _run_ranks(world, rank_main, comm):
        assert trees == want_trees
        np.testing.assert_array_equal(pred, want_pred)


@pytest.mark.parametrize("mode", ["rows", "features"])
def test_shards_with_categorical_features_match_single_rank(mode):
    """The positive-category masks travel in the shard-best record (feature shards) / are derived from
    the all-reduced histograms (row shards)."""
    n, iters, world = 40000, 5, 2
    bins, nb, na, ft, y = synth_mixed(n, 4, [6, 50, 256], seed=9)
    f = len(ft)
    want_trees, want_loss, want_pred, init = _single(bins, nb, na, y, iters, ft=ft, max_depth=6)
    assert any((np.frombuffer(t, dtype=ydf_b200.NODE_DTYPE)["condition_type"] == 1).any() for t in want_trees)
    comm = FakeComm(world)

    def rank_main(r):
        r0, r1 = ((n * r) // world, (n * (r + 1)) // world) if mode == "rows" else (0, n)
        ds = ydf_b200.Dataset(bins[:, r0:r1], nb, na, feature_types=ft)
        gbt = ydf_b200.Gbt(ds, ydf_b200.default_config(num_trees=iters, max_depth=6))
        gbt.set_labels(y[r0:r1])
        if mode == "rows":
            gbt.set_row_shard(r, world, n, init, comm.allreduce(r))
        else:
            b, e = ydf_b200.feature_shard(f, r, world)
            gbt.set_feature_shard(b, e, r, world, comm.allgather(r))
        gbt.train(iters)
        return [gbt.get_tree(i).tobytes() for i in range(iters)], gbt.get_predictions(), (r0, r1)

    for trees, pred, (r0, r1) in _run_ranks(world, rank_main, comm):
        assert trees == want_trees
        np.testing.assert_array_equal(pred, want_pred[r0:r1])


@pytest.mark.parametrize("loss,hess,world,cat", [(0, 0, 2, False), (0, 1, 3, True), (1, 0, 4, True)])
def test_row_shard_scatter_matches_single_rank(loss, hess, world, cat):
    """Reduce-scatter by feature chunk + sharded scan + all-gather of the bests (ygg_gbt_set_row_shard_scatter);
    the fake reduce-scatter poisons every chunk a rank does not own."""
    n, iters = 48000, 5
    task = "binary" if loss == 0 else "regression"
    if cat:
        bins, nb, na, ft, y = synth_mixed(n, 5, [7, 60], seed=12, task=task)
    else:
        bins, nb, na, y = synth(n, 9, seed=5, task=task, bins=64)
        ft = None
    kw = dict(loss=loss, use_hessian_gain=hess, max_depth=6)
    want_trees, want_loss, want_pred, init = _single(bins, nb, na, y, iters, ft=ft, **kw)
    comm = FakeComm(world)

    def rank_main(r):
        r0, r1 = (n * r) // world, (n * (r + 1)) // world
        ds = ydf_b200.Dataset(bins[:, r0:r1], nb, na, feature_types=ft)
        gbt = ydf_b200.Gbt(ds, ydf_b200.default_config(num_trees=iters, **kw))
        gbt.set_labels(y[r0:r1])
        gbt.set_row_shard_scatter(r, world, n, init, allreduce=comm.allreduce(r), reducescatter=comm.reducescatter(r),
                                  allgather=comm.allgather(r))
        gbt.train(iters)
        return ([gbt.get_tree(i).tobytes() for i in range(iters)], [gbt.train_loss(i) for i in range(iters)],
                gbt.get_predictions())

    res = _run_ranks(world, rank_main, comm)
    for r, (trees, losses, pred) in enumerate(res):
        assert trees == want_trees, f"rank {r}: trees differ from the single-rank run"
        r0, r1 = (n * r) // world, (n * (r + 1)) // world
        np.testing.assert_array_equal(pred, want_pred[r0:r1])
        for (l, s), (wl, ws) in zip(losses, want_loss):
            assert abs(l - wl) <= 1e-6 * abs(wl) and abs(s - ws) <= 1e-6


@pytest.mark.parametrize("loss,world,scatter", [(0, 2, False), (1, 3, True), (0, 4, True)])
def test_weighted_row_shards_match_single_rank(loss, world, scatter):
    """Example weights with row shards: the weight-sum plane is reduced with the histograms, the fixed-point scales (largest
    weight, max |w*g|) and the per-node weight sums are the job's — bit-identical trees, losses to the last digits."""
    n, iters = 48000, 5
    bins, nb, na, ft, y = synth_mixed(n, 5, [7, 60], seed=14, task="binary" if loss == 0 else "regression")
    w = np.random.default_rng(3).uniform(0.1, 5.0, n).astype(np.float32)
    w[n // 2:] *= 0.25      # the largest weight lives on the first ranks only
    kw = dict(loss=loss, max_depth=6)
    ds = ydf_b200.Dataset(bins, nb, na, feature_types=ft)
    one = ydf_b200.Gbt(ds, ydf_b200.default_config(num_trees=iters, **kw))
    one.set_weights(w)
    one.set_labels(y)
    one.train(iters)
    want_trees = [one.get_tree(i).tobytes() for i in range(iters)]
    want_loss = [one.train_loss(i) for i in range(iters)]
    want_pred, init = one.get_predictions(), one.initial_prediction()
    comm = FakeComm(world)

    def rank_main(r):
        r0, r1 = (n * r) // world, (n * (r + 1)) // world
        rds = ydf_b200.Dataset(bins[:, r0:r1], nb, na, feature_types=ft)
        gbt = ydf_b200.Gbt(rds, ydf_b200.default_config(num_trees=iters, **kw))
        gbt.set_weights(w[r0:r1])
        gbt.set_labels(y[r0:r1])
        if scatter:
            gbt.set_row_shard_scatter(r, world, n, init, allreduce=comm.allreduce(r), reducescatter=comm.reducescatter(r),
                                      allgather=comm.allgather(r))
        else:
            gbt.set_row_shard(r, world, n, init, comm.allreduce(r))
        gbt.train(iters)
        return ([gbt.get_tree(i).tobytes() for i in range(iters)], [gbt.train_loss(i) for i in range(iters)],
                gbt.get_predictions())

    res = _run_ranks(world, rank_main, comm)
    for r, (trees, losses, pred) in enumerate(res):
        assert trees == want_trees, f"rank {r}: trees differ from the single-rank run"
        r0, r1 = (n * r) // world, (n * (r + 1)) // world
        np.testing.assert_array_equal(pred, want_pred[r0:r1])
        for (l, s), (wl, ws) in zip(losses, want_loss):
            assert abs(l - wl) <= 1e-6 * abs(wl) and abs(s - ws) <= 1e-6
