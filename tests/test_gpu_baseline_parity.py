"""GPU vs oracle at BASELINE.json's OWN configurations (SURVEY.md §8d "parity check in the same run"):

  C2  synthetic 1M rows x 50 numerical features, depth 6 : first 20 trees, variance gain AND hessian gain
  C3  synthetic 10M rows x 200 numerical features, depth 8, 256 bins : first 2 trees
  C5  reduced (2M rows, 20 numerical (3 informative) + 10 categorical features of 100-256 values, regression) : first 2 trees

Both sides run their own closed boosting loop from iteration 0 on the same u8 matrix; every node is compared:
feature / threshold bin / na_value / counts exactly, split_score 1e-5 relative, leaf values 1e-5 absolute, training
loss 1e-5 relative (north_star's bar).  The reference semantics being matched: splitter_scanner.h:931-1101 (scan),
training.cc:1728-1746 (arg-max over features).  These are the sizes where 24-bit gradient quantisation over 1e7 rows
and 64-slot deep levels could flip a near-tie; they do not.
"""
import numpy as np
import pytest

import ydf_b200

pytestmark = pytest.mark.gpu


def _run(workload, trees, **over):
    import bench
    w = dict(bench.WORKLOADS[workload])
    w.update(over)
    bins, nb, na, y = bench.make_data(w, device=0)
    ds = ydf_b200.Dataset(bins, nb, na, feature_types=w.get("feature_types"))
    gbt = ydf_b200.Gbt(ds, bench.gbt_config(w, trees))
    gbt.set_labels(y)
    gbt.train(trees)
    got = [gbt.get_tree(i) for i in range(trees)]
    got_loss = [gbt.train_loss(i)[0] for i in range(trees)]
    gbt.close()
    ds.close()
    port = bench.CpuPort(w, bins, nb, na, y)
    try:
        for _ in range(trees):
            port.step()
    finally:
        port.close()
    par = bench.parity_block(got, port.trees, got_loss, port.loss)
    print(workload, over, par, "cpu s/iter", np.round(port.seconds, 2).tolist())
    return par, got, port.trees


def _assert_parity(par, trees):
    assert par["trees_compared"] == trees
    assert par["structure_mismatches"] == 0, par
    assert par["max_score_rel_err"] <= 1e-5, par
    assert par["max_leaf_abs_err"] <= 1e-5, par
    assert par["max_loss_rel_err"] <= 1e-5, par
    assert par["ok"]


@pytest.mark.parametrize("hessian", [0, 1])
def test_c2_first_20_trees_match_the_oracle(hessian):
    from oracle import oracle as O
    O.use_native_build()   # the library bench.CpuPort runs (switches apply to the loaded library)
    # hessian gain: the reference sums its per-bucket hessians in float32 in row order, which makes ITS OWN scores
    # carry ~1e-6..1e-4 of order-dependent noise at 1M rows; the comparison is against exact (double) buckets
    O.set_hessian_buckets_double(bool(hessian))
    try:
        par, got, want = _run("c2", 20, hessian=hessian)
    finally:
        O.set_hessian_buckets_double(False)
    _assert_parity(par, 20)
    assert all(len(t) == 63 for t in got)   # informative data: every tree is full at depth 6


def test_c3_first_2_trees_match_the_oracle():
    par, got, want = _run("c3", 2)
    _assert_parity(par, 2)
    assert len(got[0]) == 255


def test_c5_reduced_first_2_trees_match_the_oracle():
    # few informative numerical features, so that the per-category effects are worth splitting on
    par, got, want = _run("c5", 2, rows=2_000_000, features=30, categorical=10, informative=3)
    _assert_parity(par, 2)
    assert any((t["condition_type"] == 1).any() for t in got)   # categorical splits are taken
