import sys
sys.path.insert(0, '.')
from tests import reference_replay as R
from tests.test_reference_replay_gpu import engine_trainer
for run, iters in (("adult", 30), ("abalone", None)):
    ref, data = R.load_run(run)
    for name, tr in (("engine", engine_trainer), ("oracle", R.oracle_trainer_shuffled)):
        seen = R.replay_trees(ref, data, tr, num_iterations=iters, score_rtol=1e-5, leaf_atol=1e-5)
        print(run, name, {k: (round(v, 9) if isinstance(v, float) else v) for k, v in seen.items()})
