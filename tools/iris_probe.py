import numpy as np, sys
sys.path.insert(0, '.')
import ydf_b200
from tests import reference_replay as R
ref, data = R.load_run("iris")
for tb in ("FEATURE_ORDER", "LIBCXX_SHUFFLE", "LIBSTDCXX_SHUFFLE"):
    model = ydf_b200.GradientBoostedTreesLearner(label="class", tie_break=tb).train({k: np.asarray(v) for k, v in data.items()})
    logs = model.training_logs
    n = min(len(logs), 28)
    out = [tb, len(logs), model.num_trees(), round(model.validation_loss, 6), float(ref["validation_loss"])]
    for key, mine in (("log_training_loss", "loss"), ("log_validation_loss", "validation_loss")):
        got = np.array([e[mine] for e in logs[:n]], np.float64); want = ref[key][:n].astype(np.float64)
        out.append((mine, float(np.abs(got - want).max()), int(np.argmax(np.abs(got-want)>1e-5)) if (np.abs(got-want)>1e-5).any() else -1))
    print(out)
