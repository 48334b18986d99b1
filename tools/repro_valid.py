import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, ydf_b200
rng = np.random.default_rng(5)
n = 8000
x = rng.normal(size=(n, 3)).astype(np.float32)
lab = np.where((x[:, 0] > 0) ^ (rng.random(n) < 0.3), "a", "b")
data = {"x0": x[:, 0], "x1": x[:, 1], "x2": x[:, 2], "y": lab}
learner = ydf_b200.GradientBoostedTreesLearner(label="y", discretize_numerical_columns=True, num_trees=int(sys.argv[1]) if len(sys.argv) > 1 else 200,
                                               shrinkage=0.3, max_depth=6)
model = learner.train(data)
print(model.num_trees(), len(model.training_logs), model.validation_loss)
