import sys
import numpy as np
import tests.test_gpu_fuzz as F
from tests.util import compare_trees

seed = int(sys.argv[1])
c = F.draw_variant(seed)
print(c)
# re-run with verbose diff: monkeypatch compare to print both nodes
orig = F.compare
def verbose(a, b, kw, **more):
    e = orig(a, b, kw, **more)
    if e:
        if len(a) != len(b):
            for i in range(min(len(a), len(b))):
                if a[i]["feature"] != b[i]["feature"] or a[i]["num_examples"] != b[i]["num_examples"]:
                    print("first structural diff at", i, "\n  engine", a[i], "\n  oracle", b[i]); break
        else:
            i = int(e[0].split(":")[0].split()[1])
            print("node", i, "\n  engine", a[i], "\n  oracle", b[i], "\n  parent of it:")
            for j in range(len(a)):
                if a[j]["pos_child"] == i or a[j]["neg_child"] == i:
                    print("  engine", a[j], "\n  oracle", b[j])
    return e
F.compare = verbose
print(F.run_variant(c))
