import sys, time
sys.path.insert(0, ".")
import torch
from iris_b200 import core, scenes
sc = scenes.room(1_000_000, 16, seed=0)
torch.zeros(1, device="cuda")
for k in range(4):
    t0 = time.time(); s = core.Scene(sc.vertices, sc.faces, 0); torch.cuda.synchronize(); print("build", k, "wall %.1f ms" % ((time.time() - t0) * 1e3), "library %.1f ms" % s.stats()["build_ms"]); del s
sc5 = scenes.room(5_000_000, 16, seed=0)
for k in range(2):
    t0 = time.time(); s = core.Scene(sc5.vertices, sc5.faces, 0); torch.cuda.synchronize(); print("build 5M", k, "wall %.1f ms" % ((time.time() - t0) * 1e3), "library %.1f ms" % s.stats()["build_ms"]); del s
