"""Experiment: nearest-seen-neighbour fill against sklearn's KDTree (the reference's dependency) on one
scene of 150 k / 1 M points, 40 % of the points unseen at random plus a 1 m slab nobody has seen."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from sklearn.neighbors import KDTree
from xmask3d_b200 import ops, synthetic as syn
dev = torch.device("cuda", 0)
for n in (150_000, 1_000_000):
    sc = syn.make_scene(3, n, room=(12.0, 10.0, 3.0)) if n > 500_000 else syn.make_scene(3, n)
    xyz = torch.from_numpy(sc.xyz).to(dev)
    rng = np.random.default_rng(0)
    cnt = torch.from_numpy((rng.uniform(size=n) < 0.6).astype(np.int32)).to(dev)
    cnt[xyz[:, 0] < xyz[:, 0].min() + 1.0] = 0
    for _ in range(2):
        m = ops.nn_fill_match(xyz, cnt)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); m = ops.nn_fill_match(xyz, cnt); e1.record(); torch.cuda.synchronize()
    c = cnt.cpu().numpy()
    t = time.time(); tree = KDTree(sc.xyz[c != 0]); d, i = tree.query(sc.xyz[c == 0], k=1); t = time.time() - t
    ref = np.nonzero(c != 0)[0][i.flatten()]
    mine = m.cpu().numpy()[c == 0]
    print(f"{n} points, {int((c == 0).sum())} unseen: {e0.elapsed_time(e1):.2f} ms on the GPU, sklearn KDTree {t * 1e3:.0f} ms, "
          f"identical indices {float((mine == ref).mean()):.6f}")
