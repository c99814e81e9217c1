import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, torch.nn.functional as F
from xmask3d_b200 import ops
dev = torch.device("cuda", 0)
V, K = 160, 50
lg = torch.randn(V, K, 128, 128, device=dev)
sc = torch.rand(V, K, device=dev)
def t(fn, name, n=10):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    print(f"{name:60s} {e0.elapsed_time(e1)/n*1000:9.1f} us", flush=True)
t(lambda: ops.mask_prep(lg, (240, 320), mode="sigmoid_gt0.5", want_bits=True), "fused: bits only")
t(lambda: ops.mask_prep(lg, (240, 320), scores=sc, want_bits=False, want_partition=True), "fused: partition labels + areas")
t(lambda: ops.mask_prep(lg, (240, 320), scores=sc, want_bits=True, want_partition=True), "fused: bits + partition")
def torch_path():
    up = F.interpolate(lg, size=(240, 320), mode="bilinear", align_corners=False)
    sg = up.sigmoid()
    ids = (sc.view(V, K, 1, 1) * sg).argmax(1)
    return sg > 0.5, ids
t(torch_path, "torch CUDA ops: interpolate + sigmoid + mul + argmax + >0.5", n=3)
