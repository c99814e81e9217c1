"""Time the mask-level logits (cal_pred_logits) at the two BASELINE shapes."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from xmask3d_b200 import ops
dev = torch.device("cuda", 0)
g = torch.Generator(device=dev).manual_seed(1)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
for name, rows, t in (("B15N4 160x50 x 20", 8000, 20), ("B170N30 160x100 x 201", 16000, 201)):
    me = torch.randn(rows, 768, device=dev, generator=g); te = torch.randn(t - 1, 768, device=dev, generator=g)
    ne = torch.randn(1, 768, device=dev, generator=g)
    ws = ops._ws(ops.L.lib().xm3d_logits_ws_bytes(rows, t - 1, 768, t - 1), dev)
    for _ in range(5):
        ops.logits(me, te, ne, [1] * (t - 1), 1 / 0.07, ws=ws)
    torch.cuda.synchronize()
    e0.record()
    for _ in range(50):
        ops.logits(me, te, ne, [1] * (t - 1), 1 / 0.07, ws=ws)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 50
    print(f"{name}: {ms * 1000:.1f} us per call ({2.0 * rows * 768 * t / ms / 1e9:.1f} useful TFLOP/s)")
