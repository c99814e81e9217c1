"""Out-of-bounds WRITE check of the round-2 kernels through the raw C ABI (compute-sanitizer is not available on the
pool): every output buffer is carved out of a larger arena filled with a canary pattern, with guard bands on both
sides; after the call the guards must be untouched and the payload must equal what the torch front end returns."""
import ctypes as C

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
GUARD = 4096


@pytest.fixture(scope="module")
def dev():
    assert torch.cuda.is_available(), "GPU tests need a CUDA device"
    return torch.device("cuda", 0)


class Arena:
    def __init__(self, dev, nbytes=1 << 28):
        self.buf = torch.full((nbytes,), 0xA5, dtype=torch.uint8, device=dev)
        self.off = GUARD
        self.parts = []

    def take(self, shape, dtype):
        n = int(np.prod(shape)) * torch.empty(0, dtype=dtype).element_size()
        a = (self.off + 255) // 256 * 256
        t = self.buf[a:a + n].view(dtype).view(shape)
        self.parts.append((a, n))
        self.off = a + n + GUARD
        return t

    def check(self):
        mask = torch.ones(self.off, dtype=torch.bool, device=self.buf.device)
        for a, n in self.parts:
            mask[a:a + n] = False
        assert bool((self.buf[:self.off][mask] == 0xA5).all()), "a kernel wrote outside its output buffers"


def _p(t):
    return None if t is None else C.c_void_p(t.data_ptr())


def test_pool_mma_contra_pack_logits_stay_inside_their_buffers(dev):
    from xmask3d_b200 import _lib as L, ops
    lib = L.lib()
    st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
    g = torch.Generator(device=dev).manual_seed(1)
    ns = [1000, 0, 130, 64, 2999]
    off = torch.tensor(np.concatenate([[0], np.cumsum(ns)]), dtype=torch.int64, device=dev)
    total, n_seg = int(off[-1]), len(ns)
    for k, c in ((50, 256), (100, 128), (64, 768)):
        words = (k + 31) // 32
        feat = torch.randn(total, c, device=dev, generator=g)
        member = torch.randint(-2 ** 31, 2 ** 31 - 1, (total, words), device=dev, generator=g, dtype=torch.int64).to(torch.int32)
        ar = Arena(dev)
        s, cnt, mean = ar.take((n_seg, k, c), torch.float32), ar.take((n_seg, k), torch.int32), ar.take((n_seg, k, c), torch.float32)
        ws = ops._ws(lib.xm3d_pool_ws_bytes(n_seg, k, c, total, total * 40), dev)
        L.check(lib.xm3d_pool_batch(_p(feat), c, None, _p(member), None, n_seg, k, _p(off), total, total * 40, L.POOL_MMA,
                                    _p(s), _p(cnt), _p(mean), _p(ws), ws.numel(), None, st))
        torch.cuda.synchronize()
        ar.check()
        rs, rc, rm = ops.pool(feat, off, k, member=member, cap=total, cap_pairs=total * 40, path="mma")
        assert torch.equal(rs, s) and torch.equal(rc, cnt) and torch.equal(rm, mean)
        # loss_contra selection on the same membership words
        h, w = 24, 32
        lg = torch.randn(n_seg, k, h, w, device=dev, generator=g)
        gt = torch.randint(0, 3, (total,), device=dev, generator=g).float()
        ar = Arena(dev)
        counts, kind, score = ar.take((n_seg, k, 3), torch.int32), ar.take((n_seg, k), torch.int8), ar.take((n_seg, k), torch.float32)
        sel, n_sel, selm = ar.take((n_seg, 5), torch.int32), ar.take((n_seg,), torch.int32), ar.take((total,), torch.int32)
        ws = ops._ws(lib.xm3d_contra_ws_bytes(n_seg, k), dev)
        L.check(lib.xm3d_contra_select_batch(_p(member), k, _p(gt), _p(off), n_seg, total, _p(lg), h, w, _p(counts), _p(kind),
                                             _p(score), _p(sel), _p(n_sel), _p(selm), _p(ws), ws.numel(), st))
        torch.cuda.synchronize()
        ar.check()
        ref = ops.contra_select(member, k, gt, off, lg, cap=total)
        assert torch.equal(ref.sel, sel) and torch.equal(ref.counts, counts) and torch.equal(ref.sel_member.view(-1), selm)
    # int16 packing with a device-side row count smaller than the buffer
    src = torch.randint(0, 30000, (5001, 3), device=dev, generator=g, dtype=torch.int32)
    ar = Arena(dev)
    dst, flag = ar.take((5001, 3), torch.int16), ar.take((1,), torch.int32)
    flag.zero_()
    rows = torch.tensor([4000], dtype=torch.int64, device=dev)
    L.check(lib.xm3d_pack_i16(_p(src), _p(rows), 5001, 3, _p(dst), _p(flag), st))
    torch.cuda.synchronize()
    ar.check()
    assert torch.equal(dst[:4000].to(torch.int32), src[:4000]) and bool((dst[4000:].view(torch.uint8) == 0xA5).all())
    # per-point logits with the fused-stream epilogue, rows not a multiple of the 128-row tile
    n, t, c = 1000 + 77, 19, 256
    feat = torch.randn(n, c, device=dev, generator=g)
    te = torch.randn(t, c, device=dev, generator=g)
    binary = (torch.rand(n, device=dev, generator=g) > 0.5).float()
    is_base = (torch.arange(t, device=dev) < 14).to(torch.uint8)
    label = torch.randint(-1, 7, (n,), device=dev, generator=g, dtype=torch.int32)
    logq = torch.rand(7, t, device=dev, generator=g).softmax(-1).log()
    ar = Arena(dev)
    out, amax = ar.take((n, t), torch.float32), ar.take((n,), torch.int32)
    ws = ops._ws(lib.xm3d_point_logits_ws_bytes(t, c), dev)
    L.check(lib.xm3d_point_logits(_p(feat), n, c, _p(te), t, 1 / 0.07, _p(binary), _p(is_base), _p(label), _p(logq), 7, 0.65, 0.35,
                                  _p(out), _p(amax), _p(ws), ws.numel(), st))
    torch.cuda.synchronize()
    ar.check()
    assert bool(torch.isfinite(out[out > -1e9]).all()) and int(amax.min()) >= 0 and int(amax.max()) < t
