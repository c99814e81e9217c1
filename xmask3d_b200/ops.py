"""Torch-tensor front end of the C ABI: allocates outputs / workspaces with torch, passes raw
device pointers and the current CUDA stream to libxm3d.  Everything here is batched over
segments (one segment = one (scene, view)); the reference-shaped single-call shims live in
voxelizer.py / voxelization_utils.py / fusion_util.py / fuser.py / logits.py.

No CPU fallback: every function requires CUDA tensors (or moves numpy input to `device`).
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass
from typing import List, Optional, Sequence

import numpy as np
import torch

from . import _lib as L


def _require_cuda():
    if not torch.cuda.is_available():
        raise L.Xm3dError("xmask3d_b200 needs a CUDA device (sm_100a); there is no CPU fallback")


def _ptr(t: Optional[torch.Tensor]):
    return None if t is None else C.c_void_p(t.data_ptr())


def _stream():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def _ws(nbytes: int, device) -> torch.Tensor:
    return torch.empty(max(int(nbytes), 256), dtype=torch.uint8, device=device)


def _dev_contig(t: torch.Tensor, dtype) -> torch.Tensor:
    assert t.is_cuda, "expected a CUDA tensor"
    if t.dtype != dtype:
        t = t.to(dtype)
    return t.contiguous()


# ----------------------------------------------------------------------------- stage 2
@dataclass
class Projection:
    vis: torch.Tensor                 # uint8 [total_pts]   (views concatenated, see out_off)
    n_vis: torch.Tensor               # int32 [V]
    vis_off: torch.Tensor             # int64 [V+1]
    vis_idx: torch.Tensor             # int32 [cap_vis]
    rowcol: torch.Tensor              # int32 [cap_vis,2]   (x_label, y_label)
    xyz_vis: torch.Tensor             # float32 [cap_vis,3]
    status: torch.Tensor              # int32 [1]
    out_off: np.ndarray               # int64 [V+1] host: start of every view in `vis` / `mapping`
    mapping: Optional[torch.Tensor] = None   # int64 [total_pts,3]


def make_views(w2c: np.ndarray, intr: Sequence[float], pt_off: Sequence[int], n_pts: Sequence[int],
               depth_shape=None, depth_off: Optional[Sequence[int]] = None):
    """Build the host array of xm3d_view_t records.  w2c: [V,4,4] float64 world->camera
    (np.linalg.inv(pose), reference models/utils/fusion_util.py:70); intr = (fx,fy,cx,cy)."""
    v = len(n_pts)
    arr = (L.View * v)()
    out_off = np.zeros(v + 1, np.int64)
    for i in range(v):
        r = arr[i]
        m = np.ascontiguousarray(w2c[i], np.float64)
        for j in range(12):
            r.w2c[j] = m[j // 4, j % 4]
        r.fx, r.fy, r.cx, r.cy = (float(x) for x in intr)
        r.pt_off, r.n_pts, r.out_off = int(pt_off[i]), int(n_pts[i]), int(out_off[i])
        out_off[i + 1] = out_off[i] + int(n_pts[i])
        if depth_shape is not None:
            r.depth_h, r.depth_w = int(depth_shape[0]), int(depth_shape[1])
            r.depth_off = int(depth_off[i]) if depth_off is not None else i * int(depth_shape[0]) * int(depth_shape[1])
        else:
            r.depth_off = -1
    return arr, out_off


def views_to_device(views, device) -> torch.Tensor:
    """Device copy of the host view records (uint8 [V*192]) for graph-capturable projection calls."""
    raw = np.frombuffer(bytes(views), dtype=np.uint8).copy()
    return torch.from_numpy(raw).to(device)


def project_batch(xyz: torch.Tensor, views, out_off: np.ndarray, depth: Optional[torch.Tensor],
                  depth_scale: float = 1000.0, image_dim=(320, 240), cut_bound: int = 10,
                  vis_thres: float = 0.25, cap_vis: Optional[int] = None, want_mapping: bool = False,
                  want_compact: bool = True, ws: Optional[torch.Tensor] = None,
                  views_dev: Optional[torch.Tensor] = None) -> Projection:
    """xyz: float32 [sum N_scene,3] (all scenes concatenated); views: (View * V) host records;
    depth: uint16 / int16-viewed / float64 CUDA tensor holding every view's image, or None."""
    _require_cuda()
    dev = xyz.device
    xyz = _dev_contig(xyz, torch.float32)
    n_views = len(views)
    total = int(out_off[-1])
    cap = total if cap_vis is None else int(cap_vis)
    if depth is None:
        kind = L.DEPTH_NONE
    elif depth.dtype == torch.float64:
        kind = L.DEPTH_F64
    elif depth.dtype in (torch.uint16, torch.int16):
        kind = L.DEPTH_U16
    else:
        raise TypeError(f"depth dtype {depth.dtype}: expected uint16 (raw) or float64 (metres)")
    if depth is not None:
        assert depth.is_cuda and depth.is_contiguous()
    vis = torch.empty(max(total, 1), dtype=torch.uint8, device=dev)
    mapping = torch.empty((max(total, 1), 3), dtype=torch.int64, device=dev) if want_mapping else None
    n_vis = torch.zeros(max(n_views, 1), dtype=torch.int32, device=dev)
    vis_off = torch.zeros(n_views + 1, dtype=torch.int64, device=dev)
    status = torch.zeros(1, dtype=torch.int32, device=dev)
    vis_idx = rowcol = xyz_vis = None
    if want_compact:
        vis_idx = torch.empty(max(cap, 1), dtype=torch.int32, device=dev)
        rowcol = torch.empty((max(cap, 1), 2), dtype=torch.int32, device=dev)
        xyz_vis = torch.empty((max(cap, 1), 3), dtype=torch.float32, device=dev)
    max_pts = max((int(v.n_pts) for v in views), default=0)
    need = L.lib().xm3d_project_ws_bytes(n_views, total, max_pts)
    if ws is None or ws.numel() < need:
        ws = _ws(need, dev)
    L.check(L.lib().xm3d_project_batch(
        _ptr(xyz), C.cast(views, C.c_void_p), _ptr(views_dev), n_views, total, _ptr(depth), kind, float(depth_scale),
        int(image_dim[0]), int(image_dim[1]), int(cut_bound), float(vis_thres),
        _ptr(vis), _ptr(mapping), _ptr(n_vis), _ptr(vis_off), cap, _ptr(vis_idx), _ptr(rowcol), _ptr(xyz_vis),
        _ptr(ws), ws.numel(), _ptr(status), _stream()))
    return Projection(vis, n_vis, vis_off, vis_idx, rowcol, xyz_vis, status, out_off, mapping)


# ----------------------------------------------------------------------------- stage 1
@dataclass
class Unique:
    m: torch.Tensor          # int32 [n_seg] unique count
    uniq_off: torch.Tensor   # int64 [n_seg+1]
    first: torch.Tensor      # int32 [cap] first-occurrence index (within segment), unique order
    inverse: torch.Tensor    # int32 [cap]
    counts: Optional[torch.Tensor]
    status: torch.Tensor
    voxel_xyz: Optional[torch.Tensor] = None   # int32 [cap,3]
    grid_min: Optional[torch.Tensor] = None    # int32 [n_seg,3]
    ws: Optional[torch.Tensor] = None          # workspace of the call (voxel_path_info reads it)
    cap: Optional[int] = None


def _vox_path(mode: int = 0, unit_pts: int = 0) -> int:
    """Per-call path word of xm3d_unique_batch / xm3d_voxelize_batch: mode 0 = shared-memory units when every
    segment fits (default), 1 = multi-kernel path only, 2 = shared-memory units only (the fallback kernels are not
    launched; status flag FLAG_VOX_FALLBACK if the batch needed them); unit_pts < 7000 forces several key-range
    units per segment (tests)."""
    return (int(mode) & 0xff) | (max(int(unit_pts), 0) << 8)


def voxel_path_info(u: "Unique"):
    """(not_eligible, overflowed) of the call that produced `u` (synchronises)."""
    out = (C.c_int32 * 2)()
    n_seg = u.m.numel()
    L.check(L.lib().xm3d_voxel_path_info(_ptr(u.ws), n_seg, int(u.first.numel() if u.cap is None else u.cap),
                                         C.cast(out, C.c_void_p), _stream()))
    return int(out[0]), int(out[1])


def unique_batch(keys: torch.Tensor, seg_off: torch.Tensor, cap: Optional[int] = None, collate: bool = False,
                 want_counts: bool = False, mode: int = 0, unit_pts: int = 0) -> Unique:
    """np.unique(keys, return_index, return_inverse[, return_counts]) per segment.
    keys: int64/uint64-viewed CUDA tensor; seg_off int64 [n_seg+1] CUDA."""
    _require_cuda()
    dev = keys.device
    keys = keys.contiguous()
    assert keys.element_size() == 8
    seg_off = _dev_contig(seg_off, torch.int64)
    n_seg = seg_off.numel() - 1
    cap = int(keys.numel()) if cap is None else int(cap)
    m = torch.zeros(n_seg, dtype=torch.int32, device=dev)
    uniq_off = torch.zeros(n_seg + 1, dtype=torch.int64, device=dev)
    first = torch.empty(max(cap, 1), dtype=torch.int32, device=dev)
    inverse = torch.empty(max(cap, 1), dtype=torch.int32, device=dev)
    counts = torch.zeros(max(cap, 1), dtype=torch.int32, device=dev) if want_counts else None
    status = torch.zeros(1, dtype=torch.int32, device=dev)
    ws = _ws(L.lib().xm3d_unique_ws_bytes(n_seg, cap), dev)
    L.check(L.lib().xm3d_unique_batch(_ptr(keys), _ptr(seg_off), n_seg, cap, _ptr(m), _ptr(uniq_off), _ptr(first),
                                      _ptr(counts), _ptr(inverse), int(collate), _vox_path(mode, unit_pts), _ptr(ws),
                                      ws.numel(), _ptr(status), _stream()))
    return Unique(m, uniq_off, first, inverse, counts, status, ws=ws, cap=cap)


def voxelize_batch(xyz: torch.Tensor, seg_off: torch.Tensor, rt: torch.Tensor, cap: Optional[int] = None,
                   collate: bool = False, ws: Optional[torch.Tensor] = None, mode: int = 0, unit_pts: int = 0) -> Unique:
    """xyz float32 or float64 [cap,3] CUDA (segments concatenated; float64 = the augmented training path, where
    ElasticDistortion hands the voxelizer float64 coordinates); rt float64 [n_seg,3,4] (rows 0..2 of the
    rigid transformation, reference dataset/voxelizer.py:104-108)."""
    _require_cuda()
    dev = xyz.device
    xyz = _dev_contig(xyz, torch.float64 if xyz.dtype == torch.float64 else torch.float32)
    seg_off = _dev_contig(seg_off, torch.int64)
    rt = _dev_contig(rt, torch.float64)
    n_seg = seg_off.numel() - 1
    assert rt.numel() == n_seg * 12
    cap = int(xyz.shape[0]) if cap is None else int(cap)
    m = torch.zeros(n_seg, dtype=torch.int32, device=dev)
    uniq_off = torch.zeros(n_seg + 1, dtype=torch.int64, device=dev)
    first = torch.empty(max(cap, 1), dtype=torch.int32, device=dev)
    inverse = torch.empty(max(cap, 1), dtype=torch.int32, device=dev)
    voxel = torch.empty((max(cap, 1), 3), dtype=torch.int32, device=dev)
    gmin = torch.empty((n_seg, 3), dtype=torch.int32, device=dev)
    status = torch.zeros(1, dtype=torch.int32, device=dev)
    need = L.lib().xm3d_voxelize_ws_bytes(n_seg, cap)
    if ws is None or ws.numel() < need:
        ws = _ws(need, dev)
    L.check(L.lib().xm3d_voxelize_batch(_ptr(xyz), int(xyz.dtype == torch.float64), _ptr(seg_off), n_seg, cap, _ptr(rt),
                                        _ptr(m), _ptr(uniq_off), _ptr(first), _ptr(inverse), int(collate), _ptr(voxel),
                                        _ptr(gmin), _vox_path(mode, unit_pts), _ptr(ws), ws.numel(), _ptr(status),
                                        _stream()))
    return Unique(m, uniq_off, first, inverse, None, status, voxel, gmin, ws=ws, cap=cap)


def fnv_hash(coords: torch.Tensor) -> torch.Tensor:
    """fnv_hash_vec on a float64 [n,dim] CUDA tensor -> int64 tensor holding the uint64 bits."""
    _require_cuda()
    coords = _dev_contig(coords, torch.float64)
    n, dim = coords.shape
    keys = torch.empty(n, dtype=torch.int64, device=coords.device)
    L.check(L.lib().xm3d_fnv_hash_f64(_ptr(coords), n, dim, _ptr(keys), _stream()))
    return keys


def ravel_hash(coords: torch.Tensor) -> torch.Tensor:
    _require_cuda()
    coords = _dev_contig(coords, torch.float64)
    n, dim = coords.shape
    keys = torch.empty(n, dtype=torch.int64, device=coords.device)
    ws = _ws(L.lib().xm3d_ravel_ws_bytes(dim), coords.device)
    L.check(L.lib().xm3d_ravel_hash_f64(_ptr(coords), n, dim, _ptr(keys), _ptr(ws), ws.numel(), _stream()))
    return keys


# ----------------------------------------------------------------------------- stage 3
THR = {"ge0.5": L.THR_GE_HALF, "sigmoid_ge0.5": L.THR_SIGMOID_GE_HALF, "sigmoid_gt0.5": L.THR_SIGMOID_GT_HALF}


POOL_PATH = {"auto": L.POOL_AUTO, "pair_lists": L.POOL_PAIR_LISTS, "rows": L.POOL_ROWS, "mma": L.POOL_MMA}


def mask_words(k: int) -> int:
    return (int(k) + 31) // 32


def gather_masks(masks: torch.Tensor, rowcol: torch.Tensor, seg_off: torch.Tensor, mode: str = "ge0.5",
                 cap: Optional[int] = None, want_counts: bool = False, ws: Optional[torch.Tensor] = None):
    """masks [n_seg,k,h,w] bool/uint8/float32 CUDA; rowcol int32 [cap,2].
    Returns (member uint32-as-int32 [cap,words], counts int32 [n_seg,k] or None)."""
    _require_cuda()
    dev = masks.device
    assert masks.dim() == 4
    n_seg, k, h, w = masks.shape
    if masks.dtype == torch.bool:
        masks = masks.view(torch.uint8)
    if masks.dtype == torch.uint8:
        kind = L.MASK_U8
    else:
        masks, kind = masks.to(torch.float32), L.MASK_F32
    masks = masks.contiguous()
    rowcol = _dev_contig(rowcol, torch.int32)
    seg_off = _dev_contig(seg_off, torch.int64)
    assert seg_off.numel() == n_seg + 1
    cap = int(rowcol.shape[0]) if cap is None else int(cap)
    words = mask_words(k)
    member = torch.empty((max(cap, 1), words), dtype=torch.int32, device=dev)
    counts = torch.zeros((n_seg, k), dtype=torch.int32, device=dev) if want_counts else None
    need = L.lib().xm3d_gather_ws_bytes(n_seg, k, h, w)
    if ws is None or ws.numel() < need:
        ws = _ws(need, dev)
    L.check(L.lib().xm3d_gather_masks_batch(_ptr(masks), kind, THR[mode], n_seg, k, h, w, _ptr(rowcol),
                                            _ptr(seg_off), cap, _ptr(member), _ptr(counts), _ptr(ws), ws.numel(),
                                            _stream()))
    return member, counts


def pixel_bits(masks: torch.Tensor, mode: str = "ge0.5", ws: Optional[torch.Tensor] = None) -> "PreparedMasks":
    """masks [n_seg,k,h,w] bool/uint8/float32 -> per-pixel membership words (first half of gather_masks; it
    does not depend on the projection).  Feed the result to point_bits."""
    _require_cuda()
    assert masks.dim() == 4
    n_seg, k, h, w = masks.shape
    if masks.dtype == torch.bool:
        masks = masks.view(torch.uint8)
    if masks.dtype == torch.uint8:
        kind = L.MASK_U8
    else:
        masks, kind = masks.to(torch.float32), L.MASK_F32
    masks = masks.contiguous()
    words = mask_words(k)
    need = n_seg * words * h * w * 4
    if ws is None or ws.numel() < need:
        ws = _ws(need, masks.device)
    L.check(L.lib().xm3d_pixel_bits_batch(_ptr(masks), kind, THR[mode], n_seg, k, h, w, _ptr(ws), _stream()))
    pix = ws[:need].view(torch.int32).view(n_seg, words, h * w)
    return PreparedMasks(pix, None, None, None, k, h, w)


def _popcount32(x: torch.Tensor) -> torch.Tensor:
    x = x.to(torch.int64) & 0xFFFFFFFF
    x = x - ((x >> 1) & 0x55555555)
    x = (x & 0x33333333) + ((x >> 2) & 0x33333333)
    x = (x + (x >> 4)) & 0x0F0F0F0F
    return (x * 0x01010101 >> 24) & 0xFF


def pool(feat: torch.Tensor, seg_off: torch.Tensor, k: int, member: Optional[torch.Tensor] = None,
         label: Optional[torch.Tensor] = None, row_index: Optional[torch.Tensor] = None,
         cap: Optional[int] = None, cap_pairs: Optional[int] = None, want_mean: bool = True,
         ws: Optional[torch.Tensor] = None, status: Optional[torch.Tensor] = None, path: str = "auto", _tune: int = 0):
    """Segmented mean pooling.  feat float32 [rows,c]; member int32 [cap,words] or label int32 [cap].
    cap_pairs: bound on the number of (point, mask) memberships (None: labels -> cap, members ->
    counted on the device, which costs one host sync).
    path: "auto" | "pair_lists" | "rows" | "mma" (include/xm3d.h: XM3D_POOL_*).
    Returns (sum [n_seg,k,c], cnt int32 [n_seg,k], mean [n_seg,k,c] or None)."""
    _require_cuda()
    dev = feat.device
    feat = _dev_contig(feat, torch.float32)
    assert feat.dim() == 2
    c = feat.shape[1]
    seg_off = _dev_contig(seg_off, torch.int64)
    n_seg = seg_off.numel() - 1
    if member is not None:
        member = _dev_contig(member, torch.int32)
        n_pts = member.shape[0]
    else:
        label = _dev_contig(label, torch.int32)
        n_pts = label.shape[0]
    if row_index is not None:
        row_index = _dev_contig(row_index, torch.int32)
    cap = int(n_pts) if cap is None else int(cap)
    if cap_pairs is None:
        cap_pairs = cap if member is None else int(_popcount32(member[:cap]).sum().item())
    cap_pairs = int(cap_pairs)
    s = torch.empty((n_seg, k, c), dtype=torch.float32, device=dev)
    cnt = torch.empty((n_seg, k), dtype=torch.int32, device=dev)
    mean = torch.empty((n_seg, k, c), dtype=torch.float32, device=dev) if want_mean else None
    need = L.lib().xm3d_pool_ws_bytes(n_seg, k, c, cap, cap_pairs)
    if ws is None or ws.numel() < need:
        ws = _ws(need, dev)
    L.check(L.lib().xm3d_pool_batch(_ptr(feat), c, _ptr(row_index), _ptr(member), _ptr(label), n_seg, int(k),
                                    _ptr(seg_off), cap, cap_pairs, POOL_PATH[path] | (int(_tune) << 8), _ptr(s), _ptr(cnt), _ptr(mean),
                                    _ptr(ws), ws.numel(), _ptr(status), _stream()))
    return s, cnt, mean


def scatter(emb: torch.Tensor, seg_off: torch.Tensor, n_pts: int, member: Optional[torch.Tensor] = None,
            label: Optional[torch.Tensor] = None, want_counter: bool = True):
    """Mask -> point scatter-mean.  emb float32 [n_seg,k,c].  Returns (out [n_pts,c], counter [n_pts])."""
    _require_cuda()
    dev = emb.device
    emb = _dev_contig(emb, torch.float32)
    n_seg, k, c = emb.shape
    seg_off = _dev_contig(seg_off, torch.int64)
    if member is not None:
        member = _dev_contig(member, torch.int32)
    else:
        label = _dev_contig(label, torch.int32)
    out = torch.empty((max(n_pts, 1), c), dtype=torch.float32, device=dev)
    counter = torch.empty(max(n_pts, 1), dtype=torch.float32, device=dev) if want_counter else None
    L.check(L.lib().xm3d_scatter_batch(_ptr(member), _ptr(label), n_seg, k, _ptr(seg_off), int(n_pts), _ptr(emb), c,
                                       _ptr(out), _ptr(counter), _stream()))
    return out[:n_pts], (counter[:n_pts] if counter is not None else None)


@dataclass
class ContraSelection:
    counts: torch.Tensor      # int32 [n_seg,k,3]  (points, binary_gt == 0, binary_gt == 1) per mask, after the guard
    kind: torch.Tensor        # int8  [n_seg,k]    0 none, 1 novel candidate, 2 base candidate
    score: torch.Tensor       # float32 [n_seg,k]  mean sigmoid over pixels > 0.5 (candidates; NaN elsewhere)
    sel: torch.Tensor         # int32 [n_seg,5]    masks to pool (original indices), -1 padded
    n_sel: torch.Tensor       # int32 [n_seg]
    sel_member: torch.Tensor  # int32 [cap,1]      bit j: the point lies in sel[., j]  (pool(..., k=5, member=sel_member))


def contra_select(member: torch.Tensor, k: int, binary_gt: torch.Tensor, seg_off: torch.Tensor,
                  mask_logits: torch.Tensor, cap: Optional[int] = None) -> ContraSelection:
    """loss_contra's mask selection (models/utils/criterion.py:80-146) for all scenes of a batch on the device.
    member int32 [cap,words] (sigmoid(mask[:, x, y]) >= 0.5), binary_gt float32 [cap], mask_logits float32
    [n_seg,k,h,w] already up-sampled to cfg.mask_shape."""
    _require_cuda()
    dev = member.device
    member = _dev_contig(member, torch.int32)
    binary_gt = _dev_contig(binary_gt.reshape(-1), torch.float32)
    seg_off = _dev_contig(seg_off, torch.int64)
    mask_logits = _dev_contig(mask_logits, torch.float32)
    n_seg, kk, h, w = mask_logits.shape
    assert kk == k and seg_off.numel() == n_seg + 1
    cap = int(member.shape[0]) if cap is None else int(cap)
    assert binary_gt.numel() >= cap
    counts = torch.empty((n_seg, k, 3), dtype=torch.int32, device=dev)
    kind = torch.empty((n_seg, k), dtype=torch.int8, device=dev)
    score = torch.empty((n_seg, k), dtype=torch.float32, device=dev)
    sel = torch.empty((n_seg, 5), dtype=torch.int32, device=dev)
    n_sel = torch.empty(n_seg, dtype=torch.int32, device=dev)
    sel_member = torch.zeros((max(cap, 1), 1), dtype=torch.int32, device=dev)
    ws = _ws(L.lib().xm3d_contra_ws_bytes(n_seg, k), dev)
    L.check(L.lib().xm3d_contra_select_batch(_ptr(member), int(k), _ptr(binary_gt), _ptr(seg_off), n_seg, cap,
                                             _ptr(mask_logits), h, w, _ptr(counts), _ptr(kind), _ptr(score), _ptr(sel),
                                             _ptr(n_sel), _ptr(sel_member), _ptr(ws), ws.numel(), _stream()))
    return ContraSelection(counts, kind, score, sel, n_sel, sel_member)


# ----------------------------------------------------------------------------- stage 4
def logits(mask_embed: torch.Tensor, text_embed: torch.Tensor, null_embed: torch.Tensor,
           group_sizes: Sequence[int], logit_scale: float, ensemble: str = "max", want_argmax: bool = False,
           ws: Optional[torch.Tensor] = None):
    """mask_embed [...,c]; text_embed [n_text,c]; null_embed [1,c] -> [..., n_groups+1] float32."""
    _require_cuda()
    dev = mask_embed.device
    lead = mask_embed.shape[:-1]
    c = mask_embed.shape[-1]
    me = _dev_contig(mask_embed.reshape(-1, c), torch.float32)
    te = _dev_contig(text_embed, torch.float32)
    ne = _dev_contig(null_embed.reshape(1, c), torch.float32)
    rows, n_text, n_groups = me.shape[0], te.shape[0], len(group_sizes)
    assert ensemble in ("max", "mean")
    assert sum(group_sizes) == n_text, f"{n_text} != {sum(group_sizes)}"
    goff = (C.c_int32 * (n_groups + 1))(*np.concatenate([[0], np.cumsum(group_sizes)]).astype(np.int32).tolist())
    out = torch.empty((rows, n_groups + 1), dtype=torch.float32, device=dev)
    amax = torch.empty(max(rows, 1), dtype=torch.int32, device=dev) if want_argmax else None
    need = L.lib().xm3d_logits_ws_bytes(rows, n_text, c, n_groups)
    if ws is None or ws.numel() < need:
        ws = _ws(need, dev)
    L.check(L.lib().xm3d_logits(_ptr(me), rows, c, _ptr(te), n_text, _ptr(ne), C.cast(goff, C.c_void_p), n_groups,
                                int(ensemble == "mean"), float(logit_scale), _ptr(out), _ptr(amax), _ptr(ws),
                                ws.numel(), _stream()))
    out = out.reshape(*lead, n_groups + 1)
    if want_argmax:
        return out, amax[:rows].reshape(lead)
    return out


# ----------------------------------------------------------------------------- after the path
def point_logits(feat: torch.Tensor, text_embed: torch.Tensor, logit_scale: float,
                 binary: Optional[torch.Tensor] = None, is_base: Optional[torch.Tensor] = None,
                 want_logits: bool = True, want_argmax: bool = True, mask_label: Optional[torch.Tensor] = None,
                 mask_probs: Optional[torch.Tensor] = None, base_ratio: float = 0.0, novel_ratio: float = 0.0):
    """Per-point logits / argmax (run/infer.py:557, 606-640).  feat [n,c] float32 CUDA, text_embed
    [T,c]; binary [n] float32 (the binary head's 0/1 prediction) with is_base [T] bool selects the
    base / novel blending.  mask_label int32 [n] (-1 = in no final mask) with mask_probs [n_masks,T]
    (softmax of the masks' MaskCLIP logits) turns the values into the fused stream's softmax + geometric-mean
    ensemble of run/infer.py:568-600 (needs is_base).
    Returns (logits [n,T] or None, argmax int32 [n] or None)."""
    _require_cuda()
    dev = feat.device
    feat = _dev_contig(feat, torch.float32)
    te = _dev_contig(text_embed, torch.float32)
    n, c = feat.shape
    t = te.shape[0]
    if is_base is not None:
        is_base = _dev_contig(is_base.reshape(-1).to(torch.uint8), torch.uint8)
        assert is_base.numel() == t
    if binary is not None:
        binary = _dev_contig(binary.reshape(-1), torch.float32)
        assert binary.numel() == n and is_base is not None
    n_masks = 0
    if mask_label is not None:
        mask_label = _dev_contig(mask_label.reshape(-1), torch.int32)
        mask_probs = _dev_contig(mask_probs, torch.float32).log()      # the kernel works in the log domain
        n_masks = int(mask_probs.shape[0])
        assert mask_label.numel() == n and mask_probs.shape[1] == t and is_base is not None
    out = torch.empty((n, t), dtype=torch.float32, device=dev) if want_logits else None
    amax = torch.empty(max(n, 1), dtype=torch.int32, device=dev) if want_argmax else None
    ws = _ws(L.lib().xm3d_point_logits_ws_bytes(t, c), dev)
    L.check(L.lib().xm3d_point_logits(_ptr(feat), n, c, _ptr(te), t, float(logit_scale), _ptr(binary), _ptr(is_base),
                                      _ptr(mask_label), _ptr(mask_probs), n_masks, float(base_ratio), float(novel_ratio),
                                      _ptr(out), _ptr(amax), _ptr(ws), ws.numel(), _stream()))
    return out, (amax[:n] if amax is not None else None)



def accumulate_votes(vis_idx: torch.Tensor, seg_off: torch.Tensor, view_pt_off: torch.Tensor, cls: torch.Tensor,
                     votes: torch.Tensor, counter: torch.Tensor, cap: Optional[int] = None):
    """scene_pred[mask_2d, logits_pred] += 1; counter[mask_2d] += 1 (run/infer.py:642-647) for all
    views of a batch at once.  votes int32 [n_scene_pts, T] and counter int32 [n_scene_pts] are
    updated in place."""
    _require_cuda()
    assert votes.dtype == torch.int32 and counter.dtype == torch.int32 and votes.is_contiguous()
    vis_idx = _dev_contig(vis_idx, torch.int32)
    seg_off = _dev_contig(seg_off, torch.int64)
    view_pt_off = _dev_contig(view_pt_off, torch.int64)
    cls = _dev_contig(cls, torch.int32)
    cap = int(cls.shape[0]) if cap is None else int(cap)
    L.check(L.lib().xm3d_vote_batch(_ptr(vis_idx), _ptr(seg_off), seg_off.numel() - 1, cap, _ptr(view_pt_off),
                                    _ptr(cls), votes.shape[1], _ptr(votes), _ptr(counter), _stream()))


def vote_argmax(votes: torch.Tensor, counter: torch.Tensor) -> torch.Tensor:
    """torch.max(scene_pred, dim=1)[1] (run/infer.py:658); -1 where no view saw the point."""
    _require_cuda()
    pred = torch.empty(votes.shape[0], dtype=torch.int32, device=votes.device)
    L.check(L.lib().xm3d_vote_argmax(_ptr(votes), _ptr(counter), votes.shape[0], votes.shape[1], _ptr(pred), _stream()))
    return pred


def nn_fill_match(xyz: torch.Tensor, counter: torch.Tensor, seg_off: Optional[torch.Tensor] = None) -> torch.Tensor:
    """Nearest seen neighbour of every unseen point (run/infer.py:651-656, 684-694: KDTree over the seen
    points, k = 1).  xyz float32 [N,3], counter int32 [N] (seen iff != 0), seg_off int64 [S+1] scene offsets
    (one scene if None).  Returns int64 [N] GLOBAL indices: pred[match] is the filled prediction."""
    _require_cuda()
    dev = xyz.device
    xyz = _dev_contig(xyz, torch.float32)
    counter = _dev_contig(counter.to(torch.int32), torch.int32)
    n = int(xyz.shape[0])
    if seg_off is None:
        seg_off = torch.tensor([0, n], dtype=torch.int64, device=dev)
    seg_off = _dev_contig(seg_off, torch.int64)
    n_seg = seg_off.numel() - 1
    match = torch.empty(max(n, 1), dtype=torch.int32, device=dev)
    ws = _ws(L.lib().xm3d_nn_fill_ws_bytes(n_seg, n), dev)
    L.check(L.lib().xm3d_nn_fill_batch(_ptr(xyz), _ptr(counter), _ptr(seg_off), n_seg, n, _ptr(match), _ptr(ws),
                                       ws.numel(), _stream()))
    match = match[:n].to(torch.int64)
    seg = torch.bucketize(torch.arange(n, device=dev), seg_off[1:-1], right=True) if n_seg > 1 else None
    base = seg_off[seg] if seg is not None else seg_off[0]
    return torch.where(match >= 0, match + base, torch.arange(n, device=dev))


def nn_fill(pred: torch.Tensor, xyz: torch.Tensor, counter: torch.Tensor,
            seg_off: Optional[torch.Tensor] = None) -> torch.Tensor:
    """scene_pred[false_idx] = scene_pred[true_idx[nearest]] (run/infer.py:686-694); pred [N] or [N, ...]."""
    return pred[nn_fill_match(xyz, counter, seg_off)]


def segment_max(feat: torch.Tensor, seg_off: torch.Tensor) -> torch.Tensor:
    """torch.stack([feat[seg_off[s]:seg_off[s+1]].max(0)[0] for s ...]) (models/xmask3d.py:154-159)."""
    _require_cuda()
    feat = _dev_contig(feat, torch.float32)
    seg_off = _dev_contig(seg_off, torch.int64)
    n_seg = seg_off.numel() - 1
    out = torch.empty((n_seg, feat.shape[1]), dtype=torch.float32, device=feat.device)
    L.check(L.lib().xm3d_segment_max(_ptr(feat), _ptr(seg_off), n_seg, int(feat.shape[1]), _ptr(out), _stream()))
    return out


# ----------------------------------------------------------------------------- after the path: mask preparation
@dataclass
class PreparedMasks:
    pixbits: Optional[torch.Tensor]    # int32 (uint32 bits) [n_seg, words, h*w]
    label: Optional[torch.Tensor]      # int16 [n_seg, h, w]; -1 = in no final mask
    areas: Optional[torch.Tensor]      # int32 [n_seg, k, 3]: mask_area, original_area, intersection
    upsampled: Optional[torch.Tensor]  # float32 [n_seg, k, h, w]
    k: int
    h: int
    w: int


def mask_prep(logits_lowres: torch.Tensor, size, scores: Optional[torch.Tensor] = None,
              keep: Optional[torch.Tensor] = None, mode: str = "sigmoid_gt0.5", want_bits: bool = True,
              want_partition: bool = False, want_upsampled: bool = False) -> PreparedMasks:
    """Bilinear upsample (align_corners=False) + sigmoid / threshold (+ score-weighted argmax partition
    with its areas) of low-resolution mask logits [n_seg,k,hs,ws], one dense pass, nothing of size
    k*h*w written unless want_upsampled (reference models/xmask3d.py:326-331, 356-358, 391-435)."""
    _require_cuda()
    dev = logits_lowres.device
    lg = _dev_contig(logits_lowres, torch.float32)
    assert lg.dim() == 4
    n_seg, k, hs, ws_ = lg.shape
    h, w = int(size[0]), int(size[1])
    words = mask_words(k)
    sc = _dev_contig(scores.reshape(n_seg, k), torch.float32) if scores is not None else None
    kp = _dev_contig(keep.reshape(n_seg, k).to(torch.uint8), torch.uint8) if keep is not None else None
    pixbits = torch.empty((n_seg, words, h * w), dtype=torch.int32, device=dev) if want_bits else None
    label = torch.empty((n_seg, h, w), dtype=torch.int16, device=dev) if want_partition else None
    areas = torch.empty((n_seg, k, 3), dtype=torch.int32, device=dev) if want_partition else None
    up = torch.empty((n_seg, k, h, w), dtype=torch.float32, device=dev) if want_upsampled else None
    L.check(L.lib().xm3d_mask_prep_batch(_ptr(lg), n_seg, k, hs, ws_, h, w, _ptr(sc), _ptr(kp), THR[mode], _ptr(pixbits),
                                         _ptr(label), _ptr(areas), _ptr(up), _stream()))
    return PreparedMasks(pixbits, label, areas, up, k, h, w)


def point_bits(prep: PreparedMasks, rowcol: torch.Tensor, seg_off: torch.Tensor, cap: Optional[int] = None,
               want_counts: bool = False):
    """member words of every visible point from PreparedMasks.pixbits (mask[:, x_label, y_label] > thr)."""
    _require_cuda()
    dev = prep.pixbits.device
    rowcol = _dev_contig(rowcol, torch.int32)
    seg_off = _dev_contig(seg_off, torch.int64)
    n_seg = prep.pixbits.shape[0]
    cap = int(rowcol.shape[0]) if cap is None else int(cap)
    member = torch.empty((max(cap, 1), mask_words(prep.k)), dtype=torch.int32, device=dev)
    counts = torch.empty((n_seg, prep.k), dtype=torch.int32, device=dev) if want_counts else None
    L.check(L.lib().xm3d_point_bits_batch(_ptr(prep.pixbits), n_seg, prep.k, prep.h, prep.w, _ptr(rowcol), _ptr(seg_off),
                                          cap, _ptr(member), _ptr(counts), _stream()))
    return member, counts


def gather_labels(label_img: torch.Tensor, rowcol: torch.Tensor, seg_off: torch.Tensor,
                  cap: Optional[int] = None) -> torch.Tensor:
    """point_label[i] = label_img[seg(i), row_i, col_i] (int32; the `label` input of pool / scatter)."""
    _require_cuda()
    label_img = _dev_contig(label_img, torch.int16)
    n_seg, h, w = label_img.shape
    rowcol = _dev_contig(rowcol, torch.int32)
    seg_off = _dev_contig(seg_off, torch.int64)
    cap = int(rowcol.shape[0]) if cap is None else int(cap)
    out = torch.full((max(cap, 1),), -1, dtype=torch.int32, device=label_img.device)
    L.check(L.lib().xm3d_gather_labels_batch(_ptr(label_img), n_seg, h, w, _ptr(rowcol), _ptr(seg_off), cap, _ptr(out),
                                             _stream()))
    return out


# ----------------------------------------------------------------------------- after the path: batch layout
def collate(proj: Projection, vox: Unique, cap: Optional[int] = None):
    """collation_fn (dataset/data_loader.py:319-357) on the device.  Returns (ori_coords float32 [cap,4],
    coords int32 [cap,4]); rows beyond vis_off[-1] / uniq_off[-1] are unspecified.  inds_reconstruct is
    `vox.inverse` of voxelize_batch(collate=True), x_label / y_label = proj.rowcol[:, 0] / [:, 1]."""
    _require_cuda()
    dev = proj.xyz_vis.device
    n_seg = proj.vis_off.numel() - 1
    cap = int(proj.xyz_vis.shape[0]) if cap is None else int(cap)
    ori = torch.empty((max(cap, 1), 4), dtype=torch.float32, device=dev)
    coords = torch.empty((max(cap, 1), 4), dtype=torch.int32, device=dev)
    L.check(L.lib().xm3d_collate_batch(_ptr(proj.xyz_vis), _ptr(proj.vis_off), _ptr(vox.voxel_xyz), _ptr(vox.uniq_off),
                                       n_seg, cap, _ptr(ori), _ptr(coords), _stream()))
    return ori, coords


def pack_i16(src: torch.Tensor, rows_dev: Optional[torch.Tensor] = None, out: Optional[torch.Tensor] = None,
             status: Optional[torch.Tensor] = None) -> torch.Tensor:
    """int32 [rows, width] -> int16 (host-bound copies of rowcol / voxel coordinates).  rows_dev: int64 device
    scalar (e.g. vis_off[-1:]) bounding the rows actually converted."""
    _require_cuda()
    src = _dev_contig(src, torch.int32)
    rows = int(src.shape[0])
    width = int(src.numel() // max(rows, 1))
    if out is None:
        out = torch.empty(src.shape, dtype=torch.int16, device=src.device)
    L.check(L.lib().xm3d_pack_i16(_ptr(src), _ptr(rows_dev), rows, width, _ptr(out), _ptr(status), _stream()))
    return out


def accept_views(proj: Projection, scene_labels: torch.Tensor, view_pt_off: torch.Tensor,
                 ignore_labels: Sequence[int] = (255,), min_points: int = 400, max_points: int = 65000,
                 min_valid: int = 10) -> torch.Tensor:
    """The loaders' view filter (dataset/data_loader.py:190-199, data_loader_infer.py: a frame is kept iff
    `400 < sum(mask) < 65000` and more than 10 of its visible points carry a non-ignored label), for every
    view of a batch at once and without a host round trip.  scene_labels: integer label per scene point
    [sum N]; view_pt_off int64 [V]: first point of each view's scene in scene_labels.  Returns bool [V].
    Plain torch ops on the projection's compaction outputs (works on CPU tensors too)."""
    n_vis = proj.n_vis.to(torch.int64)
    v = n_vis.numel()
    cap = proj.vis_idx.shape[0]
    dev = n_vis.device
    j = torch.arange(cap, device=dev)
    seg = torch.bucketize(j, proj.vis_off[1:].contiguous(), right=True).clamp_(max=v - 1)
    live = j < proj.vis_off[-1]
    lab = scene_labels[(view_pt_off[seg] + proj.vis_idx.to(torch.int64)).clamp_(0, scene_labels.numel() - 1)]
    ok = live.clone()
    for ig in ignore_labels:
        ok &= lab != ig
    valid = torch.zeros(v, dtype=torch.int64, device=dev).index_add_(0, seg, ok.to(torch.int64))
    return (n_vis > min_points) & (n_vis < max_points) & (valid > min_valid)
