"""ORACLE — test infrastructure, NOT product code.

CPU restatement (numpy / torch-CPU) of the reference's cross-modal correspondence
path, op for op, so that every float->int decision rounds exactly as the reference's
numpy does.  Only `tests/`, `__graft_entry__.smoke()` and `bench.py`'s CPU-baseline /
`--impl reference` legs may import this module; the product package
(`xmask3d_b200`) never does.

Parity status: the reference ships no golden vectors for this path, so
this port is pinned against outputs of the reference's own functions executed in the
build container — `tests/golden/make_golden.py` imports them from /root/reference and
writes `tests/golden/*.npz`; `tests/test_oracle_golden.py` replays them.

Each function cites the reference lines it follows (paths relative to /root/reference).
"""
from __future__ import annotations

import math
from typing import List, Optional, Sequence, Tuple

import numpy as np

FNV_OFFSET = np.uint64(14695981039346656037)
FNV_PRIME = np.uint64(1099511628211)


# --------------------------------------------------------------------------- stage 1
def fnv_hash_vec(arr: np.ndarray) -> np.ndarray:
    """dataset/voxelization_utils.py:6-18 — FNV-1 over whole uint64 words per row."""
    assert arr.ndim == 2
    words = arr.astype(np.uint64)
    h = np.full(words.shape[0], FNV_OFFSET, dtype=np.uint64)
    with np.errstate(over="ignore"):
        for j in range(words.shape[1]):
            h = h * FNV_PRIME
            h = h ^ words[:, j]
    return h


def ravel_hash_vec(arr: np.ndarray) -> np.ndarray:
    """dataset/voxelization_utils.py:21-35 — mixed-radix ravel after min-subtract."""
    assert arr.ndim == 2
    shifted = (arr - arr.min(0)).astype(np.uint64)
    radix = shifted.max(0).astype(np.uint64) + np.uint64(1)
    key = np.zeros(shifted.shape[0], dtype=np.uint64)
    with np.errstate(over="ignore"):
        for j in range(shifted.shape[1] - 1):
            key = (key + shifted[:, j]) * radix[j + 1]
        key = key + shifted[:, -1]
    return key


def sparse_quantize(coords, feats=None, labels=None, ignore_label=255,
                    set_ignore_label_when_collision=False, return_index=False,
                    hash_type="fnv", quantization_size=1):
    """dataset/voxelization_utils.py:38-102."""
    has_lab, has_feat = labels is not None, feats is not None
    if not has_lab and not has_feat:
        return_index = True
    assert hash_type in ("ravel", "fnv"), \
        "Invalid hash_type. Either ravel, or fnv allowed. You put hash_type=" + hash_type
    assert coords.ndim == 2, \
        "The coordinates must be a 2D matrix. The shape of the input is " + str(coords.shape)
    if has_feat:
        assert feats.ndim == 2 and coords.shape[0] == feats.shape[0]
    if has_lab:
        assert coords.shape[0] == len(labels)
    dim = coords.shape[1]
    if np.isscalar(quantization_size):
        q = [quantization_size] * dim
    elif isinstance(quantization_size, (list, tuple, np.ndarray)) or hasattr(quantization_size, "__len__"):
        assert len(quantization_size) == dim, "Quantization size and coordinates size mismatch."
        q = [v for v in quantization_size]
    else:
        raise ValueError("Not supported type for quantization_size.")
    cells = np.floor(coords / np.array(q))
    key = ravel_hash_vec(cells) if hash_type == "ravel" else fnv_hash_vec(cells)
    if has_lab:
        _, first, cnt = np.unique(key, return_index=True, return_counts=True)
        lab = labels[first]
        if set_ignore_label_when_collision:
            lab[cnt > 1] = ignore_label
        if return_index:
            return first, lab
        return cells[first], feats[first], lab
    _, first, inverse = np.unique(key, return_index=True, return_inverse=True)
    if return_index:
        return first, inverse
    if has_feat:
        return cells[first], feats[first]
    return cells[first]


def axis_rotation(axis: np.ndarray, theta: float) -> np.ndarray:
    """dataset/voxelizer.py:7-8 — expm of the skew matrix of the scaled axis."""
    from scipy.linalg import expm, norm
    return expm(np.cross(np.eye(3), axis / norm(axis) * theta))


class Voxelizer:
    """dataset/voxelizer.py:11-132 (ctor :12-30)."""

    def __init__(self, voxel_size=1, clip_bound=None, use_augmentation=False,
                 scale_augmentation_bound=None, rotation_augmentation_bound=None,
                 translation_augmentation_ratio_bound=None, ignore_label=255):
        self.voxel_size = voxel_size
        self.clip_bound = clip_bound
        self.ignore_label = ignore_label
        self.use_augmentation = use_augmentation
        self.scale_augmentation_bound = scale_augmentation_bound
        self.rotation_augmentation_bound = rotation_augmentation_bound
        self.translation_augmentation_ratio_bound = translation_augmentation_ratio_bound

    def get_transformation_matrix(self):
        """:32-58 — same np.random draw order: 3x uniform, shuffle, 1x uniform."""
        m_vox, m_rot = np.eye(4), np.eye(4)
        rot = np.eye(3)
        if self.use_augmentation and self.rotation_augmentation_bound is not None:
            if not hasattr(self.rotation_augmentation_bound, "__iter__"):
                raise ValueError()
            mats = []
            for ax, bound in enumerate(self.rotation_augmentation_bound):
                theta = 0
                axis = np.zeros(3)
                axis[ax] = 1
                if bound is not None:
                    theta = np.random.uniform(*bound)
                mats.append(axis_rotation(axis, theta))
            np.random.shuffle(mats)
            rot = mats[0] @ mats[1] @ mats[2]
        m_rot[:3, :3] = rot
        scale = 1 / self.voxel_size
        if self.use_augmentation and self.scale_augmentation_bound is not None:
            scale *= np.random.uniform(*self.scale_augmentation_bound)
        np.fill_diagonal(m_vox[:3, :3], scale)
        return m_vox, m_rot

    def clip(self, coords, center=None, trans_aug_ratio=None):
        """:60-79."""
        lo = np.min(coords, 0).astype(float)
        hi = np.max(coords, 0).astype(float)
        size = hi - lo
        if center is None:
            center = lo + size * 0.5
        if trans_aug_ratio is not None:
            center += np.multiply(trans_aug_ratio, size)
        lim = self.clip_bound
        keep = np.ones(coords.shape[0], dtype=bool)
        for a in range(3):
            keep &= (coords[:, a] >= (lim[a][0] + center[a])) & (coords[:, a] < (lim[a][1] + center[a]))
        return keep

    def rigid_transformation(self):
        m_vox, m_rot = self.get_transformation_matrix()
        rt = m_vox
        if self.use_augmentation:
            rt = m_rot @ rt
        return rt, m_rot

    def voxelize(self, coords, feats, labels, center=None, link=None, return_ind=False, _rt=None):
        """:81-132.  `_rt=(RT, M_r)` injects the matrices instead of drawing them."""
        assert coords.shape[1] == 3 and coords.shape[0] == feats.shape[0] and coords.shape[0]
        if self.clip_bound is not None:
            ratio = np.zeros(3)
            if self.use_augmentation and self.translation_augmentation_ratio_bound is not None:
                for ax, b in enumerate(self.translation_augmentation_ratio_bound):
                    ratio[ax] = np.random.uniform(*b)
            keep = self.clip(coords, center, ratio)
            if keep.sum():
                coords, feats = coords[keep], feats[keep]
                if labels is not None:
                    labels = labels[keep]
        rt, m_rot = self.rigid_transformation() if _rt is None else _rt
        homo = np.hstack((coords, np.ones((coords.shape[0], 1), dtype=coords.dtype)))
        grid = np.floor(homo @ rt.T[:, :3])
        grid = np.floor(grid - grid.min(0))
        first, inverse = sparse_quantize(grid, return_index=True)
        grid, feats, labels = grid[first], feats[first], labels[first]
        if feats.shape[1] > 6:
            feats[:, 3:6] = feats[:, 3:6] @ (m_rot[:3, :3].T)
        if return_ind:
            return grid, feats, labels, np.array(inverse), first
        if link is not None:
            return grid, feats, labels, np.array(inverse), link[first]
        return grid, feats, labels, np.array(inverse)


# --------------------------------------------------------------------------- stage 2
def make_intrinsic(fx, fy, mx, my):
    """models/utils/fusion_util.py:7-15."""
    k = np.eye(4)
    k[0][0], k[1][1], k[0][2], k[1][2] = fx, fy, mx, my
    return k


def adjust_intrinsic(intrinsic, intrinsic_image_dim, image_dim):
    """models/utils/fusion_util.py:18-33."""
    if intrinsic_image_dim == image_dim:
        return intrinsic
    rw = int(math.floor(image_dim[1] * float(intrinsic_image_dim[0]) / float(intrinsic_image_dim[1])))
    intrinsic[0, 0] *= float(rw) / float(intrinsic_image_dim[0])
    intrinsic[1, 1] *= float(image_dim[1]) / float(intrinsic_image_dim[1])
    intrinsic[0, 2] *= float(image_dim[0] - 1) / float(intrinsic_image_dim[0] - 1)
    intrinsic[1, 2] *= float(image_dim[1] - 1) / float(intrinsic_image_dim[1] - 1)
    return intrinsic


class PointCloudToImageMapper:
    """models/utils/fusion_util.py:36-142."""

    def __init__(self, image_dim, visibility_threshold=0.25, cut_bound=0, intrinsics=None):
        self.image_dim = image_dim
        self.vis_thres = visibility_threshold
        self.cut_bound = cut_bound
        self.intrinsics = intrinsics

    def compute_mapping(self, camera_to_world, coords, depth=None, intrinsic=None):
        """:46-142 — returns int64 [N,3] rows (pixel row, pixel col, visible)."""
        if self.intrinsics is not None:
            intrinsic = self.intrinsics
        n = coords.shape[0]
        out = np.zeros((3, n), dtype=int)
        homo = np.concatenate([coords, np.ones([n, 1])], axis=1).T
        assert homo.shape[0] == 4, "[!] Shape error"
        w2c = np.linalg.inv(camera_to_world)
        p = np.matmul(w2c, homo)
        zdiv = p[2].copy()
        zdiv[np.abs(zdiv) < 1e-8] = 1.0
        px = (p[0] * intrinsic[0][0]) / zdiv + intrinsic[0][2]
        py = (p[1] * intrinsic[1][1]) / zdiv + intrinsic[1][2]
        ix = np.round(px).astype(int)
        iy = np.round(py).astype(int)
        cut = self.cut_bound
        keep = (p[2] > 0) & (ix >= cut) & (iy >= cut) & \
               (ix < self.image_dim[0] - cut) & (iy < self.image_dim[1] - cut)
        if depth is not None and np.any(keep):
            ky, kx, kz = iy[keep], ix[keep], p[2][keep]
            in_depth = (ky >= 0) & (ky < depth.shape[0]) & (kx >= 0) & (kx < depth.shape[1])
            if np.any(in_depth):
                d = depth[ky[in_depth], kx[in_depth]]
                ok = np.abs(d - kz[in_depth]) <= self.vis_thres * d
                survivors = np.where(keep)[0][in_depth][ok]
                keep = np.zeros_like(keep)
                keep[survivors] = True
        out[0, keep] = iy[keep]
        out[1, keep] = ix[keep]
        out[2, keep] = 1
        return out.T


def getMapping():
    """models/utils/mapping_util.py:10-39 (without the RNG reseeding side effect)."""
    k = adjust_intrinsic(make_intrinsic(577.870605, 577.870605, 319.5, 239.5), [640, 480], (320, 240))
    return PointCloudToImageMapper(image_dim=(320, 240), intrinsics=k,
                                   visibility_threshold=0.25, cut_bound=10)


def compact_mapping(mapping: np.ndarray):
    """Caller-side compaction, dataset/data_loader_infer.py:176-182, 263-268:
    visible mask, rows whose three entries are all non-zero, row/col lists."""
    vis = mapping[:, 2]
    nz = np.all(mapping != 0, axis=1)
    kept = mapping[nz]
    x_label = kept[:, 0][kept[:, 0] != 0]
    y_label = kept[:, 1][kept[:, 1] != 0]
    return vis == 1, x_label, y_label


# --------------------------------------------------------------------------- stage 3
def gather_masks(mask, x_label, y_label, mode: str):
    """Mask-at-point gather + threshold.
      mode "ge0.5"          models/utils/fuser.py:16-17       (mask >= 0.5)
      mode "sigmoid_ge0.5"  models/utils/criterion.py:83-85   (sigmoid(mask) >= 0.5)
      mode "sigmoid_gt0.5"  models/xmask3d.py:356-357         (sigmoid(mask) >  0.5)
    mask: torch [K,H,W]; x_label = pixel row, y_label = pixel col (int64). -> bool [K,n]."""
    m = mask[:, x_label, y_label].clone()
    if mode == "ge0.5":
        return m >= 0.5
    if mode == "sigmoid_ge0.5":
        return m.sigmoid() >= 0.5
    if mode == "sigmoid_gt0.5":
        return m.sigmoid() > 0.5
    raise ValueError(mode)


def scatter_mask_embed(member, mask_embed, n_rows_like):
    """models/utils/fuser.py:22-34 (twin: models/xmask3d.py:441-455): mask -> point
    scatter-mean.  member bool [K,n], mask_embed [K,C], returns (feat2d [n,C], counter [n,1])
    where counter already has the 0 -> 1e-5 replacement applied."""
    import torch
    member = member.clone()
    if len(member[torch.sum(member, dim=1) != 0]) == 0:
        member[0][0] = True                                    # fuser.py:19-20
    counter = torch.zeros((n_rows_like.shape[0], 1))
    acc = torch.zeros_like(n_rows_like)
    for row, emb in zip(member, mask_embed):
        if torch.sum(row) == 0:
            continue
        acc[row] += emb
        counter[row] += 1
    counter[counter == 0] = 1e-5
    return acc / counter, counter


def mask_mapper_ref(x_list, y_list, masks, mask_embeds, pred_3ds, fuser, fc1, fc2, caption_contra_2d_pre=True):
    """models/utils/fuser.py:6-53 over the caller's per-scene lists (ragged K, bool or float masks —
    models/utils/criterion.py:262-340 passes K' kept bool masks or K float zero masks): gather + `>= 0.5`,
    the `[0][0] = True` guard, scatter-mean, fusion of covered rows.  CPU torch tensors."""
    import torch
    output, output_2d, output_3d, output_2d_pre = [], [], [], []
    for x_label, y_label, mask, mask_embed, pred_3d in zip(x_list, y_list, masks, mask_embeds, pred_3ds):
        member = gather_masks(mask, x_label, y_label, "ge0.5")
        feat2d, counter = scatter_mask_embed(member, mask_embed, pred_3d)
        covered = torch.sum(counter, dim=1) >= 1
        final = torch.zeros_like(pred_3d)
        final[covered] = fuser(feat2d[covered], pred_3d[covered])
        final[~covered] = pred_3d[~covered]
        output.append(final)
        output_2d.append(fc2(feat2d))
        output_3d.append(fc1(pred_3d))
        if caption_contra_2d_pre:
            output_2d_pre.append(feat2d[covered])
    return output, output_2d, output_3d, output_2d_pre


def masked_mean_pool(feat, member):
    """models/utils/criterion.py:152-157: per mask `feat[member_k].mean(0)`.
    feat torch [n,C] f32, member bool [K,n] -> (mean [K,C] f32, cnt [K] int64).
    Empty masks give NaN rows in torch; we return zeros for them and cnt 0."""
    import torch
    k, c = member.shape[0], feat.shape[1]
    out = torch.zeros(k, c, dtype=feat.dtype)
    cnt = member.sum(1)
    for i in range(k):
        if cnt[i] > 0:
            out[i] = torch.mean(feat[member[i]], dim=0)
    return out, cnt


def masked_score_pool(scores, member):
    """models/xmask3d.py:362-367: sum(score*mask)/(sum(mask)+1e-10) per mask."""
    import torch
    s = scores.view(1, -1) * member
    return torch.sum(s, dim=1) / (torch.sum(member, dim=1) + 1e-10)


def masked_sum_f64(feat: np.ndarray, member: np.ndarray):
    """High-precision truth for the tolerance tests: fp64 sums per mask."""
    m = member.astype(np.float64)
    return m @ feat.astype(np.float64), member.sum(1).astype(np.int64)


def contra_select_ref(x_label, y_label, mask, binary_gt):
    """models/utils/criterion.py:80-136 for ONE scene: which masks `loss_contra` pools.
    mask: torch [K,H,W] float32 logits, already up-sampled to cfg.mask_shape (:53-55); binary_gt torch [n]
    (0 = novel, 1 = base, anything else = ignored).  Returns a dict:
      member   bool [K,n]  sigmoid(mask[:, x, y]) >= 0.5, after the "row 0 all True if no mask has >= 10 points" guard
      keep     bool [K]    masks with >= 10 points (:90)
      kind     int  [K]    0 = neither, 1 = novel candidate (:109-112), 2 = base candidate (:114-117)   (original index)
      score    float [K]   torch.mean(sigmoid(mask)[sigmoid(mask) > 0.5]) of the candidates, NaN elsewhere
      sel      list        ORIGINAL mask indices pooled: up to 4 novel by descending score, then up to 1 base (:124-139)
      sel_kept list        the same as indices into the kept arrays (the reference's final_list_idx)"""
    import torch
    mask_3d = mask[:, x_label, y_label].clone().sigmoid() >= 0.5
    if len(mask_3d[torch.sum(mask_3d, dim=1) >= 10]) == 0:
        mask_3d[0, :] = True
    keep = torch.sum(mask_3d, dim=1) >= 10
    kept = torch.nonzero(keep).flatten().tolist()
    k = mask.shape[0]
    kind = torch.zeros(k, dtype=torch.int64)
    score = torch.full((k,), float("nan"))
    novel, base = [], []
    for j, m in enumerate(kept):
        sig = mask[m].clone().sigmoid()
        gt = binary_gt[mask_3d[m]]
        novel_num = gt.eq(0).sum().item()
        base_num = len(gt) - novel_num
        base_num_ = gt.eq(1).sum().item()
        novel_num_ = len(gt) - base_num_
        if novel_num > 1.8 * base_num and novel_num > 10:
            kind[m] = 1
            score[m] = torch.mean(sig[sig > 0.5])
            novel.append((j, score[m]))
        elif base_num_ > 20 * novel_num_ and base_num_ > 150:
            kind[m] = 2
            score[m] = torch.mean(sig[sig > 0.5])
            base.append((j, score[m]))
    novel = [i for i, _ in sorted(novel, key=lambda t: t[1], reverse=True)][:4]
    base = [i for i, _ in sorted(base, key=lambda t: t[1], reverse=True)][:1]
    sel_kept = novel + base
    return {"member": mask_3d, "keep": keep, "kind": kind, "score": score, "sel": [kept[j] for j in sel_kept],
            "sel_kept": sel_kept}


def loss_contra_ref(x_list, y_list, pred_masks, mask_embeds, clip_mask_embeddings, features_fused, features_3d,
                    binary_gts, mask_shape):
    """models/utils/criterion.py:39-182 (`Criterion.loss_contra`): up-sample, select per scene
    (contra_select_ref), pool the fused / 3D features under the selected masks (`feature[mask].mean(0)`,
    :148-157), cosine loss against the selected MaskCLIP embeddings.  Returns (loss_3d_contra, final_2d_mask,
    embedding_fused, embedding_3d) — the last two are what the reference pools and then only uses inside the loss."""
    import torch
    import torch.nn.functional as F
    masks = F.interpolate(pred_masks, size=tuple(mask_shape), mode="bilinear", align_corners=False)
    cos = torch.nn.CosineSimilarity()
    e_fused, e_3d, e_gt, final_2d_mask = [], [], [], []
    for b, (x, y, mask, ff, f3, clip_e, gt) in enumerate(zip(x_list, y_list, masks, features_fused, features_3d,
                                                             clip_mask_embeddings, binary_gts)):
        r = contra_select_ref(x, y, mask, gt)
        if not r["sel"]:
            continue
        e_gt.append(torch.stack([clip_e[m] for m in r["sel"]]))
        e_fused.append(torch.stack([torch.mean(ff[r["member"][m]], dim=0) for m in r["sel"]]))
        e_3d.append(torch.stack([torch.mean(f3[r["member"][m]], dim=0) for m in r["sel"]]))
        final_2d_mask.append((b, torch.stack([mask[m] for m in r["sel"]])))
    if e_fused:
        e_fused, e_3d, e_gt = torch.cat(e_fused), torch.cat(e_3d), torch.cat(e_gt).detach()
        loss = (1 - cos(e_3d, e_gt)).mean()
    else:
        inv = torch.stack([mask_embeds[-1][0]])
        loss = (1 - cos(inv, inv)).mean()
        e_fused = e_3d = None
    return loss, final_2d_mask, e_fused, e_3d


def scene_mean_ref(features):
    """models/xmask3d.py:239-258: `feature.mean(0)` per scene, stacked (the K = 1 all-ones pooling)."""
    import torch
    return torch.stack([f.mean(0, keepdim=False) for f in features])


# --------------------------------------------------------------------------- stage 4
def ensemble_logits_with_labels(logits, labels: List[List[str]], ensemble_method: str = "max"):
    """models/modeling/meta_arch/helper.py:72-97."""
    import torch
    sizes = [len(g) for g in labels]
    assert logits.shape[-1] == sum(sizes), f"{logits.shape[-1]} != {sum(sizes)}"
    assert ensemble_method in ["mean", "max"]
    out = torch.zeros(*logits.shape[:-1], len(labels), dtype=logits.dtype)
    start = 0
    for i, s in enumerate(sizes):
        chunk = logits[..., start:start + s]
        out[..., i] = chunk.max(dim=-1).values if ensemble_method == "max" else chunk.mean(dim=-1)
        start += s
    return out


def cal_pred_logits(outputs: dict):
    """models/xmask3d.py:129-143."""
    import torch
    import torch.nn.functional as F
    me = F.normalize(outputs["mask_embed"], dim=-1)
    te = F.normalize(outputs["text_embed"], dim=-1)
    scale = outputs["logit_scale"]
    pred = scale * (me @ te.t())
    pred = ensemble_logits_with_labels(pred, outputs["labels"], ensemble_method="max")
    ne = F.normalize(outputs["null_embed"], dim=-1)
    null = scale * (me @ ne.t())
    return torch.cat([pred, null], dim=-1)


# --------------------------------------------------------------------------- after the path
def fused_stream_ref(fused_feature, text_features, logit_scale, final_pred_open_embedding, final_mask_3d,
                     base_category, novel_category, all_category_count, base_ratio, novel_ratio, binary_pred):
    """run/infer.py:516-608, the FUSED stream of the inference loop, statement for statement (torch CPU):
    normalise, logits, softmax; per final mask the geometric-mean ensemble with the mask's MaskCLIP class
    probabilities, `.log()`, base / novel halves through category_overlapping_mask (:577-590); -1e10 masking of the
    other category group, blend with the binary head's prediction (:592-604), argmax (:606).
    fused_feature [n,C], text_features [T,C], final_pred_open_embedding [Kf,C], final_mask_3d bool [Kf,n] (an
    argmax partition: every point in at most one mask), binary_pred [n,1] in {0,1}.
    Returns (logits_pred [n,T] before the argmax, argmax int64 [n])."""
    import torch
    import torch.nn.functional as F
    fused_feature = F.normalize(fused_feature, dim=-1)
    text_features = F.normalize(text_features, dim=-1)
    logits_pred = logit_scale * (fused_feature @ text_features.t())
    final_pred_open_embedding = F.normalize(final_pred_open_embedding, dim=-1)
    final_pred_open_logits = logit_scale * (final_pred_open_embedding @ text_features.t())
    logits_pred = logits_pred.softmax(dim=-1)
    final_pred_open_logits = final_pred_open_logits.softmax(dim=-1)
    overlap = torch.tensor([int(t in base_category) for t in range(all_category_count)], dtype=torch.long)
    for single_mask, open_logit in zip(final_mask_3d, final_pred_open_logits):
        base = (logits_pred[single_mask] ** base_ratio * open_logit ** (1 - base_ratio)).log() * overlap
        novel = (logits_pred[single_mask] ** novel_ratio * open_logit ** (1 - novel_ratio)).log() * (1 - overlap)
        logits_pred[single_mask] = base + novel
    logits_pred_novel = logits_pred.clone()
    logits_pred_base = logits_pred.clone()
    logits_pred_novel[:, base_category] = -1e10
    logits_pred_base[:, novel_category] = -1e10
    logits_pred = binary_pred * logits_pred_base + (1 - binary_pred) * logits_pred_novel
    return logits_pred, torch.max(logits_pred, 1)[1]


def accumulate_votes(scene_pred, counter, mask_2d, logits_pred):
    """run/infer.py:642-647 — one view's votes (numpy, in place)."""
    idx = np.nonzero(mask_2d)[0]
    scene_pred[idx, logits_pred] += 1
    counter[idx] += 1


def vote_argmax(scene_pred):
    """run/infer.py:658 — torch.max(scene_pred, dim=1)[1]: first maximum."""
    return scene_pred.argmax(1)


def nn_fill_match(scene_coords: np.ndarray, counter: np.ndarray) -> np.ndarray:
    """run/infer.py:651-656, 684-686 — index of the nearest seen point for every point (own index for a
    seen point), restated as an exhaustive float64 search: the KD tree of the reference returns the exact
    Euclidean nearest neighbour of the float32 coordinates promoted to float64; ties -> lowest index.
    O(n_unseen x n_seen): small cases only."""
    xyz = scene_coords.astype(np.float64)
    seen = np.nonzero(counter != 0)[0]
    out = np.arange(len(xyz), dtype=np.int64)
    for i in np.nonzero(counter == 0)[0]:
        d = xyz[seen] - xyz[i]
        d2 = (d[:, 0] * d[:, 0] + d[:, 1] * d[:, 1]) + d[:, 2] * d[:, 2]
        out[i] = seen[np.argmin(d2)] if len(seen) else -1
    return out


def segment_max(feat: np.ndarray, idx: np.ndarray) -> np.ndarray:
    """models/xmask3d.py:154-159 — torch.max(imp_condition[_idx_ == scene_idx], dim=0)[0] per scene."""
    return np.stack([feat[idx == s].max(0) for s in np.unique(idx)])


# ----------------------------------------------------------------------------- mask preparation
def mask_prep_ref(mask_pred_lowres, scores, mask_shape, score_thresh=0.0):
    """Restatement of the reference's inline mask preparation for one view (torch CPU ops, same
    order): models/xmask3d.py:326-331 (interpolate), :391-397 (sigmoid, keep), :415-435 (argmax
    partition + area checks); twin models/utils/criterion.py:239-244, 273-320.
    Returns dict(up, sig, keep, ids, final_keep (indices into the unfiltered masks), final_mask,
    areas [K,3], prob)."""
    import torch
    import torch.nn.functional as F
    x = torch.as_tensor(mask_pred_lowres, dtype=torch.float32)
    sc = torch.as_tensor(scores, dtype=torch.float32)
    up = F.interpolate(x[None], size=tuple(mask_shape), mode="bilinear", align_corners=False)[0]
    mask_pred = up.sigmoid()
    keep = sc > score_thresh
    kept = torch.nonzero(keep).flatten()
    cur_scores, cur_masks = sc[keep], mask_pred[keep]
    K = x.shape[0]
    areas = torch.zeros((K, 3), dtype=torch.int64)
    final_keep, final_mask = [], []
    ids = torch.full(tuple(mask_shape), -1, dtype=torch.int64)
    prob = None
    if cur_masks.shape[0] > 0:
        prob = cur_scores.view(-1, 1, 1) * cur_masks
        cur_mask_ids = prob.argmax(0)
        ids = kept[cur_mask_ids]
        for k in range(cur_masks.shape[0]):
            mask_area = int((cur_mask_ids == k).sum().item())
            original_area = int((cur_masks[k] >= 0.5).sum().item())
            mask = (cur_mask_ids == k) & (cur_masks[k] >= 0.5)
            areas[kept[k]] = torch.tensor([mask_area, original_area, int(mask.sum().item())])
            if mask_area > 0 and original_area > 0 and mask.sum().item() > 0:
                if mask_area / original_area <= 0:
                    continue
                final_keep.append(int(kept[k]))
                final_mask.append(mask)
    return {"up": up, "sig": mask_pred, "keep": keep, "ids": ids, "final_keep": final_keep,
            "final_mask": final_mask, "areas": areas, "prob": prob, "kept": kept}


# ----------------------------------------------------------------------------- batch layout
def collation_ref(locals_3d, coords, inds_reconstruct):
    """The index-carrying part of collation_fn (dataset/data_loader.py:337-342, 345-346, 355): every
    sample arrives with a leading column of ones (data_loader.py:262-265, 283-285), which `[:, 0] *= i` turns
    into the batch index; inds_reconstruct is shifted by the voxels of the samples before it."""
    import torch
    loc_out, crd_out, ind_out = [], [], []
    acc = 0
    for i in range(len(coords)):
        l3 = torch.cat((torch.ones(len(locals_3d[i]), 1, dtype=torch.float32),
                        torch.as_tensor(locals_3d[i], dtype=torch.float32)), dim=1)
        cd = torch.cat((torch.ones(len(coords[i]), 1, dtype=torch.int32),
                        torch.as_tensor(coords[i]).to(torch.int32)), dim=1)
        cd[:, 0] *= i
        l3[:, 0] *= i
        ind_out.append(acc + torch.as_tensor(inds_reconstruct[i], dtype=torch.int64))
        acc += cd.shape[0]
        loc_out.append(l3)
        crd_out.append(cd)
    return torch.cat(loc_out), torch.cat(crd_out), torch.cat(ind_out)
