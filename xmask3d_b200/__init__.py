"""xmask3d_b200 — B200-native cross-modal correspondence path of XMask3D.

Host modules mirror the reference's own (same names / signatures):
  voxelization_utils  <- dataset/voxelization_utils.py
  voxelizer           <- dataset/voxelizer.py
  fusion_util         <- models/utils/fusion_util.py
  mapping_util        <- models/utils/mapping_util.py
  fuser               <- models/utils/fuser.py (mask_mapper)
  logits              <- XMASK3d.cal_pred_logits, ensemble_logits_with_labels
  pooling             <- the inline masked pooling of criterion.py / xmask3d.py
`ops` is the batched torch front end of the C ABI (include/xm3d.h); `pipeline` chains the
stages for a batch of scenes x views without host round trips.
"""
__version__ = "0.1.0"
