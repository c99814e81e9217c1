"""Drop-in for the reference's `models/utils/mapping_util.py` (:10-39): the fixed ScanNet camera.

The colour camera of ScanNet is calibrated at 640x480 (fx = fy = 577.870605, principal point
319.5 / 239.5); the model works on 320x240 images, points closer than 10 px to the border are
dropped and a point counts as visible when its depth agrees with the depth image within 25 %.
"""
from .fusion_util import PointCloudToImageMapper, adjust_intrinsic, make_intrinsic

SCANNET_CAMERA = {
    "calibrated_dim": [640, 480],
    "focal": (577.870605, 577.870605),
    "principal": (319.5, 239.5),
    "image_dim": (320, 240),
    "depth_scale": 1000.0,          # depth PNGs are millimetres; the loaders divide by this
    "visibility_threshold": 0.25,
    "border_px": 10,
    "rng_seed": 1457,
}


def _reseed(seed):
    import numpy as np
    import torch
    torch.manual_seed(seed)
    if torch.cuda.is_available():
        torch.cuda.manual_seed_all(seed)
    np.random.seed(seed)


def getMapping(reseed=True):
    """Mapper for the ScanNet camera.  Like the reference (:11-14) the call reseeds the torch and
    numpy RNGs as a side effect; reseed=False leaves them alone."""
    cam = SCANNET_CAMERA
    if reseed:
        _reseed(cam["rng_seed"])
    k = make_intrinsic(fx=cam["focal"][0], fy=cam["focal"][1], mx=cam["principal"][0], my=cam["principal"][1])
    k = adjust_intrinsic(k, intrinsic_image_dim=cam["calibrated_dim"], image_dim=cam["image_dim"])
    return PointCloudToImageMapper(image_dim=cam["image_dim"], intrinsics=k,
                                   visibility_threshold=cam["visibility_threshold"], cut_bound=cam["border_px"])
