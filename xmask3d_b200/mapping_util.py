"""Drop-in for the reference's `models/utils/mapping_util.py` (:10-39): the fixed ScanNet camera
(640x480 intrinsics rescaled to 320x240, visibility threshold 0.25, 10 px border)."""
from .fusion_util import PointCloudToImageMapper, adjust_intrinsic, make_intrinsic


def getMapping(reseed=True):
    """Like the reference, reseeds the torch / numpy RNGs to 1457 as a side effect (:11-14);
    pass reseed=False to leave the RNG state alone."""
    if reseed:
        import numpy as np
        import torch
        seed = 1457
        torch.manual_seed(seed)
        if torch.cuda.is_available():
            torch.cuda.manual_seed_all(seed)
        np.random.seed(seed)
    img_dim = (320, 240)
    depth_scale = 1000.0  # noqa: F841  (millimetres; the loaders divide the PNG by it)
    fx, fy, mx, my = 577.870605, 577.870605, 319.5, 239.5
    visibility_threshold = 0.25
    cut_num_pixel_boundary = 10
    intrinsic = make_intrinsic(fx=fx, fy=fy, mx=mx, my=my)
    intrinsic = adjust_intrinsic(intrinsic, intrinsic_image_dim=[640, 480], image_dim=img_dim)
    return PointCloudToImageMapper(image_dim=img_dim, intrinsics=intrinsic,
                                   visibility_threshold=visibility_threshold,
                                   cut_bound=cut_num_pixel_boundary)
