"""Constants shared by make_golden.py and the replay tests: the loader's augmentation
bounds (/root/reference/dataset/point_loader.py:54-60), restated so that replay does not
need the reference tree."""
import numpy as np

SCALE_AUGMENTATION_BOUND = (0.9, 1.1)
ROTATION_AUGMENTATION_BOUND = ((-np.pi / 64, np.pi / 64), (-np.pi / 64, np.pi / 64), (-np.pi, np.pi))
TRANSLATION_AUGMENTATION_RATIO_BOUND = ((-0.2, 0.2), (-0.2, 0.2), (0, 0))


def vox_kwargs(voxel_size=0.02):
    return dict(voxel_size=voxel_size, clip_bound=None, use_augmentation=True,
                scale_augmentation_bound=SCALE_AUGMENTATION_BOUND,
                rotation_augmentation_bound=ROTATION_AUGMENTATION_BOUND,
                translation_augmentation_ratio_bound=TRANSLATION_AUGMENTATION_RATIO_BOUND)
