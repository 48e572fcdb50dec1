"""mvsnet_b200 -- B200-native (sm_100a) MVSNet cost-volume hot path behind the reference's Python names.

    mvsnet_b200.homography_warping   get_homographies, tf_transform_homography, homography_warping, ...
    mvsnet_b200.model                inference, inference_mem, get_probability_map(_slice)
    mvsnet_b200.cnn_wrapper.mvsnetworks.RegNetUS0
    mvsnet_b200.engine.HotPath       device-resident whole-path runner (what bench.py times)

All arithmetic runs in lib/libmvsnet_b200.so (hand-written CUDA, C ABI in include/mvsnet_b200.h).
"""
__version__ = "0.1.0"
