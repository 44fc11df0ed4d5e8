"""Seeded inputs for the geometry goldens (shared with tools/make_golden_geometry.py)."""
import numpy as np

PC_RANGE = [-51.2, -51.2, -5.0, 51.2, 51.2, 3.0]
BEV = (12, 10)
IMG_SHAPE = (928, 1600, 3)


def rig(bs=2):
    """nuScenes-like 6-camera rig: yaw {0, +-55, +-110, 180} deg, f = 1266 px, pp (816, 491), 1.5 m."""
    rng = np.random.default_rng(0)
    metas = []
    for b in range(bs):
        mats = []
        for yaw in (0.0, 55.0, -55.0, 110.0, -110.0, 180.0):
            th = np.deg2rad(yaw + rng.normal(0, 0.5))
            # lidar (x fwd, y left, z up) -> camera (x right, y down, z fwd), rotated by yaw
            R = np.array([[np.sin(th), -np.cos(th), 0.0], [0.0, 0.0, -1.0], [np.cos(th), np.sin(th), 0.0]])
            t = np.array([0.0, 1.5, 0.0]) + rng.normal(0, 0.05, 3)
            K = np.array([[1266.0, 0, 816.0], [0, 1266.0, 491.0], [0, 0, 1.0]])
            E = np.eye(4)
            E[:3, :3], E[:3, 3] = R, t
            P = np.eye(4)
            P[:3, :3] = K
            mats.append(P @ E)
        metas.append(dict(lidar2img=np.stack(mats), img_shape=[IMG_SHAPE] * 6))
    return metas
