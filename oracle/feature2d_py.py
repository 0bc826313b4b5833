"""CPU oracle of the DETECT stage — TEST INFRASTRUCTURE ONLY (tests/, smoke(), bench.py CPU legs).

rtabmap::Feature2D with Kp/DetectorStrategy=2 (ORB) is a thin wrapper over OpenCV's cv::ORB
(corelib/src/Features2d.cpp:1614 create, :1657 detect, :1714 compute).  OpenCV is a third-party
dependency that is not under /root/reference; the opencv-python 4.13 build of this
image IS that dependency, so the oracle calls it directly and restates only RTAB-Map's own wrapper
logic around it:
  depth -> mask            Feature2D::generateKeypoints        Features2d.cpp:783-808
  strongest-N selection    Feature2D::limitKeypoints           Features2d.cpp:356-399 (multimap walk)
  3-D lifting              util3d::generateKeypoints3DDepth    util3d_features.cpp:67-120
                           util3d::projectDepthTo3D            util3d.cpp:215-245
                           util2d::getDepth (smoothing=true)   util2d.cpp:947-1108
  colour conversion        cv::cvtColor(BGR2GRAY)              Memory.cpp:5447
"""
from __future__ import annotations

from dataclasses import dataclass

import cv2
import numpy as np


@dataclass
class OrbParams:
    n_features: int = 1000      # Kp/MaxFeatures
    n_levels: int = 3           # ORB/NLevels
    scale_factor: float = 2.0   # ORB/ScaleFactor
    edge_threshold: int = 19    # ORB/EdgeThreshold
    fast_threshold: int = 20    # FAST/Threshold
    patch_size: int = 31        # ORB/PatchSize
    min_depth: float = 0.0      # Kp/MinDepth
    max_depth: float = 0.0      # Kp/MaxDepth
    depth_as_mask: bool = True  # Mem/DepthAsMask


def make_orb(p: OrbParams):
    return cv2.ORB_create(p.n_features, p.scale_factor, p.n_levels, p.edge_threshold, 0, 2, cv2.ORB_HARRIS_SCORE, p.patch_size, p.fast_threshold)


def depth_mask(depth: np.ndarray, p: OrbParams) -> np.ndarray:
    if depth.dtype == np.uint16:
        value = np.where((depth > 0) & (depth < 65535), depth.astype(np.float32) * np.float32(0.001), np.float32(0))
    else:
        value = depth.astype(np.float32)
    ok = (value > np.float32(p.min_depth)) & np.isfinite(value)
    if p.max_depth != 0.0:
        ok &= value <= np.float32(p.max_depth)
    return np.where(ok, 255, 0).astype(np.uint8)


def limit_keypoints(kps, max_keypoints: int):
    """Feature2D::limitKeypoints without SSC: multimap<fabs(response), index> walked in reverse."""
    if max_keypoints > 0 and len(kps) > max_keypoints:
        # std::multimap keeps equal keys in insertion order; the reverse walk therefore yields, among equal
        # responses, the LATER index first
        order = sorted(range(len(kps)), key=lambda i: (abs(np.float32(kps[i].response)), i), reverse=True)
        return [kps[i] for i in order[:max_keypoints]]
    return list(kps)


def get_depth(depth: np.ndarray, x: np.float32, y: np.float32, error_ratio=np.float32(0.02)) -> np.float32:
    f = np.float32
    rows, cols = depth.shape
    u = int(f(x) + f(0.5))
    v = int(f(y) + f(0.5))
    if u == cols and x < f(cols):
        u = cols - 1
    if v == rows and y < f(rows):
        v = rows - 1
    if not (0 <= u < cols and 0 <= v < rows):
        return f(0)
    mm = depth.dtype == np.uint16

    def at(vv, uu):
        if mm:
            d = depth[vv, uu]
            return f(d) * f(0.001) if 0 < d < 65535 else f(0)
        return f(depth[vv, uu])

    d0 = at(v, u)
    if d0 == 0 or not np.isfinite(d0):
        return f(0)
    sw = f(0)
    sd = f(0)
    for uu in range(max(u - 1, 0), min(u + 1, cols - 1) + 1):
        for vv in range(max(v - 1, 0), min(v + 1, rows - 1) + 1):
            if uu == u and vv == v:
                continue
            d = at(vv, uu)
            err = f(error_ratio * d0)
            if d != 0 and np.isfinite(d) and abs(f(d - d0)) < err:
                if uu == u or vv == v:
                    sw = f(sw + f(2))
                    d = f(d * f(2))
                else:
                    sw = f(sw + f(1))
                sd = f(sd + d)
    dd = f(d0 * f(4))
    sw = f(sw + f(4))
    return f(f(dd + sd) / sw)


def keypoints_3d(kps, depth: np.ndarray, K4, p: OrbParams) -> np.ndarray:
    """generateKeypoints3DDepth, one camera, depth the size of the image, identity local transform."""
    f = np.float32
    fx, fy, cx, cy = [f(v) for v in K4]
    out = np.full((len(kps), 3), np.nan, np.float32)
    for i, k in enumerate(kps):
        x, y = f(k.pt[0]), f(k.pt[1])
        d = get_depth(depth, x, y)
        if d > 0:
            ccx = cx if cx > 0 else f(depth.shape[1] // 2) - f(0.5)
            ccy = cy if cy > 0 else f(depth.shape[0] // 2) - f(0.5)
            px = f(f(f(x - ccx) * d) / fx)
            py = f(f(f(y - ccy) * d) / fy)
            if (p.min_depth < 0 or d > f(p.min_depth)) and (p.max_depth <= 0 or d <= f(p.max_depth)):
                out[i] = (px, py, d)
    return out


def detect_describe(image: np.ndarray, depth, K4, p: OrbParams = OrbParams()):
    """Memory::createSignature's feature block: gray -> mask -> generateKeypoints -> generateDescriptors ->
    generateKeypoints3D.  Returns (keypoints [n,6] float32 = x,y,size,angle,response,octave; desc [n,32]; xyz [n,3])."""
    gray = cv2.cvtColor(image, cv2.COLOR_BGR2GRAY) if image.ndim == 3 else image
    mask = depth_mask(depth, p) if (depth is not None and p.depth_as_mask) else None
    orb = make_orb(p)
    kps = orb.detect(gray, mask)
    kps = limit_keypoints(kps, p.n_features)
    kps, desc = orb.compute(gray, kps)
    if desc is None:
        desc = np.zeros((0, 32), np.uint8)
    xyz = keypoints_3d(kps, depth, K4, p) if depth is not None else np.full((len(kps), 3), np.nan, np.float32)
    arr = np.array([[k.pt[0], k.pt[1], k.size, k.angle, k.response, k.octave] for k in kps], np.float32).reshape(-1, 6)
    return arr, desc, xyz
