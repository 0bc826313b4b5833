"""Stage-by-stage comparison of the CUDA ORB with OpenCV (run under gpurun)."""
import sys
from pathlib import Path

import cv2
import numpy as np

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from oracle import feature2d_py as f2d  # noqa: E402
from rtabmap_b200 import Engine, synth  # noqa: E402

img = synth.make_image(480, 640, 5)
use_depth = len(sys.argv) > 1 and sys.argv[1] == "depth"
depth = synth.make_depth(480, 640, 15) if use_depth else None
eng = Engine()
got = eng.orb_detect_describe(img[None], depth[None] if use_depth else None, Engine.orb_params(synth.CAMERA_K4, depth_as_mask=use_depth))[0]
want = f2d.detect_describe(img, depth, synth.CAMERA_K4, f2d.OrbParams(depth_as_mask=use_depth))
print("n got/want", len(got[0]), len(want[0]))
stride = 640 * 480 + 320 * 240 + 160 * 120
gray = eng.debug_orb_buffer(0, stride)
l0 = gray[:640 * 480].reshape(480, 640)
l1 = gray[640 * 480:640 * 480 + 320 * 240].reshape(240, 320)
l2 = gray[640 * 480 + 320 * 240:].reshape(120, 160)
r1 = cv2.resize(img, (320, 240), interpolation=cv2.INTER_LINEAR_EXACT)
r2 = cv2.resize(r1, (160, 120), interpolation=cv2.INTER_LINEAR_EXACT)
print("pyramid equal:", np.array_equal(l0, img), np.array_equal(l1, r1), np.array_equal(l2, r2))
counts = eng.debug_orb_buffer(4, 12, np.int32)
print("candidate counts", counts)
cand = eng.debug_orb_buffer(3, 3 * 16384 * 4, np.uint32).reshape(3, 16384)
fd = cv2.FastFeatureDetector_create(20, True)
if use_depth:
    mask0 = f2d.depth_mask(depth, f2d.OrbParams())
for lv, (im, w, h) in enumerate(((img, 640, 480), (r1, 320, 240), (r2, 160, 120))):
    kf = fd.detect(im, None)
    ref = set()
    for k in kf:
        x, y = int(k.pt[0]), int(k.pt[1])
        if 19 <= x < w - 19 and 19 <= y < h - 19:
            ref.add((y * w + x, int(k.response)))
    mine = set((int(c >> 8), int(c & 255)) for c in cand[lv, :counts[lv]])
    print("level", lv, "cv FAST (border-filtered, no mask)", len(ref), "mine", len(mine), "common", len(ref & mine),
          "only mine", list(mine - ref)[:3], "only cv", list(ref - mine)[:3])
if not use_depth:
    import os
    scm = eng.debug_orb_buffer(6, stride)
    if len(scm) == stride:
        sys.path.insert(0, "/tmp")
        a = img.astype(np.int32)
        h_, w_ = a.shape
        OFF16 = [(0,3),(1,3),(2,2),(3,1),(3,0),(3,-1),(2,-2),(1,-3),(0,-3),(-1,-3),(-2,-2),(-3,-1),(-3,0),(-3,1),(-2,2),(-1,3)]
        c = a[3:h_-3, 3:w_-3]
        d = np.stack([c - a[3+dy:h_-3+dy, 3+dx:w_-3+dx] for dx, dy in OFF16], 0)
        d2 = np.concatenate([d, d[:9]], 0)
        bb = np.full(c.shape, -999); bd = np.full(c.shape, -999)
        for s_ in range(16):
            arc = d2[s_:s_+9]
            bb = np.maximum(bb, arc.min(0)); bd = np.maximum(bd, (-arc).min(0))
        m = np.maximum(bb, bd)
        ref_sc = np.zeros((h_, w_), np.int32); ref_sc[3:h_-3, 3:w_-3] = np.where(m > 20, m - 1, 0)
        g_sc = scm[:640*480].reshape(480, 640).astype(np.int32)
        diff = np.argwhere(g_sc != ref_sc)
        print("score map mismatches:", len(diff), "first", diff[:5].tolist(), [(int(g_sc[y, x]), int(ref_sc[y, x])) for y, x in diff[:5]])
        print("mismatch x%16 hist", np.bincount(diff[:, 1] % 16, minlength=16).tolist(), "y%16", np.bincount(diff[:, 0] % 16, minlength=16).tolist())
print("level kp counts", eng.debug_orb_buffer(5, 12, np.int32))
if len(got[0]) == len(want[0]):
    print("kp equal", np.array_equal(got[0][:, [0, 1, 2, 5]], want[0][:, [0, 1, 2, 5]]), "resp", np.array_equal(got[0][:, 4], want[0][:, 4]),
          "angle", np.array_equal(got[0][:, 3], want[0][:, 3]), "desc bad bits", int(np.unpackbits(got[1] ^ want[1]).sum()))
    blur = eng.debug_orb_buffer(2, stride)
    import math
    kf = cv2.getGaussianKernel(7, 2, cv2.CV_32F)
    print("blur L0 equal sepFilter2D:", np.array_equal(blur[:640 * 480].reshape(480, 640), cv2.sepFilter2D(img, cv2.CV_8U, kf, kf, borderType=cv2.BORDER_REFLECT_101)))
