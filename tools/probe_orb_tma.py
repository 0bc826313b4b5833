"""Diagnostic: one small frame through the ORB stage with the TMA-staged FAST kernel, compared with the plain-load kernel.
Run under compute-sanitizer on the GPU box when the TMA path misbehaves:  compute-sanitizer python tools/probe_orb_tma.py"""
import os
import sys
from pathlib import Path

import numpy as np

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from rtabmap_b200 import Engine, synth  # noqa: E402

img = synth.make_image(240, 320, 7)
op = Engine.orb_params(synth.CAMERA_K4, n_features=300)
eng = Engine()
out = eng.orb_detect_describe(img[None], None, op)
print("tma path:", os.environ.get("LCD_ORB_TMA", "1"), "keypoints", len(out[0][0]))
