"""Pack the reference's own end-to-end fixture for BASELINE configs[0] — the 84 images of /root/reference/data/samples and the
ground-truth loop-closure matrix data/samples_GT.bmp that tools/ConsoleApp/main.cpp:321,401 reads — into ONE small test fixture,
tests/golden/samples_c1.npz (the GPU box has no /root/reference).  The JPEG files are stored byte for byte (data, not source);
both the CUDA path and the checker decode them with the same cv2.imdecode call inside the test.

    python tests/golden/make_samples_fixture.py
"""
from pathlib import Path

import cv2
import numpy as np

REF = Path("/root/reference/data")
OUT = Path(__file__).resolve().parent / "samples_c1.npz"


def main():
    files = sorted((REF / "samples").glob("*.jpg"), key=lambda p: int(p.stem))  # CameraImages order: 1.jpg, 2.jpg, ...
    blobs = [np.frombuffer(f.read_bytes(), np.uint8) for f in files]
    offsets = np.concatenate([[0], np.cumsum([len(b) for b in blobs])]).astype(np.int64)
    gt = cv2.imread(str(REF / "samples_GT.bmp"), cv2.IMREAD_GRAYSCALE)
    assert gt.shape == (len(files), len(files))
    np.savez_compressed(OUT, jpeg=np.concatenate(blobs), offsets=offsets, gt=gt, names=np.array([f.name for f in files]))
    print(f"{OUT}: {len(files)} images, {OUT.stat().st_size} bytes, GT loop-closure cells: {(gt == 255).sum()}")


if __name__ == "__main__":
    main()
