#!/usr/bin/env python3
"""Writes the PIZ-compressed OpenEXR fixtures under tests/golden/ with OpenCV's OpenEXR codec (an independent
encoder) and the pixels OpenCV itself decodes from them:

  piz_half_37x45.exr   HALF, odd sizes, < 2^14 distinct 16-bit values -> the 14-bit wavelet butterflies
  piz_float_70x33.exr  FLOAT, two 32-line blocks (the second one short), >= 2^14 distinct values -> 16-bit butterflies
  piz_flat_40x40.exr   constant image: run-length symbol of the Huffman coder

usage: python tools/make_golden_piz.py   (needs cv2 built with OpenEXR)"""
import os
os.environ["OPENCV_IO_ENABLE_OPENEXR"] = "1"
import cv2
import numpy as np

OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")
rng = np.random.default_rng(7272)


def write(name, img, half):
    path = os.path.join(OUT, name + ".exr")
    ok = cv2.imwrite(path, img, [cv2.IMWRITE_EXR_TYPE, cv2.IMWRITE_EXR_TYPE_HALF if half else cv2.IMWRITE_EXR_TYPE_FLOAT,
                                 cv2.IMWRITE_EXR_COMPRESSION, cv2.IMWRITE_EXR_COMPRESSION_PIZ])
    assert ok
    back = cv2.imread(path, cv2.IMREAD_UNCHANGED)
    np.save(os.path.join(OUT, name + ".npy"), back[..., [2, 1, 0, 3]])  # BGRA -> RGBA
    print(name, back.shape, os.path.getsize(path), "bytes")


smooth = lambda h, w: (np.add.outer(np.arange(h), np.arange(w))[..., None] * np.array([0.01, 0.02, 0.03, 0.0]) + np.array([0, 0, 0, 1.0])).astype(np.float32)
write("piz_half_37x45", (smooth(45, 37) + 0.05 * rng.random((45, 37, 4))).astype(np.float32), True)
write("piz_float_70x33", (smooth(33, 70) * rng.random((33, 70, 4)) * 100).astype(np.float32), False)
write("piz_flat_40x40", np.full((40, 40, 4), 0.25, np.float32), True)
