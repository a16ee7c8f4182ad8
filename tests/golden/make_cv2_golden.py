"""Generates tests/golden/cv2_golden.npz — the pin for the two OpenCV calls on the hot path.

The reference calls cv::dilate (src/alfa_gs.cpp:750) with the bytes GetUniformKernel really produces
(src/alfa_gs.cpp:630-639: two `float 1.0` stores into a CV_8U column, then `/= 2`) and cv::filter2D
(src/alfa_gs.cpp:710) with the Savitzky-Golay column kernels of src/alfa_gs.cpp:653-699.  OpenCV's C++
library is absent from the build image, but Python cv2 is importable in the authoring container, so
the real OpenCV answers are recorded here once and committed; the oracle (and through it the GPU
path) is checked against them on every run.  Run from the repo root:  python tests/golden/make_cv2_golden.py
"""
import os

import cv2
import numpy as np

OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "cv2_golden.npz")
rng = np.random.Generator(np.random.PCG64(20261018))
blob = {"cv2_version": np.array(cv2.__version__)}


def uniform_kernel_literal(ks):
    """GetUniformKernel(ks, CV_8U) executed on bytes (the second store overruns the ks-byte buffer)."""
    buf = np.zeros(ks + 3, np.uint8)
    one = np.frombuffer(np.float32(1.0).tobytes(), np.uint8)
    buf[0:4] = one                # kernel.at<float>(0, 0) = 1
    buf[ks - 1:ks + 3] = one      # kernel.at<float>(ks-1, 0) = 1   (row step of a ks x 1 u8 Mat is 1 byte)
    k = buf[:ks].reshape(ks, 1).copy()
    # `kernel /= 2` is Mat::convertTo(.., alpha = 0.5): saturate_cast<uchar>(cvRound(v * 0.5))
    return cv2.convertScaleAbs(k, alpha=0.5)


# ---- dilation: kernel sizes the pipeline can produce: max(window - 2, 3) for window 3..15 odd
dil_ks = [3, 5, 7, 9, 11, 13]
for ks in dil_ks:
    kern = uniform_kernel_literal(ks)
    lab = (rng.random((20, 24)) < 0.15).astype(np.uint8)
    lab[0, :4] = 1
    lab[-1, -4:] = 1
    out = cv2.dilate(lab, kern)  # defaults: anchor centre, 1 iteration, BORDER_CONSTANT / default border value
    blob["dil_kernel_%d" % ks] = kern.reshape(-1)
    blob["dil_in_%d" % ks] = lab
    blob["dil_out_%d" % ks] = out
blob["dil_ks"] = np.array(dil_ks)

# ---- Savitzky-Golay filter2D on angle-range images (0..157) and on a rougher 0..1000 image
SG = {5: ([-3, 12, 17, 12, -3], 35.0), 7: ([-2, 3, 6, 7, 6, 3, -2], 21.0),
      9: ([-21, 14, 39, 54, 59, 54, 39, 14, -21], 231.0),
      11: ([-36, 9, 44, 69, 84, 89, 84, 69, 44, 9, -36], 429.0)}
for ws, (num, norm) in SG.items():
    # `kernel /= norm` on a CV_32F Mat is convertTo with alpha = 1/norm evaluated in float
    kern = (np.array(num, np.float32) * np.float32(1.0 / norm)).reshape(ws, 1)
    for tag, hi, shape in (("a", 158, (16, 40)), ("b", 158, (64, 33)), ("c", 1001, (12, 17))):
        img = rng.integers(0, hi, size=shape).astype(np.uint16)
        if tag == "b":  # angle-image-like: mostly small values with zero holes
            img = (img // 6).astype(np.uint16)
            img[rng.random(shape) < 0.1] = 0
        out = cv2.filter2D(img, -1, kern, anchor=(-1, -1), delta=0, borderType=cv2.BORDER_REFLECT101)
        blob["sg_in_%d_%s" % (ws, tag)] = img
        blob["sg_out_%d_%s" % (ws, tag)] = out
    blob["sg_kernel_%d" % ws] = kern.reshape(-1)
blob["sg_ws"] = np.array(sorted(SG))

np.savez_compressed(OUT, **blob)
print("wrote", OUT, os.path.getsize(OUT), "bytes; cv2", cv2.__version__)
for ks in dil_ks:
    print("ks", ks, "kernel bytes", blob["dil_kernel_%d" % ks].tolist())
