"""Generate tests/golden/line_ref_*.npz from the reference tree's OWN line-descriptor sources.

oracle/_ref/libref_line.so is Thirdparty/line_descriptor/src/{LSDDetector_custom,binary_descriptor_custom}.cpp compiled unmodified
where they lie (oracle/Makefile target `ref`; the OpenCV primitives behind oracle/shim/ - the line segment detector, GaussianBlur,
Sobel - are the oracle's cv2-4.13-pinned restatements).  Committed because the reference does not travel to the GPU box:
  keylines : LSDDetectorC::detect(image, scale 1, one octave) - every KeyLine of the frame, detection order, 68-byte records
  top      : the LINEextractor selection of them (stable sort by response, nfeatures + 1 kept, class_id = rank)
  desc/desvec : BinaryDescriptor::compute on `top` - 32-byte descriptors and the 72-float LBD vectors
Run from the repo root, in the container that has /root/reference:  python tools/gen_golden_line_ref.py
"""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import plslam_b200  # noqa  (synth only; no GPU needed)
from plslam_b200 import synth
import oracle

out = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")
CASES = {"640x480_s1": (640, 480, 1, 200), "640x480_s2": (640, 480, 2, 200), "752x480_s5": (752, 480, 5, 200),
         "1241x376_s4": (1241, 376, 4, 300)}


def select(kl, nfeatures):
    """LineExtractor.cpp:42-67 for min_line_length 0 and more lines than nfeatures: sort by response (ties keep detection order),
    keep nfeatures + 1, class_id = rank."""
    order = np.argsort(-kl["response"], kind="stable")
    top = kl[order][:nfeatures + 1].copy()
    top["class_id"] = np.arange(len(top))
    return top


if __name__ == "__main__":
    assert oracle.ref_line_available(), "needs /root/reference (make -C oracle ref)"
    for name, (w, h, seed, nf) in CASES.items():
        im = synth.synth_frame(w, h, seed)
        kl = oracle.ref_lsd_keylines(im)
        assert len(kl) > nf + 1
        top = select(kl, nf)
        desc, dv = oracle.ref_lbd_compute(im, top, want_float=True)
        np.savez_compressed(os.path.join(out, f"line_ref_{name}.npz"), keylines=kl, top=top, desc=desc, desvec=dv,
                            img_sum=np.int64(im.astype(np.int64).sum()), params=np.array([w, h, seed, nf]))
        print(name, len(kl), len(top))
