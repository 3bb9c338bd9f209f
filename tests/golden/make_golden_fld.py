"""Pins the FastLineDetector restatement (oracle/csrc/oracle_shim.cpp, PARITY UNPINNED so far) against the real
opencv_contrib implementation, in the first environment that has it.

    python tests/golden/make_golden_fld.py

* Always (re)writes tests/golden/fld_inputs.npz: the half-resolution equalised images the line detector sees (a 320x192 and
  a 640x280 one from the KAIST-shaped generator, plus a line-heavy 640x280 one).  Committed, so that the inputs do not depend
  on the generator's version.
* Where cv2.ximgproc exists (opencv-contrib-python), runs cv::ximgproc::createFastLineDetector(20, 1.414213562f, 50, 50, 3,
  false)->detect — the reference's call, PL-VIWO/src/update/cam/TrackLSD.cpp:200-205 with the parameters of
  TrackLSD.h:269-273 — on each of them and writes tests/golden/fld_golden.npz (segments in detection order, cv2 version).
  tests/test_oracle_pins.py::test_fld_restatement_against_contrib_golden and the GPU test of the same name then compare
  the restatement and the CUDA path with it; both tests are skipped only while the golden file is absent.
The authoring image has opencv-python-headless without contrib, so only the inputs are committed from there.
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
import cv2  # noqa: E402

INPUTS = os.path.join(HERE, "fld_inputs.npz")
GOLDEN = os.path.join(HERE, "fld_golden.npz")
FLD_ARGS = (20, 1.414213562, 50.0, 50.0, 3, False)   # TrackLSD.h:269-273


def make_inputs():
    import plviwo_b200  # noqa: F401
    from plviwo_b200 import synth
    out = {}
    for name, kw in (("kaist_320x192", dict(seed=77, width=640, height=384)),
                     ("kaist_640x280", dict(seed=1000, width=1280, height=560)),
                     ("lines_640x280", dict(seed=1004, width=1280, height=560, line_heavy=True))):
        seq = synth.SynthSequence(n_frames=4, **kw)
        eq = cv2.equalizeHist(seq.frame(1))                                            # TrackLSD.cpp:83
        out[name] = cv2.resize(eq, None, fx=0.5, fy=0.5, interpolation=cv2.INTER_LINEAR)   # TrackLSD.cpp:204
    np.savez_compressed(INPUTS, **out)
    return out


def main():
    inputs = dict(np.load(INPUTS)) if os.path.exists(INPUTS) and "--regen-inputs" not in sys.argv else make_inputs()
    print("inputs:", {k: v.shape for k, v in inputs.items()})
    if not hasattr(cv2, "ximgproc") or not hasattr(cv2.ximgproc, "createFastLineDetector"):
        print("cv2 %s has no ximgproc.createFastLineDetector (opencv_contrib missing): inputs written, no golden segments"
              % cv2.__version__)
        return 1
    fld = cv2.ximgproc.createFastLineDetector(*FLD_ARGS)
    out = {"cv2_version": np.array(cv2.__version__)}
    for name, img in inputs.items():
        lines = fld.detect(np.ascontiguousarray(img))
        out[name] = np.zeros((0, 4), np.float32) if lines is None else np.asarray(lines, np.float32).reshape(-1, 4)
        print(name, len(out[name]), "segments")
    np.savez_compressed(GOLDEN, **out)
    print("wrote", GOLDEN)
    return 0


if __name__ == "__main__":
    sys.exit(main())
