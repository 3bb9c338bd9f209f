"""ORACLE — test infrastructure, NOT part of the product.

CPU restatement of the PL-VIWO visual front end (ov_core::TrackKLT::feed_new_camera + viw::TrackLSD::
feed_new_camera) used only as the checker in tests/, in __graft_entry__.smoke() and as bench.py's CPU
baseline.  The product path (pl-viwo_b200/csrc, the C-ABI in include/plviwo_fe.h) never imports, links or
executes anything in this package, and fails loudly when its CUDA library is missing.

What pins it
------------
The reference is C++ glue around OpenCV; it cannot be compiled in this image (no
OpenCV C++ headers, Eigen, Boost, ROS — SURVEY.md 8c), and it ships no tests or golden vectors (section 4).

* Every pixel-touching operation is a call into the third-party dependency OpenCV (README.md:19 says
  "OpenCV 4.2"; CMake accepts any 4.x).  The SAME library is importable here as ``cv2`` 4.13.0
  (opencv-python-headless), so ``oracle.cvops`` calls the real OpenCV kernels at exactly the reference's call
  sites, and ``oracle.npops`` restates their published arithmetic in NumPy, pinned against cv2 by
  tests/test_oracle_pins.py and by the committed vectors in tests/golden/ (made by tests/golden/make_golden.py).
* ``oracle.frontend`` restates the reference's own glue (state machine, occupancy grids, grid FAST selection,
  ID assignment, line/point association) line by line, citing file:line.
* ``cv::ximgproc::FastLineDetector`` (opencv_contrib) exists neither under /root/reference nor in this image:
  PARITY UNPINNED for the line-segment extractor — ``oracle/csrc/oracle_shim.cpp`` restates the published
  algorithm; its building blocks (Canny, fitLine) are pinned against cv2.
* libstdc++ ``std::sort`` (Grider_GRID.h:128) is executed for real through the same shim.
"""
