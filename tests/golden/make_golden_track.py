"""Generates tests/golden/ref_track_calibration.npz by running the REFERENCE'S OWN CODE (oracle/_ref/libref_tc.so:
track_calibration.cc + weight_calculation.cc compiled unmodified from /root/reference against the MatrixXd shim) through
the LD:57-83 loop on a seeded synthetic SLAM / ENU track pair.  Only runs where /root/reference exists.
    python tests/golden/make_golden_track.py
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def main():
    from oracle import ref
    from test_track_calibration import make_tracks
    assert ref.tc_available(), "build oracle/_ref first (python -c 'import __graft_entry__ as g; g.build()')"
    slam, enu = make_tracks(256, 0xC0FFEE, outliers=6)
    w, cal = ref.tc_long(slam, enu, 5)
    np.savez_compressed(os.path.join(os.path.dirname(os.path.abspath(__file__)), "ref_track_calibration.npz"), slam=slam, enu=enu, w=w, cal=cal)
    print("wrote ref_track_calibration.npz", w[:4], cal[:2])


if __name__ == "__main__":
    main()
