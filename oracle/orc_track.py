"""ORACLE — TEST INFRASTRUCTURE ONLY.  numpy / plain-Python restatement of the reference's track calibration (N4):

  trackCalibration   src/gpsCalibration/src/gps_calibration/track_calibration.cc   (TC)
  WeightCoeCal       src/gpsCalibration/src/gps_calibration/weight_calculation.cc  (WC)
  longDisTrackPro    src/gpsCalibration/src/long_distance_track_process/long_distance_track_process.cpp (LD:57-83)

fp64 throughout, every sum in the reference's order (np.cumsum adds sequentially; Python floats are IEEE doubles).
Third-party piece: Eigen::JacobiSVD (TC:506) -> orc.svd3 (orc_linalg.h svd3_jacobi).  Eigen's dense products are taken to
accumulate in ascending inner index (cannot be checked here: Eigen is absent from the image).
PARITY PIN: checked bit for bit against the reference's own track_calibration.cc + weight_calculation.cc compiled
against a minimal MatrixXd shim (oracle/ref_build/ref_tc.cpp -> oracle/_ref/libref_tc.so) in tests/test_track_calibration.py,
and against numpy's LAPACK SVD Kabsch solution within 1e-9.
Quirk fence (stated in the product too): WC:18-19 / WC:41-42 read element n of an n-element vector for the last point;
the index is clamped to n - 1 (distance 0, weight 0).  The shim run reproduces that by keeping a copy of the last element
behind the vector's end.
Only tests/ may import this module.
"""
import math

import numpy as np

from . import orc

SPEED = 2.2   # weight_calculation.h:6
DELTA = 0.01  # weight_calculation.h:7


def _seqsum(v):
    """Left-to-right sum starting from 0.0, like `s += v[i]` in a loop."""
    v = np.asarray(v, np.float64)
    return float(np.cumsum(v)[-1]) if v.size else 0.0


def speed_weights(slam):  # WC:4-27
    slam = np.asarray(slam, np.float64)
    n = slam.shape[0]
    w = np.ones(n)
    for i in range(1, n):
        j = min(i + 1, n - 1)  # quirk fence
        dx, dy = slam[j, 0] - slam[i, 0], slam[j, 1] - slam[i, 1]
        w[i] = min(math.sqrt(dx * dx + dy * dy) / SPEED, 1.0)
    return w


def residual_weights(slam, enu, cal):  # WC:30-78
    w = speed_weights(slam)
    enu, cal = np.asarray(enu, np.float64), np.asarray(cal, np.float64)
    for i in range(w.shape[0]):
        dx, dy = enu[i, 0] - cal[i, 0], enu[i, 1] - cal[i, 1]
        w[i] = w[i] * 1.0 / max(DELTA, math.sqrt(dx * dx + dy * dy))
    return w


def _mat3_mul(X, Y):
    Z = np.zeros((3, 3))
    for i in range(3):
        for j in range(3):
            s = X[i, 0] * Y[0, j]
            s += X[i, 1] * Y[1, j]
            s += X[i, 2] * Y[2, j]
            Z[i, j] = s
    return Z


def _det3(m):
    return (m[0, 0] * (m[1, 1] * m[2, 2] - m[1, 2] * m[2, 1]) - m[0, 1] * (m[1, 0] * m[2, 2] - m[1, 2] * m[2, 0])
            + m[0, 2] * (m[1, 0] * m[2, 1] - m[1, 1] * m[2, 0]))


def best_fit_weighted(A, B, w):  # BFTWithWeight TC:366-545; A, B: N x 4 homogeneous rows, columns 0..2 used
    sa = [_seqsum(A[:, j] * w) for j in range(3)]  # TC:417-439
    sb = [_seqsum(B[:, j] * w) for j in range(3)]
    sw = _seqsum(w)
    sa = [x / sw for x in sa]                     # TC:450-456
    sb = [x / sw for x in sb]
    AA = (A[:, :3] - np.array(sa)) * w[:, None]    # TC:489-503
    BB = (B[:, :3] - np.array(sb)) * w[:, None]
    H = np.zeros((3, 3))
    for r in range(3):
        for c in range(3):
            H[r, c] = _seqsum(AA[:, r] * BB[:, c])  # TC:506 H = AA^T * BB
    U, S, V = orc.svd3(H)
    R = _mat3_mul(V, U.T.copy())
    if _det3(R) < 0:                              # TC:514-521
        V[:, 2] = -1 * V[:, 2]
        R = _mat3_mul(V, U.T.copy())
    T = np.eye(4)
    for i in range(3):
        ra = R[i, 0] * sa[0]
        ra += R[i, 1] * sa[1]
        ra += R[i, 2] * sa[2]
        T[i, :3] = R[i]
        T[i, 3] = sb[i] - ra
    return T


def _rows_times_Tt(M, T, ncols):
    out = np.zeros((M.shape[0], ncols))
    for j in range(ncols):
        s = M[:, 0] * T[j, 0]
        for k in range(1, M.shape[1]):
            s = s + M[:, k] * T[j, k]
        out[:, j] = s
    return out


class TrackCalibration:
    def __init__(self, slam, enu, w):  # dataInitial TC:39-96
        slam, enu = np.asarray(slam, np.float64), np.asarray(enu, np.float64)
        self.n = slam.shape[0]
        self.slam = np.ones((self.n, 4))
        self.enu = np.ones((self.n, 4))
        self.slam[:, 0] = slam[:, 0] - slam[0, 0]
        self.slam[:, 1] = slam[:, 1] - slam[0, 1]
        self.x0, self.y0 = enu[0, 0], enu[0, 1]
        self.enu[:, 0] = enu[:, 0] - self.x0
        self.enu[:, 1] = enu[:, 1] - self.y0
        self.w = np.asarray(w, np.float64).copy()
        self.z, self.t = enu[:, 2].copy(), enu[:, 3].copy()

    def do_icp(self):  # icp TC:98-201 + coordRotated TC:583-618
        src = self.slam.copy()
        prev = 0.0
        for _ in range(2):
            dx, dy = src[:, 0] - self.enu[:, 0], src[:, 1] - self.enu[:, 1]
            dist = np.sqrt(dx * dx + dy * dy)
            T = best_fit_weighted(src, self.enu, self.w)
            src = _rows_times_Tt(src, T, 4)
            mean = _seqsum(dist) / self.n
            if abs(prev - mean) < 0.003:
                break
            prev = mean
        self.T = best_fit_weighted(self.slam, src, self.w)
        self.rotated = _rows_times_Tt(self.slam[:, :3], self.T, 2) + self.T[:2, 3]
        return self.T

    def do_calibration(self):  # calibrateGPSWithSLAMTrack TC:631-689
        n = self.n
        out = np.zeros((n, 4))
        for j in range(2):
            S, E = self.rotated[:, j], self.enu[:, j]
            acc = np.zeros(n)
            for i in range(n):           # iCoord loop, vectorised over iNum: same additions in the same order per iNum
                acc = acc + (E[i] - (S[i] - S))
            acc = acc / n
            out[:, j] = (acc + S) / 2.0 + (self.x0 if j == 0 else self.y0)
        out[:, 2], out[:, 3] = self.z, self.t
        return out


def calibrate_long(slam, enu, iterations=5):  # LD:57-83
    w = speed_weights(slam)
    tc = TrackCalibration(slam, enu, w)
    tc.do_icp()
    pro = tc.do_calibration()
    for _ in range(iterations):
        w = residual_weights(slam, enu, pro)
        tc = TrackCalibration(pro, enu, w)
        tc.do_icp()
        pro = tc.do_calibration()
    return w, pro
