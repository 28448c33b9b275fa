"""CPU tests: the oracle's restated third-party semantics against independent implementations (scipy, numpy, cv2),
oracle invariants on seeded sweeps, and the host-only parts of the C ABI.  No GPU needed."""
import ctypes as C

import numpy as np
import pytest


def _cloud(rng, n, extent=20.0):
    p = np.empty((n, 4), np.float32)
    p[:, :3] = rng.uniform(-extent, extent, (n, 3))
    p[:, 3] = rng.integers(0, 16, n)
    return p


# ---------------------------------------------------------------------------------------------- kNN (FLANN contract)
@pytest.mark.parametrize("k", [1, 5])
def test_kdtree_equals_brute_force(orc, k):
    rng = np.random.default_rng(0)
    cloud, q = _cloud(rng, 5000), _cloud(rng, 400)
    ib, db = orc.knn(cloud, q, k, brute=True)
    it, dt = orc.knn(cloud, q, k, brute=False)
    assert np.array_equal(ib, it) and np.array_equal(db, dt)


def test_knn_tie_rule_is_d2_then_index(orc):
    # duplicates produce exact distance ties: the smaller index must come first, in both implementations
    rng = np.random.default_rng(1)
    base = _cloud(rng, 300)
    cloud = np.concatenate([base, base, base])  # every point three times
    q = base[:50] + np.float32(0.001)
    for brute in (True, False):
        idx, d2 = orc.knn(cloud, q, 5, brute=brute)
        assert (np.diff(d2, axis=1) >= 0).all()
        for r in range(q.shape[0]):
            for j in range(4):
                if d2[r, j] == d2[r, j + 1]:
                    assert idx[r, j] < idx[r, j + 1]
        assert np.array_equal(idx[:, :3], np.stack([np.arange(50), np.arange(50) + 300, np.arange(50) + 600], 1))


def test_knn_matches_scipy(orc):
    scipy_spatial = pytest.importorskip("scipy.spatial")
    rng = np.random.default_rng(2)
    cloud, q = _cloud(rng, 8000), _cloud(rng, 500)
    idx, d2 = orc.knn(cloud, q, 5)
    tree = scipy_spatial.cKDTree(cloud[:, :3].astype(np.float64))
    dd, ii = tree.query(q[:, :3].astype(np.float64), k=5)
    assert np.array_equal(idx, ii.astype(np.int32))
    assert np.allclose(np.sqrt(d2), dd, rtol=1e-5, atol=1e-6)


# ---------------------------------------------------------------------------------------------- VoxelGrid (PCL contract)
def _voxel_numpy(pts, leaf):
    inv = np.float32(1.0) / np.float32(leaf)
    ijk = np.floor(pts[:, :3] * inv).astype(np.int64)
    ijk -= ijk.min(0)
    div = ijk.max(0) + 1
    cell = ijk[:, 0] + ijk[:, 1] * div[0] + ijk[:, 2] * div[0] * div[1]
    order = np.lexsort((np.arange(len(cell)), cell))
    out = []
    start = 0
    cs = cell[order]
    for end in list(np.nonzero(np.diff(cs))[0] + 1) + [len(cs)]:
        acc = np.zeros(4, np.float32)
        for i in order[start:end]:
            acc = (acc + pts[i]).astype(np.float32)
        out.append(acc / np.float32(end - start))
        start = end
    return np.array(out, np.float32)


@pytest.mark.parametrize("leaf", [0.2, 0.4])
def test_voxel_grid_matches_numpy_restatement(orc, leaf):
    rng = np.random.default_rng(3)
    pts = _cloud(rng, 3000, extent=3.0)
    assert np.array_equal(orc.voxel_grid(pts, leaf), _voxel_numpy(pts, leaf))


def test_voxel_grid_edge_cases(orc):
    assert orc.voxel_grid(np.zeros((0, 4), np.float32), 0.2).shape == (0, 4)
    one = np.array([[1.5, -2.25, 0.125, 3.0]], np.float32)
    assert np.array_equal(orc.voxel_grid(one, 0.2), one)
    far = np.array([[0, 0, 0, 1], [3000, 3000, 3000, 2]], np.float32)  # > INT_MAX cells: returned unfiltered
    assert np.array_equal(orc.voxel_grid(far, 0.2), far)
    same = np.tile(one, (7, 1))
    assert np.allclose(orc.voxel_grid(same, 0.2), one)


# ---------------------------------------------------------------------------------------------- OpenCV contract
def test_gemm_double_accumulation(orc):
    rng = np.random.default_rng(4)
    A = rng.normal(size=(6, 3000)).astype(np.float32)
    B = A.T.copy()
    ref = (A.astype(np.float64) @ B.astype(np.float64)).astype(np.float32)
    assert np.abs(orc.gemm(A, B) - ref).max() <= np.abs(ref).max() * 2e-7


def test_qr_solve_vs_numpy_and_cv2(orc):
    rng = np.random.default_rng(5)
    for m, n in ((6, 6), (5, 3)):
        for _ in range(20):
            A = rng.normal(size=(m, n)).astype(np.float32)
            b = rng.normal(size=m).astype(np.float32)
            x = orc.qr_solve(A, b)
            ref = np.linalg.lstsq(A.astype(np.float64), b.astype(np.float64), rcond=None)[0]
            assert np.allclose(x, ref, rtol=2e-4, atol=2e-4)
    cv2 = pytest.importorskip("cv2")
    A = rng.normal(size=(5, 3)).astype(np.float32) + np.array([50, 3, 80], np.float32)  # map-scale coordinates
    b = -np.ones((5, 1), np.float32)
    ok, xc = cv2.solve(A, b, flags=cv2.DECOMP_QR)
    assert ok and np.allclose(orc.qr_solve(A, b[:, 0]), xc[:, 0], rtol=1e-3, atol=1e-5)


def test_jacobi_eigen_vs_cv2(orc):
    cv2 = pytest.importorskip("cv2")
    rng = np.random.default_rng(6)
    for n in (3, 6):
        for _ in range(20):
            M = rng.normal(size=(n + 4, n)).astype(np.float32)
            A = (M.T @ M).astype(np.float32)
            W, V = orc.jacobi_eigen(A)
            ok, Wc, Vc = cv2.eigen(A)
            assert (np.diff(W) <= 0).all()  # descending
            assert np.allclose(W, Wc[:, 0], rtol=1e-4, atol=1e-4)
            for r in range(n):  # eigenvectors are rows, equal up to sign
                assert min(np.abs(V[r] - Vc[r]).max(), np.abs(V[r] + Vc[r]).max()) < 5e-3
            assert np.allclose(V @ A @ V.T, np.diag(W), atol=1e-3 * max(1.0, W[0]))


def test_lu_inverse(orc):
    rng = np.random.default_rng(7)
    A = rng.normal(size=(6, 6)).astype(np.float32) + 3 * np.eye(6, dtype=np.float32)
    assert np.allclose(orc.lu_inverse(A) @ A, np.eye(6), atol=1e-4)


# ---------------------------------------------------------------------------------------------- oracle invariants
def test_extract_invariants(orc, sweeps16):
    sr = orc.ScanRegistration()
    f = sr.extract(sweeps16[0])
    n = f["full"].shape[0]
    rings = f["full"][:, 3].astype(np.int32)
    assert n == sweeps16[0].shape[0] and (np.diff(rings) >= 0).all()  # ring-major, nothing dropped with the table angles
    assert f["sharp"].shape[0] <= 16 * 6 * 16 and f["less_sharp"].shape[0] <= 20 * 6 * 16 and f["flat"].shape[0] <= 32 * 6 * 16
    label = sr.ints("label")
    assert (label[5:n - 5] == 2).sum() == f["sharp"].shape[0]
    assert ((label[5:n - 5] == 2) | (label[5:n - 5] == 1)).sum() == f["less_sharp"].shape[0]
    assert (label[5:n - 5] == -1).sum() == f["flat"].shape[0]
    # sharp points are a subsequence of less sharp
    ls = {tuple(p) for p in f["less_sharp"][:, :3]}
    assert all(tuple(p) in ls for p in f["sharp"][:, :3])
    # relative time in [0, 1] (+- a hair) => intensity fraction < 0.1
    frac = f["full"][:, 3] - rings
    assert frac.min() > -1e-3 and frac.max() < 0.1 + 1e-3


def test_true_vlp16_angles_drop_beams(orc):
    from gpscalibration_b200 import SweepGenerator
    xyz = SweepGenerator(sensor=1).sweep(0)[0]
    f = orc.ScanRegistration().extract(xyz)
    rings = set(f["full"][:, 3].astype(np.int32).tolist())
    assert rings == {0, 1, 2, 3, 4, 5, 7, 9, 11, 12, 13, 14, 15}  # +11/+13/+15 deg dropped, rings 6/8/10 empty (SR:301-320)
    assert f["full"].shape[0] < xyz.shape[0]


def test_oracle_pipeline_tracks_motion(orc, sweeps16):
    pipe = orc.Pipeline()
    for k, xyz in enumerate(sweeps16[:8]):
        r = pipe.process(xyz)
        assert r.odom_published == (1 if k > 0 else 0)
        assert r.mapping_ran == (1 if k % 2 == 1 else 0)  # frames 2, 4, 6, ... (skipFrameNum = 1)
    # 1 m/sweep forward (z), 0.02 rad/sweep yaw (ry), the reference's 1.05 fudge on both (LO:1037,1043)
    assert abs(r.odom[5] - 6 * 1.05) < 0.6 and abs(r.odom[1] - 6 * 0.021) < 0.02
    nc, ns = pipe.map_size()
    assert nc > 3000 and ns > 10000


def test_first_frames_state_machine(orc, sweeps16):
    """frame 1 initialises, frame 2 skips the optimisation (LastNum still 0), frame 3 registers (Appendix A)."""
    pipe = orc.Pipeline()
    it = [pipe.process(x).odom_iters for x in sweeps16[:4]]
    assert it[0] == 0 and it[1] == 0 and it[2] > 0 and it[3] > 0


# ---------------------------------------------------------------------------------------------- C ABI (no compute)
def test_abi_exports_every_declared_symbol():
    import re
    import os
    from gpscalibration_b200 import capi
    lib = capi.load_library()
    hdr = open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "include", "loamgpu.h")).read()
    declared = set(re.findall(r"\b(loam_[a-z0-9_]+)\s*\(", hdr))
    assert declared == set(capi.SYMBOLS), declared ^ set(capi.SYMBOLS)
    for s in declared:
        assert hasattr(lib, s), s
    assert lib.loam_strerror(0) == b"ok" and lib.loam_strerror(-3) == b"buffer or capacity too small"


def test_abi_struct_layouts():
    from gpscalibration_b200 import capi
    assert C.sizeof(capi.Counts) == 20 and C.sizeof(capi.OdomResult) == 72 and C.sizeof(capi.MapResult) == 108
    assert C.sizeof(capi.SweepResult) == 20 + 72 + 108 + 4
    p = capi.Params()
    capi.load_library().loam_default_params(C.byref(p))
    assert (p.n_scans, p.ring_mode, p.skip_frame_num) == (16, 0, 1)


def test_host_gn_solve_bitwise_equal_to_oracle(orc):
    """The 6x6 solve / degeneracy projection the host keeps (LO:975-1004) — product copy vs oracle copy."""
    from gpscalibration_b200 import capi
    rng = np.random.default_rng(8)
    for trial in range(30):
        A = rng.normal(size=(300, 6)).astype(np.float32)
        if trial % 2:
            A[:, 1] *= 1e-3
        AtA = orc.gemm(A.T.copy(), A)
        AtB = rng.normal(size=6).astype(np.float32)
        s1, s2 = np.zeros(37, np.float32), np.zeros(37, np.float32)
        for it in (0, 1, 2):
            assert np.array_equal(capi.gn_solve(AtA, AtB, it, 10.0, s1), orc.gn_solve(AtA, AtB, it, 10.0, s2))
            assert np.array_equal(s1, s2)
        if trial % 2:
            assert s1[36] == 1.0  # degenerate direction detected


def test_create_fails_loudly_without_gpu():
    from gpscalibration_b200 import LoamGpu, LoamError
    try:
        import torch
        if torch.cuda.is_available():
            pytest.skip("GPU present")
    except ImportError:
        pass
    with pytest.raises(LoamError):
        LoamGpu(device=0)
