"""Scan-to-map iteration at map sizes the pipeline tests do not reach (SURVEY 8d cfg 3 / cfg 5): a voxel-filtered 1 M-point
map, ~40 k stack points, un-sharded and split into 2 and 4 x-slabs (+ 1 m halo) -- 5-NN indices equal to the oracle's, the
28 sums equal up to summation order, and the whole Gauss-Newton loop on the device equal to the oracle's loop bit for bit."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

T_TRUE = np.array([0.002, 0.01, -0.003, 0.05, -0.02, 0.08], np.float32)


@pytest.fixture(scope="module")
def big_map():
    from gpscalibration_b200 import mapsynth
    corner_map, surf_map, extent = mapsynth.synth_map(950_000, 50_000)
    cs, ss = mapsynth.synth_queries(corner_map, surf_map, 40_000, T_TRUE)
    assert corner_map.shape[0] + surf_map.shape[0] > 980_000
    return corner_map, surf_map, cs, ss, extent


@pytest.mark.parametrize("world", [1, 2, 4])
def test_sharded_iteration_equals_oracle_on_1M_map(orc, big_map, world):
    import torch
    from gpscalibration_b200 import LoamGpu, sharding
    from helpers.routing import owner_mask
    corner_map, surf_map, cs, ss, extent = big_map
    T = np.array([0.001, 0.004, -0.001, 0.02, -0.01, 0.03], np.float32)
    rAtA, rAtB, rn, rcc, rcs = orc.map_iteration(cs, ss, corner_map, surf_map, T)
    rsum = orc.map_iteration_sums28(cs, ss, corner_map, surf_map, T)
    assert rn > 20_000 and (rcs[:, 0] >= 0).sum() > 20_000
    edges = sharding.slab_edges(-extent, extent, world)
    total = np.zeros(28)
    got_c, got_s = np.full_like(rcc, -2), np.full_like(rcs, -2)
    part = torch.zeros(32, dtype=torch.float64, device="cuda")
    gpu = LoamGpu(max_map_points=1 << 20)
    for r in range(world):
        ic, isf = sharding.shard_indices(corner_map, edges, r), sharding.shard_indices(surf_map, edges, r)
        gpu.shard_set_slab(edges[r], edges[r + 1])
        gpu.map_set_inputs(cs, ss, corner_map[ic], surf_map[isf])  # the whole stack; this rank's slab + halo of the map
        gpu.map_iter_partial(0, T, part.data_ptr())
        total += part[:28].cpu().numpy()
        lc, ls = gpu.map_corr(cs.shape[0], ss.shape[0])
        mc, ms = owner_mask(orc, cs, T, edges, r), owner_mask(orc, ss, T, edges, r)
        got_c[mc] = np.where(lc[mc] >= 0, ic[np.maximum(lc[mc], 0)], -1)  # shard-local -> global map indices
        got_s[ms] = np.where(ls[ms] >= 0, isf[np.maximum(ls[ms], 0)], -1)
    gpu.close()
    assert np.array_equal(got_c, rcc) and np.array_equal(got_s, rcs)  # pointSearchInd, (d2, index) tie rule
    assert int(total[27]) == rn
    assert np.abs(total - rsum).max() <= 1e-11 * np.abs(rsum).max()    # exact products, double sums, another order
    from gpscalibration_b200 import capi
    AtA, AtB, n = capi.finish_reduced(total)
    assert np.abs(AtA - rAtA).max() <= 1e-6 * np.abs(rAtA).max() and np.abs(AtB - rAtB).max() <= 1e-6 * np.abs(rAtB).max()


def test_device_gauss_newton_loop_equals_oracle_loop(orc, big_map):
    """loam_map_optimize (LM:753-1017 in one launch) against the oracle's iteration + host solve, step by step."""
    from gpscalibration_b200 import LoamGpu
    corner_map, surf_map, cs, ss, extent = big_map
    T = np.zeros(6, np.float32)
    state = np.zeros(37, np.float32)
    iters = 0
    for it in range(10):  # LM:753-1017 with the oracle's pieces
        iters = it + 1
        AtA, AtB, n, _, _ = orc.map_iteration(cs, ss, corner_map, surf_map, T)
        if n < 50:
            continue
        X = orc.gn_solve(AtA, AtB, it, 100.0, state)
        T = (T + X).astype(np.float32)
        dR = np.float32(np.sqrt(((X[:3].astype(np.float64) * 180.0 / np.pi) ** 2).sum()))
        dT = np.float32(np.sqrt(((X[3:].astype(np.float64) * 100) ** 2).sum()))
        if dR < 0.05 and dT < 0.05:
            break
    gpu = LoamGpu(max_map_points=1 << 20)
    gpu.map_set_inputs(cs, ss, corner_map, surf_map)
    Tg, ig = gpu.map_optimize(np.zeros(6, np.float32), 10)
    gpu.close()
    assert ig == iters and 2 <= iters <= 10
    assert np.array_equal(Tg, T), (Tg, T)
    assert np.abs(Tg[3:] - T_TRUE[3:]).max() < 5e-3 and np.abs(Tg[:3] - T_TRUE[:3]).max() < 5e-4  # and it converged to the truth


@pytest.fixture(scope="module")
def cfg3_map():
    """SURVEY 8d cfg 3 at spec size: 2.0 M-point voxel-filtered local map (0.4 M corner + 1.6 M surf), stacks of the size an
    HDL-64 sweep leaves after LM:736-747, all within 80 m of the sensor."""
    from gpscalibration_b200 import mapsynth
    corner_map, surf_map, extent = mapsynth.synth_map(1_600_000, 400_000)
    T_true = np.array([0.001, 0.004, -0.001, 0.05, -0.02, 0.08], np.float32)
    cs, ss = mapsynth.synth_queries_local(corner_map, surf_map, 6_400, 10_300, T_true)
    assert corner_map.shape[0] + surf_map.shape[0] > 1_980_000
    return corner_map, surf_map, cs, ss, T_true


def _oracle_loop(orc, cs, ss, corner_map, surf_map):
    T = np.zeros(6, np.float32)
    state = np.zeros(37, np.float32)
    iters = 0
    for it in range(10):  # LM:753-1017 with the oracle's pieces
        iters = it + 1
        AtA, AtB, n, _, _ = orc.map_iteration(cs, ss, corner_map, surf_map, T)
        if n < 50:
            continue
        X = orc.gn_solve(AtA, AtB, it, 100.0, state)
        T = (T + X).astype(np.float32)
        dR = np.float32(np.sqrt(((X[:3].astype(np.float64) * 180.0 / np.pi) ** 2).sum()))
        dT = np.float32(np.sqrt(((X[3:].astype(np.float64) * 100) ** 2).sum()))
        if dR < 0.05 and dT < 0.05:
            break
    return T, iters


def test_cfg3_spec_size_2M_map_equals_oracle(orc, cfg3_map):
    """Index build + one iteration + the whole device loop on the 2 M-point map against the oracle: pointSearchInd equal,
    sums equal up to summation order, final pose and iteration count equal bit for bit."""
    import torch
    from gpscalibration_b200 import LoamGpu
    corner_map, surf_map, cs, ss, T_true = cfg3_map
    T = np.array([0.0005, 0.002, -0.0005, 0.02, -0.01, 0.03], np.float32)
    rAtA, rAtB, rn, rcc, rcs = orc.map_iteration(cs, ss, corner_map, surf_map, T)
    rsum = orc.map_iteration_sums28(cs, ss, corner_map, surf_map, T)
    assert rn > 8_000
    gpu = LoamGpu(max_map_points=1 << 21)
    gpu.map_set_inputs(cs, ss, corner_map, surf_map)
    part = torch.zeros(32, dtype=torch.float64, device="cuda")
    gpu.map_iter_partial(0, T, part.data_ptr())
    total = part[:28].cpu().numpy()
    lc, ls = gpu.map_corr(cs.shape[0], ss.shape[0])
    assert np.array_equal(lc, rcc) and np.array_equal(ls, rcs)
    assert int(total[27]) == rn
    assert np.abs(total - rsum).max() <= 1e-11 * np.abs(rsum).max()
    Tr, ir = _oracle_loop(orc, cs, ss, corner_map, surf_map)
    Tg, ig = gpu.map_optimize(np.zeros(6, np.float32), 10)
    gpu.close()
    assert ig == ir and np.array_equal(Tg, Tr), (ig, ir, Tg, Tr)
    assert np.abs(Tg[3:] - T_true[3:]).max() < 5e-3 and np.abs(Tg[:3] - T_true[:3]).max() < 5e-4


@pytest.mark.parametrize("sub", ["1", "8"])
def test_both_search_layouts_equal_oracle(orc, big_map, sub):
    """The two layouts of the search (eight lanes per query / one thread per query with cell pruning, chosen by stack size)
    forced on the same inputs in a fresh process each: pointSearchInd, sums and the device loop equal to the oracle's."""
    import os, subprocess, sys, json, tempfile
    corner_map, surf_map, cs, ss, extent = big_map
    T = np.array([0.001, 0.004, -0.001, 0.02, -0.01, 0.03], np.float32)
    rAtA, rAtB, rn, rcc, rcs = orc.map_iteration(cs, ss, corner_map, surf_map, T)
    rsum = orc.map_iteration_sums28(cs, ss, corner_map, surf_map, T)
    Tr, ir = _oracle_loop(orc, cs, ss, corner_map, surf_map)
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    with tempfile.TemporaryDirectory() as td:
        np.savez(os.path.join(td, "in.npz"), cm=corner_map, sm=surf_map, cs=cs, ss=ss, T=T)
        code = (
            "import sys, numpy as np, torch\n"
            f"sys.path.insert(0, {root!r})\n"
            "from gpscalibration_b200 import LoamGpu\n"
            f"d = np.load({os.path.join(td, 'in.npz')!r})\n"
            "gpu = LoamGpu(max_map_points=1 << 20)\n"
            "gpu.map_set_inputs(d['cs'], d['ss'], d['cm'], d['sm'])\n"
            "part = torch.zeros(32, dtype=torch.float64, device='cuda')\n"
            "gpu.map_iter_partial(0, d['T'], part.data_ptr())\n"
            "lc, ls = gpu.map_corr(d['cs'].shape[0], d['ss'].shape[0])\n"
            "Tg, ig = gpu.map_optimize(np.zeros(6, np.float32), 10)\n"
            f"np.savez({os.path.join(td, 'out.npz')!r}, total=part[:28].cpu().numpy(), lc=lc, ls=ls, Tg=Tg, ig=ig)\n")
        env = dict(os.environ, LOAM_GN_SUB=sub)
        subprocess.run([sys.executable, "-c", code], check=True, env=env)
        o = np.load(os.path.join(td, "out.npz"))
        assert np.array_equal(o["lc"], rcc) and np.array_equal(o["ls"], rcs)
        assert int(o["total"][27]) == rn and np.abs(o["total"] - rsum).max() <= 1e-11 * np.abs(rsum).max()
        assert int(o["ig"]) == ir and np.array_equal(o["Tg"], Tr)
