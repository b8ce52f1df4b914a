"""CPU tests: the oracle against analytic known answers, brute force and the reference's own kd-tree
(oracle/_ref, built from /root/reference/mloam_loop/.../nanoflann.hpp).  The reference ships no golden
vectors for this path, so known answers are derived analytically with the conventions of the
reference's check() printers (eps 1e-6, right-multiplied deltaQ; lidar_map_factor.hpp:72-120)."""
import math
import os

import numpy as np
import pytest

import oracle_lib as orc
import synthetic as syn


def rand_pose(rng):
    q = rng.normal(size=4)
    return syn.pose7(rng.normal(size=3) * 3, q)


def test_knn_tree_vs_brute_and_reference_nanoflann():
    rng = np.random.default_rng(1)
    m = np.concatenate([rng.uniform(-20, 20, (20000, 3)), np.zeros((20000, 1))], 1).astype(np.float32)
    q = np.concatenate([rng.uniform(-20, 20, (500, 3)), np.zeros((500, 1))], 1).astype(np.float32)
    for k in (1, 5, 10):
        i_t, d_t = orc.knn(m, q, k)
        i_b, d_b = orc.knn(m, q, k, brute=True)
        assert np.array_equal(i_t, i_b) and np.array_equal(d_t, d_b)
        if orc.ref_lib() is not None:  # the reference's vendored nanoflann, compiled where it lies
            i_r, d_r = orc.ref_knn(m, q, k)
            assert np.array_equal(d_t, d_r)
            assert np.array_equal(i_t, i_r)


def test_knn_small_map_missing_slots():
    m = np.array([[0, 0, 0, 0], [1, 0, 0, 0], [0, 2, 0, 0]], np.float32)
    q = np.array([[0.1, 0, 0, 0]], np.float32)
    idx, sqd = orc.knn(m, q, 5)
    assert list(idx[0]) == [0, 1, 2, -1, -1]
    assert np.isinf(sqd[0, 3])


def test_eig3f_known():
    # diag + rotation: eigenvalues known
    rng = np.random.default_rng(2)
    for _ in range(50):
        R = syn.quat_to_mat(rand_pose(rng)[3:])
        lam = np.sort(rng.uniform(0.01, 5, 3))
        A = (R @ np.diag(lam) @ R.T).astype(np.float32)
        w, V = orc.eig3f(A)
        assert np.allclose(w, lam, rtol=2e-5, atol=1e-6)
        assert np.allclose(np.abs(V.T @ R), np.eye(3), atol=2e-3)
        assert np.allclose(V.T @ V, np.eye(3), atol=1e-5)


def test_lsq_plane_known():
    rng = np.random.default_rng(3)
    for _ in range(50):
        n = rng.normal(size=3)
        n /= np.linalg.norm(n)
        d = rng.uniform(1, 20)
        # points on plane n.x + d = 0
        B = np.linalg.svd(n[None, :])[2][1:]
        P = (-d * n)[None, :] + rng.uniform(-1, 1, (5, 2)) @ B
        ok, sol = orc.lsq_plane(P.astype(np.float32))
        assert ok
        # A sol = -1  ->  sol = n/d
        assert np.allclose(sol, n / d, rtol=2e-3, atol=1e-5)
        ref = np.linalg.lstsq(P.astype(np.float32).astype(np.float64), -np.ones(5), rcond=None)[0]
        assert np.allclose(sol, ref, rtol=5e-3, atol=1e-5)
    ok, _ = orc.lsq_plane(np.zeros((5, 3), np.float32))
    assert not ok


def test_voxel_grid_hand_placed():
    # 4 points, leaf 1: two share a voxel (mean), two alone; output ascending voxel index
    pts = np.array([[0.1, 0.1, 0.1, 1], [0.3, 0.5, 0.7, 3], [1.5, 0.2, 0.2, 5], [0.2, 1.6, 0.1, 7]], np.float32)
    out, ok = orc.voxel_grid(pts, 1.0)
    assert ok and out.shape[0] == 3
    assert np.allclose(out[0], [0.2, 0.3, 0.4, 2.0], atol=1e-6)
    assert np.allclose(out[1], pts[2]) and np.allclose(out[2], pts[3])
    out2, _ = orc.voxel_grid(pts, 1.0, intensity_last=True)
    assert out2[0, 3] == 3.0  # VoxelGridCovarianceMLOAM keeps the last point's intensity
    # the 4-point example of mloam_test/src/test_pointiwithcov.cpp:23-40 (leaf 3): all four in one voxel
    pts = np.array([[0, 0, 0, 0], [1, 0, 0, 0], [0, 1, 0, 0], [1, 1, 0, 0]], np.float32)
    out, _ = orc.voxel_grid(pts, 3.0)
    assert out.shape[0] == 1 and np.allclose(out[0, :3], [0.5, 0.5, 0])
    # empty and overflow edge cases
    out, ok = orc.voxel_grid(np.zeros((0, 4), np.float32), 0.2)
    assert ok and out.shape[0] == 0
    big = np.array([[0, 0, 0, 0], [1e6, 1e6, 1e6, 0]], np.float32)
    out, ok = orc.voxel_grid(big, 0.01)
    assert (not ok) and out.shape[0] == 2  # "Leaf size is too small": input copied


@pytest.mark.parametrize("kind", [orc.F_PLANE, orc.F_EDGE, orc.F_EDGE_VEC])
def test_single_pose_factor_jacobians_fd(kind):
    """Analytic Jacobian vs forward differences, exactly the reference's check() convention."""
    rng = np.random.default_rng(10 + kind)
    eps = 1e-6
    for _ in range(20):
        x = rand_pose(rng)
        p = rng.normal(size=3) * 5
        if kind == orc.F_PLANE:
            n = rng.normal(size=3)
            n /= np.linalg.norm(n)
            coeff = np.array([*n, rng.normal()])
        else:
            a = rng.normal(size=3) * 5
            coeff = np.array([*a, *(a + rng.normal(size=3))])
        s = 0.7 if kind != orc.F_EDGE_VEC else 1.0
        r, J = orc.factor_eval(kind, p, coeff, s, x)
        rows = 3 if kind == orc.F_EDGE_VEC else 1
        J = J[: rows * 7].reshape(rows, 7)
        assert np.all(J[:, 6] == 0)
        for k in range(6):
            d = np.zeros(6)
            d[k] = eps
            xp = orc.plus(x, d)  # t += d ; q = q * deltaQ(d)
            rp, _ = orc.factor_eval(kind, p, coeff, s, xp, want_jac=False)
            num = (rp[:rows] - r[:rows]) / eps
            assert np.allclose(num, J[:, k], rtol=1e-4, atol=2e-4), (kind, k, num, J[:, k])


def test_plane_residual_zero_on_plane_and_edge_distance():
    x = syn.pose7([1, 2, 3], syn.quat_from_rpy(0.1, -0.2, 0.3))
    R = syn.quat_to_mat(x[3:])
    p = np.array([0.5, -1.0, 2.0])
    pw = R @ p + x[:3]
    n = np.array([0.0, 0.6, 0.8])
    r, _ = orc.factor_eval(orc.F_PLANE, p, [*n, -n @ pw], 1.0, x)
    assert abs(r[0]) < 1e-12
    # line through pw + (0,0,1)*t shifted by 0.3 in x: distance 0.3
    a = pw + np.array([0.3, 0, 1.0])
    b = pw + np.array([0.3, 0, -1.0])
    r, _ = orc.factor_eval(orc.F_EDGE, p, [*a, *b], 1.0, x)
    assert abs(r[0] - 0.3) < 1e-12
    rv, _ = orc.factor_eval(orc.F_EDGE_VEC, p, [*a, *b], 1.0, x)
    assert abs(np.linalg.norm(rv) - 0.3) < 1e-12


@pytest.mark.parametrize("kind", [orc.F_ODOM_PLANE, orc.F_ODOM_EDGE])
def test_odom_factor_chain_and_jacobians(kind):
    """Three-pose chain: residual equals the single-pose factor at the composed pose; the pose_i block and the
    translation part of the ext block match forward differences.  The ext ROTATION block is the reference's own
    closed form ([R_e p]x instead of R_e [p]x, lidar_pure_odom_factor.hpp:94-95,273-275) and the pivot block
    (:67-70,247-249) are kept verbatim, inconsistencies included (SURVEY.md §7 "quirks"), so they are not
    compared with finite differences."""
    rng = np.random.default_rng(30 + kind)
    eps = 1e-6
    for _ in range(10):
        xp, xi, xe = rand_pose(rng), rand_pose(rng), rand_pose(rng)
        p = rng.normal(size=3) * 4
        if kind == orc.F_ODOM_PLANE:
            n = rng.normal(size=3)
            n /= np.linalg.norm(n)
            coeff = np.array([*n, rng.normal(), 0, 0])
            single = orc.F_PLANE
        else:
            a = rng.normal(size=3) * 5
            coeff = np.array([*a, *(a + rng.normal(size=3))])
            single = orc.F_EDGE
        x = np.concatenate([xp, xi, xe])
        r, J = orc.factor_eval(kind, p, coeff, 1.0, x)
        comp = syn.pose_mul(syn.pose_mul(syn.pose_inv(xp), xi), xe)
        r1, _ = orc.factor_eval(single, p, coeff, 1.0, comp)
        assert abs(r[0] - r1[0]) < 1e-9
        J = J[:21].reshape(3, 7)
        for blk in (1, 2):
            for k in range(6 if blk == 1 else 3):
                d = np.zeros(6)
                d[k] = eps
                xx = x.copy()
                xx[blk * 7:(blk + 1) * 7] = orc.plus(x[blk * 7:(blk + 1) * 7], d)
                rp, _ = orc.factor_eval(kind, p, coeff, 1.0, xx, want_jac=False)
                num = (rp[0] - r[0]) / eps
                assert abs(num - J[blk, k]) < 5e-4 * max(1.0, abs(num)), (kind, blk, k, num, J[blk, k])


def test_plus_and_huber_and_sqrt_info():
    x = syn.pose7([1, 2, 3], syn.quat_from_rpy(0.3, 0.2, 0.1))
    assert np.allclose(orc.plus(x, np.zeros(6)), x, atol=1e-15)
    d = np.array([0.1, -0.2, 0.3, 0.01, 0.02, -0.03])
    out = orc.plus(x, d)
    assert np.allclose(out[:3], x[:3] + d[:3])
    q = syn.quat_mul(x[3:], np.array([*(d[3:] / 2), 1.0]))
    assert np.allclose(out[3:], q / np.linalg.norm(q))
    V = np.diag([1, 1, 0, 1, 1, 1.0])  # degenerate z translation is not updated
    assert orc.plus(x, d, V)[2] == x[2]
    assert np.allclose(orc.huber(0.1, 0.005), [0.005, 1.0])
    s = 0.04
    assert np.allclose(orc.huber(0.1, s), [2 * 0.1 * math.sqrt(s) - 0.01, 0.1 / math.sqrt(s)])
    assert orc.lib().orc_map_sqrt_info(0.0075) == 1.0  # sqrt(1/0.0075)=11.5 >= 3 -> 1
    assert abs(orc.lib().orc_map_sqrt_info(1.0) - 1.0 / 3.0) < 1e-15


def test_eval_degeneracy_remap():
    rng = np.random.default_rng(5)
    Q = np.linalg.qr(rng.normal(size=(6, 6)))[0]
    lam = np.array([5.0, 50.0, 500.0, 1e3, 1e4, 1e5])
    H = Q @ np.diag(lam) @ Q.T
    V, eig, deg = orc.eval_degeneracy(H, 100.0)
    assert deg and np.allclose(eig, lam, rtol=1e-9)
    # V_update projects out the two weakest eigen-directions
    assert np.allclose(V @ Q[:, 0], 0, atol=1e-9) and np.allclose(V @ Q[:, 1], 0, atol=1e-9)
    assert np.allclose(V @ Q[:, 3], Q[:, 3], atol=1e-9)
    V, eig, deg = orc.eval_degeneracy(H, 1.0)
    assert (not deg) and np.allclose(V, np.eye(6))


def test_extract_cloud_structure():
    scene = syn.make_scene()
    pose = syn.trajectory(1)[0]
    cloud, ss, se = syn.make_sweep(scene, pose, 16, 1024, seed=0)
    f = orc.extract_cloud(cloud, ss, se)
    ns = 16
    assert f["corner_points_sharp"].shape[0] <= 2 * 6 * ns
    assert f["corner_points_less_sharp"].shape[0] <= 20 * 6 * ns
    assert f["surf_points_flat"].shape[0] <= 4 * 6 * ns
    lab = f["label"]
    assert (lab == 2).sum() == f["corner_points_sharp"].shape[0]
    assert ((lab == 2) | (lab == 1)).sum() == f["corner_points_less_sharp"].shape[0]
    assert (lab == -1).sum() == f["surf_points_flat"].shape[0]
    # curvature known answer: collinear equally spaced points have zero curvature
    line = np.zeros((40, 4), np.float32)
    line[:, 0] = np.arange(40) * 0.5
    g = orc.extract_cloud(line, np.array([5], np.int32), np.array([34], np.int32))
    assert np.all(g["curvature"][5:35] == 0)
    assert g["corner_points_sharp"].shape[0] == 0 and g["surf_points_flat"].shape[0] == 6 * 4
    # short ring (< 6 usable points) is skipped (feature_extract.cpp:155)
    g = orc.extract_cloud(line[:16], np.array([5], np.int32), np.array([10], np.int32))
    assert sum(g[k].shape[0] for k in ("corner_points_sharp", "surf_points_flat", "surf_points_less_flat")) == 0


def test_scan2map_converges_and_schedules_agree():
    scene = syn.make_scene()
    traj = syn.trajectory(4)
    surf_map, corner_map = syn.make_submap(scene, 50000)
    cloud, ss, se = syn.make_sweep(scene, traj[3], 16, 1024, seed=3)
    f = orc.extract_cloud(cloud, ss, se)
    cs, _ = orc.voxel_grid(f["corner_points_less_sharp"], 0.2, True)
    sf, _ = orc.voxel_grid(f["surf_points_less_flat"], 0.4, True)
    init = syn.perturb_pose(traj[3], np.random.Generator(np.random.PCG64(9)))
    out, st = orc.scan2map(surf_map, corner_map, sf, cs, init)
    assert st["ran"] == 1 and st["n_surf"] > 1000 and st["n_corner"] > 50
    dt, dr = syn.pose_err(out, traj[3])
    dt0, dr0 = syn.pose_err(init, traj[3])
    assert dt < 0.01 and dr < 1e-3 and dt < dt0
    o = orc.default_opts()
    o[orc.O_MAX_OUTER], o[orc.O_MAX_INNER] = 5, 1  # north-star schedule: re-associate every GN iteration
    out2, _ = orc.scan2map(surf_map, corner_map, sf, cs, init, o)
    d2 = syn.pose_err(out, out2)
    assert d2[0] < 2e-3 and d2[1] < 2e-4
    # map-size gate (lidar_mapper_keyframe.cpp:429)
    out3, st3 = orc.scan2map(surf_map[:40], corner_map, sf, cs, init)
    assert st3["ran"] == 0 and np.array_equal(out3, init)


def test_track_cloud_recovers_motion():
    scene = syn.make_scene()
    traj = syn.trajectory(3)
    a, ssa, sea = syn.make_sweep(scene, traj[0], 16, 1024, seed=0)
    b, ssb, seb = syn.make_sweep(scene, traj[1], 16, 1024, seed=1)
    fa, fb = orc.extract_cloud(a, ssa, sea), orc.extract_cloud(b, ssb, seb)
    rel = syn.pose_mul(syn.pose_inv(traj[0]), traj[1])
    out, st = orc.track_cloud(fa["corner_points_less_sharp"], fa["surf_points_less_flat"], fb["corner_points_sharp"],
                              fb["surf_points_flat"], syn.pose7([0, 0, 0], [0, 0, 0, 1]))
    assert st["n_corner"] + st["n_surf"] >= 10
    dt, dr = syn.pose_err(out, rel)
    assert dt < 0.05 and dr < 5e-3


def _odom_problem(rng, n=400):
    """Features consistent with a 3-pose chain: planes/lines through the chained point, plus noise."""
    xp, xi, xe = (rand_pose(rng) for _ in range(3))
    comp = syn.pose_mul(syn.pose_mul(syn.pose_inv(xp), xi), xe)
    R = syn.quat_to_mat(comp[3:])
    pts = rng.normal(size=(n, 3)) * 6
    types = np.array([ord("s") if k % 3 else ord("c") for k in range(n)], np.uint8)
    coeffs = np.zeros((n, 6))
    for k in range(n):
        lp = R @ pts[k] + comp[:3]
        if types[k] == ord("s"):
            nrm = rng.normal(size=3)
            nrm /= np.linalg.norm(nrm)
            coeffs[k, :4] = [*nrm, -nrm @ lp + rng.normal() * 0.01]
        else:
            d = rng.normal(size=3)
            d /= np.linalg.norm(d)
            off = np.cross(d, rng.normal(size=3)) * 0.01
            coeffs[k] = [*(lp + off + 0.1 * d), *(lp + off - 0.1 * d)]
    # float-valued like the reference's PointPlaneFeature (built from float clouds)
    return xp, xi, xe, types, pts.astype(np.float32).astype(np.float64), coeffs.astype(np.float32).astype(np.float64)


@pytest.mark.parametrize("free_mask", [1, 2, 3])
def test_odom_solve_recovers_perturbed_blocks(free_mask):
    rng = np.random.default_rng(70 + free_mask)
    xp, xi, xe, types, pts, coeffs = _odom_problem(rng)
    d = lambda: syn.pose7(rng.normal(size=3) * 0.02, syn.quat_from_rpy(*(rng.normal(size=3) * 0.004)))
    xi0 = syn.pose_mul(xi, d()) if free_mask & 1 else xi
    xe0 = syn.pose_mul(xe, d()) if free_mask & 2 else xe
    oi, oe, st = orc.odom_solve(types, pts, coeffs, xp, xi0, xe0, free_mask, max_it=30)
    assert st["lm_iterations"] >= 2
    if not free_mask & 1:
        assert np.array_equal(oi, xi0)
    if not free_mask & 2:
        assert np.array_equal(oe, xe0)
    # the composed pivot<-sensor transform is what the residuals constrain
    c_true = syn.pose_mul(syn.pose_mul(syn.pose_inv(xp), xi), xe)
    c_est = syn.pose_mul(syn.pose_mul(syn.pose_inv(xp), oi), oe)
    c_ini = syn.pose_mul(syn.pose_mul(syn.pose_inv(xp), xi0), xe0)
    assert syn.pose_err(c_est, c_true)[0] < 0.25 * syn.pose_err(c_ini, c_true)[0]


def test_good_feature_selection_oracle_properties():
    """orc_gf.hpp: selection sizes, determinism per seed, no duplicates, only matched features, greedy > random in log det."""
    rng = np.random.default_rng(5)
    n = 600
    matched = rng.random(n) < 0.7
    jaco = rng.normal(size=(n, 6)) * matched[:, None]
    xyz = np.concatenate([rng.uniform(-20, 20, (n, 3)), np.zeros((n, 1))], 1).astype(np.float32)
    sel_all, H_all = orc.gf_select(orc.GF_WO, 1.0, 1, matched, jaco, xyz)
    assert np.array_equal(sel_all, np.flatnonzero(matched))
    assert np.allclose(H_all, 1e-6 * np.eye(6) + jaco.T @ jaco)
    ld = {}
    for name, m in (("rnd", orc.GF_RND), ("fps", orc.GF_FPS), ("gd", orc.GF_GD)):
        s1, H1 = orc.gf_select(m, 0.2, 42, matched, jaco, xyz)
        s2, H2 = orc.gf_select(m, 0.2, 42, matched, jaco, xyz)
        s3, _ = orc.gf_select(m, 0.2, 43, matched, jaco, xyz)
        assert np.array_equal(s1, s2) and np.array_equal(H1, H2) and not np.array_equal(s1, s3)
        assert len(s1) == int(n * 0.2) and len(set(s1.tolist())) == len(s1) and matched[s1].all()
        ld[name] = np.linalg.slogdet(H1)[1]
        if m != orc.GF_FPS:  # fps never adds its start point to H (reference quirk, lidar_mapper.h:375-379)
            assert np.allclose(H1, 1e-6 * np.eye(6) + jaco[s1].T @ jaco[s1])
    assert ld["gd"] > ld["rnd"]
    # nothing matched: every method terminates empty
    none = np.zeros(n, bool)
    for m in (orc.GF_RND, orc.GF_FPS, orc.GF_GD):
        assert len(orc.gf_select(m, 0.3, 1, none, np.zeros((n, 6)), xyz)[0]) == 0


# ------------------------------------------------------------------------------------------------ committed golden fixtures
GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def test_golden_reference_nanoflann_pins_oracle_knn():
    """tests/golden/knn_nanoflann.npz holds the answers of the REFERENCE's own kd-tree (nanoflann.hpp compiled in place,
    see make_golden.py): the oracle's kNN must reproduce indices and float distances exactly — also on hosts without
    /root/reference."""
    g = np.load(os.path.join(GOLDEN, "knn_nanoflann.npz"))
    for k in (1, 5, 10):
        idx, sqd = orc.knn(g["map"], g["query"], k)
        assert np.array_equal(idx, g[f"idx{k}"]) and np.array_equal(sqd, g[f"sqd{k}"])


def test_golden_oracle_regression_vectors():
    """The oracle reproduces its committed outputs bit for bit (guards the checker itself against accidental change)."""
    g = np.load(os.path.join(GOLDEN, "oracle_small.npz"))
    f = orc.extract_cloud(g["cloud"], g["ss"], g["se"])
    for key, name in (("corner_points_sharp", "sharp"), ("corner_points_less_sharp", "less_sharp"), ("surf_points_flat", "flat"),
                      ("surf_points_less_flat", "less_flat")):
        assert np.array_equal(f[key].view(np.uint32), g[name].view(np.uint32)), key
    cs, _ = orc.voxel_grid(g["less_sharp"], 0.2, True)
    sf, _ = orc.voxel_grid(g["less_flat"], 0.4, True)
    assert np.array_equal(cs.view(np.uint32), g["corner_ds"].view(np.uint32)) and np.array_equal(sf.view(np.uint32), g["surf_ds"].view(np.uint32))
    vs, cfs, nns = orc.match_from_map("s", g["surf_map"], g["surf_ds"], g["init"])
    assert np.array_equal(vs, g["surf_valid"]) and np.array_equal(nns, g["surf_nn"]) and np.array_equal(cfs, g["surf_coeff"])
    o = orc.default_opts()
    o[orc.O_MAX_OUTER], o[orc.O_MAX_INNER] = 3, 4
    pose, st = orc.scan2map(g["surf_map"], g["corner_map"], g["surf_ds"], g["corner_ds"], g["init"], o)
    assert np.array_equal(pose, g["pose"]) and int(st["n_surf"]) == int(g["n_surf"]) and int(st["n_corner"]) == int(g["n_corner"])
    gf = orc.good_features("s", g["surf_map"], g["surf_ds"], g["init"], orc.GF_GD, 0.25, 11)
    assert np.array_equal(gf["sel"], g["gf_sel"])


def test_solver_restatement_converges_to_independent_optimum():
    """The ceres::Solve restatement (orc_solver.hpp) + the factor restatements against an INDEPENDENT solver: with the
    correspondences of one association fixed, scipy's trust-region least squares with the Huber loss minimises the same
    objective (ceres::HuberLoss(a): rho(s) = s | 2a sqrt(s) - a^2  ==  scipy loss='huber', f_scale=a) written directly
    in numpy.  Both must land on the same pose."""
    from scipy.optimize import least_squares
    from scipy.spatial.transform import Rotation

    scene = syn.make_scene()
    traj = syn.trajectory(6)
    surf_map, corner_map = syn.make_submap(scene, 30000)
    cloud, ss, se = syn.make_sweep(scene, traj[4], 16, 512, seed=21)
    f = orc.extract_cloud(cloud, ss, se)
    cs, _ = orc.voxel_grid(f["corner_points_less_sharp"], 0.2, True)
    sf, _ = orc.voxel_grid(f["surf_points_less_flat"], 0.4, True)
    init = syn.perturb_pose(traj[4], np.random.Generator(np.random.PCG64(2)))
    o = orc.default_opts()
    o[orc.O_MAX_OUTER], o[orc.O_MAX_INNER] = 1, 60  # one association, LM to convergence
    pose, st = orc.scan2map(surf_map, corner_map, sf, cs, init, o)
    assert st["n_surf"] > 300 and st["n_corner"] > 50

    vs, cfs, _ = orc.match_from_map("s", surf_map, sf, init)
    vc, cfc, _ = orc.match_from_map("c", corner_map, cs, init)
    ps, ws, ds = sf[vs][:, :3].astype(np.float64), cfs[vs][:, :3], cfs[vs][:, 3]
    pc, la, lb = cs[vc][:, :3].astype(np.float64), cfc[vc][:, :3], cfc[vc][:, 3:6]
    sinfo = orc.map_sqrt_info(0.0075)
    t0, R0 = init[:3], Rotation.from_quat(init[3:7])  # x, y, z, w

    def residuals(x):
        R = (R0 * Rotation.from_rotvec(x[3:6])).as_matrix()
        t = t0 + x[:3]
        rs = sinfo * (np.einsum("ij,ij->i", ws, ps @ R.T + t) + ds)            # plane: s (w.(Rp+t) + d)
        lp = pc @ R.T + t
        rc = sinfo * np.linalg.norm(np.cross(lp - la, lp - lb), axis=1) / np.linalg.norm(la - lb, axis=1)  # edge: s |..x..| / |a-b|
        return np.concatenate([rs, rc])

    sol = least_squares(residuals, np.zeros(6), loss="huber", f_scale=0.1, xtol=1e-14, ftol=1e-14, gtol=1e-14, max_nfev=400)
    R_ref = (R0 * Rotation.from_rotvec(sol.x[3:6]))
    pose_ref = np.concatenate([t0 + sol.x[:3], R_ref.as_quat()])
    dt, dr = syn.pose_err(pose, pose_ref)
    # Ceres' default function tolerance (1e-6 relative cost change) stops the LM a few 1e-5 m short of the exact optimum
    assert dt < 1e-4 and dr < 1e-4, (dt, dr)
    assert dt < 0.05 * syn.pose_err(init, pose_ref)[0]
    # and the objective value agrees: 1/2 sum rho
    r = residuals(sol.x)
    rho = np.where(np.abs(r) <= 0.1, r * r, 2 * 0.1 * np.abs(r) - 0.01)
    assert abs(0.5 * rho.sum() - st["final_cost"]) < 1e-6 * max(1.0, st["final_cost"])
