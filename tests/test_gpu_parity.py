"""End-to-end parity of the CUDA path against the CPU oracle through the reference-facing interface.

Tolerances are BASELINE.json's: pose 1e-4 rad / 1e-4 m, fitness and RMSE 1e-5 (tie-free synthetic inputs).
"""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

ROT_TOL, TRANS_TOL, FIT_TOL = 1e-4, 1e-4, 1e-5
VOXELS, DISTS = [1.0, 0.5, 0.25], [3.0, 1.0, 0.25]


def _check(pkg, got, ref):
    rot, tr = pkg.synthetic.pose_error(got.transformation, ref.transformation)
    assert rot < ROT_TOL and tr < TRANS_TOL, (rot, tr, got.iterations, ref.iterations)
    assert abs(got.fitness - ref.fitness) < FIT_TOL
    assert abs(got.inlier_rmse - ref.inlier_rmse) < FIT_TOL


def test_multiscale_gicp_30k_l2(pkg, oracle, engine, pair30k):
    """config 1 of BASELINE.json against the reference-faithful oracle, contractive L2 kernel: far inside tolerance"""
    src, tgt, T_init, T_true = pair30k
    ref = oracle.multiscale_gicp(src, tgt, VOXELS, DISTS, 100, T_init, loss="l2")
    got = pkg.multiscale_gicp(src, tgt, VOXELS, DISTS, 100, T_init, loss="l2", engine=engine)
    _check(pkg, got, ref)
    assert got.iterations == ref.iterations
    rot, tr = pkg.synthetic.pose_error(got.transformation, ref.transformation)
    assert rot < 1e-12 and tr < 1e-11
    assert pkg.synthetic.pose_error(got.transformation, T_true)[1] < 0.25 * pkg.synthetic.pose_error(T_init, T_true)[1]


@pytest.mark.parametrize("cl", [1, 8, 48, -1, -4])      # > 0: static gang of cl blocks, < 0: task mode, -cl chunks per pass
@pytest.mark.parametrize("which", ["30k", "small"])
def test_multiscale_gicp_l1_strict(pkg, oracle, engine, pair30k, pair_small, cl, which):
    """The reference's own setting (L1 kernel, ALL_FUNCTIONS.py:284) at the north-star tolerance.

    The L1-IRLS loop is chaotic: re-associating its sums moves the result by 1e-5..1e-3 m (the oracle does that to
    itself, tests/test_oracle.py::test_l1_self_sensitivity; so does Open3D between thread counts).  The strict
    check therefore runs the oracle's ICP loop in the kernel's documented reduction order (oracle/engine_order.cpp):
    preprocessing comes from the GPU stages (proved bit-exact against the faithful oracle in test_gpu_stages.py), the
    nearest neighbours from the oracle's KD-tree."""
    from mgicp_b200 import _lib as L
    src, tgt, T_init, _ = pair30k if which == "30k" else pair_small
    opts = engine.make_opts(loss="l1", ctas_per_pair=cl)
    flat, off, _ = engine.pack_clouds([src, tgt])
    engine.preprocess_device(engine.upload(flat), off, VOXELS, opts)
    T0d = engine.upload(np.ascontiguousarray(T_init).reshape(1, 16))
    T, fit, rm, it, nc, st = (t.cpu().numpy() for t in engine.register_device([0], [1], [DISTS], [100] * 3, T0d, opts))
    engine.check()
    Tc = T_init
    for s in range(3):
        sp, sn = (engine.get_stage(0, s, w, len(src)) for w in (L.STAGE_ICP_POINTS, L.STAGE_ICP_NORMALS))
        tp, tn = (engine.get_stage(1, s, w, len(tgt)) for w in (L.STAGE_ICP_POINTS, L.STAGE_ICP_NORMALS))
        ref = oracle.gicp_engine_order(sp, sn, tp, tn, DISTS[s], Tc, 100, cl=abs(cl), loss="l1")
        Tc = ref.transformation
        assert it[0, s] == ref.iterations[0], (s, it[0], ref.iterations)
        assert abs(st[0, s, 4] - ref.fitness) < FIT_TOL and abs(st[0, s, 5] - ref.inlier_rmse) < FIT_TOL
    rot, tr = pkg.synthetic.pose_error(T[0], Tc)
    print(f"L1 {which} cl={cl}: iters {it[0].tolist()} vs engine-order oracle {rot:.2e} rad {tr:.2e} m")
    assert rot < ROT_TOL and tr < TRANS_TOL
    assert abs(fit[0] - ref.fitness) < FIT_TOL and abs(rm[0] - ref.inlier_rmse) < FIT_TOL
    assert np.abs(T[0] - Tc).max() < 1e-12                               # in practice bit-identical


def test_multiscale_gicp_30k_l1_vs_faithful_oracle(pkg, oracle, engine, pair30k):
    """L1 against the reference-faithful oracle (independent arithmetic order): inside the chaos envelope the oracle
    shows against itself when only its summation chunk changes."""
    src, tgt, T_init, T_true = pair30k
    ref = oracle.multiscale_gicp(src, tgt, VOXELS, DISTS, 100, T_init, loss="l1")
    oracle.set_sum_chunk(333)
    try:
        ref2 = oracle.multiscale_gicp(src, tgt, VOXELS, DISTS, 100, T_init, loss="l1")
    finally:
        oracle.set_sum_chunk(1024)
    got = pkg.multiscale_gicp(src, tgt, VOXELS, DISTS, 100, T_init, loss="l1", engine=engine)
    rot, tr = pkg.synthetic.pose_error(got.transformation, ref.transformation)
    rot_s, tr_s = pkg.synthetic.pose_error(ref2.transformation, ref.transformation)
    print(f"L1 30k: GPU vs faithful oracle {rot:.2e} rad {tr:.2e} m, drmse {abs(got.inlier_rmse - ref.inlier_rmse):.2e}; "
          f"oracle vs itself (re-associated sums) {rot_s:.2e} rad {tr_s:.2e} m, drmse {abs(ref2.inlier_rmse - ref.inlier_rmse):.2e}")
    assert rot < 1e-3 and tr < 2e-3 and abs(got.inlier_rmse - ref.inlier_rmse) < 5e-4 and abs(got.fitness - ref.fitness) < 5e-3
    assert pkg.synthetic.pose_error(got.transformation, T_true)[1] < 0.25 * pkg.synthetic.pose_error(T_init, T_true)[1]


def test_reference_signature_script2(pkg, oracle, engine, pair_small):
    src, tgt, T_init, _ = pair_small
    ref = oracle.Multiscale_GICP(src, tgt, 3, 30, T_init, schedule="script2", loss="l2")
    got = pkg.Multiscale_GICP(src, tgt, 3, 30, T_init, loss="l2", engine=engine)
    _check(pkg, got, ref)


def test_reference_signature_all_functions(pkg, oracle, engine, pair_small):
    src, tgt, T_init, _ = pair_small
    ref = oracle.Multiscale_GICP(src, tgt, 3, 10, T_init, schedule="all_functions", loss="l2")
    got = pkg.Multiscale_GICP(src, tgt, 3, 10, T_init, schedule="all_functions", loss="l2", engine=engine)
    _check(pkg, got, ref)


def test_float32_input_equals_float64(pkg, engine, pair_small):
    src, tgt, T_init, _ = pair_small
    a = pkg.multiscale_gicp(src, tgt, VOXELS, DISTS, 20, T_init, engine=engine)
    b = pkg.multiscale_gicp(src.astype(np.float32), tgt.astype(np.float32), VOXELS, DISTS, 20, T_init, engine=engine)
    assert np.array_equal(a.transformation, b.transformation) and a.fitness == b.fitness


def test_deterministic_and_inputs_untouched(pkg, engine, pair_small):
    src, tgt, T_init, _ = pair_small
    s0, t0, T0 = src.copy(), tgt.copy(), T_init.copy()
    a = pkg.multiscale_gicp(src, tgt, VOXELS, DISTS, 50, T_init, engine=engine)
    b = pkg.multiscale_gicp(src, tgt, VOXELS, DISTS, 50, T_init, engine=engine)
    assert np.array_equal(a.transformation, b.transformation)       # bit-identical run to run
    assert a.inlier_rmse == b.inlier_rmse and a.iterations == b.iterations
    assert np.array_equal(src, s0) and np.array_equal(tgt, t0) and np.array_equal(T_init, T0)


@pytest.mark.parametrize("ctas", [1, 2, 8, 37])
def test_cluster_sizes_agree(pkg, oracle, engine, pair_small, ctas):
    src, tgt, T_init, _ = pair_small
    ref = oracle.multiscale_gicp(src, tgt, VOXELS, DISTS, 30, T_init, loss="l2")
    got = pkg.multiscale_gicp(src, tgt, VOXELS, DISTS, 30, T_init, loss="l2", engine=engine, ctas_per_pair=ctas)
    _check(pkg, got, ref)


def test_batch_matches_single(pkg, engine):
    scans, inits, truths = pkg.synthetic.make_sequence(5, azimuth_steps=250, seed=1)
    pairs = [(i + 1, i) for i in range(4)]
    r = pkg.multiscale_gicp_batch(scans, pairs, VOXELS, DISTS, 30, np.stack(inits), engine=engine, loss="l2", ctas_per_pair=1)
    for b, (s, t) in enumerate(pairs):
        one = pkg.multiscale_gicp(scans[s], scans[t], VOXELS, DISTS, 30, inits[b], engine=engine, loss="l2", ctas_per_pair=1)
        assert np.array_equal(one.transformation, r.transformation[b])
        assert one.fitness == r.fitness[b] and one.inlier_rmse == r.inlier_rmse[b]
        assert pkg.synthetic.pose_error(r.transformation[b], truths[b])[1] < 0.05


@pytest.mark.parametrize("V", [1, 2, 4, 7])
def test_task_mode_equals_static_gangs(pkg, engine, V):
    """Dynamic scheduling must not change a single bit: a batch run in task mode with V chunks per pass (persistent
    blocks pulling (pair, chunk) tasks) equals the same batch run with a static gang of V blocks per pair; L1 kernel,
    pairs with different iteration counts, one pair without any overlap and one with an empty source."""
    scans, inits, truths = pkg.synthetic.make_sequence(7, azimuth_steps=300, seed=3)
    scans = list(scans) + [scans[0] + 500.0, np.zeros((0, 3))]
    pairs = [(i + 1, i) for i in range(6)] + [(7, 0), (8, 1), (2, 0)]
    T0 = np.stack(list(inits[:6]) + [np.eye(4), np.eye(4), inits[0]])
    its = [40, 5, 100]
    a = pkg.multiscale_gicp_batch(scans, pairs, VOXELS, DISTS, its, T0, engine=engine, loss="l1", ctas_per_pair=V)
    b = pkg.multiscale_gicp_batch(scans, pairs, VOXELS, DISTS, its, T0, engine=engine, loss="l1", ctas_per_pair=-V)
    assert np.array_equal(a.transformation, b.transformation)
    assert np.array_equal(a.fitness, b.fitness) and np.array_equal(a.inlier_rmse, b.inlier_rmse)
    assert np.array_equal(a.iterations, b.iterations) and np.array_equal(a.num_correspondences, b.num_correspondences)
    assert np.array_equal(a.stats, b.stats)
    assert len(set(map(tuple, a.iterations.tolist()))) > 3          # the pairs really do need different numbers of passes
    assert b.fitness[6] == 0.0 and b.fitness[7] == 0.0 and np.array_equal(b.transformation[7], np.eye(4))


def test_task_mode_adaptive_chunks_strict(pkg, oracle, engine):
    """Automatic mode on a batch with fewer pairs than thread blocks: task scheduling with a per-scale chunk count
    V(ns) = clamp(round(ns * blocks / (4096 * pairs)), 1, 16) (with at least one pair per block: round(ns / 8192)).  Two pairs of the batch are checked bit for bit against the oracle's ICP loop run in the
    kernel's reduction order with cl = V(ns) per scale."""
    import torch
    from mgicp_b200 import _lib as L
    scans, inits, truths = pkg.synthetic.make_sequence(25, azimuth_steps=1000, seed=4)
    pairs = [(i + 1, i) for i in range(24)]
    opts = engine.make_opts(loss="l1")
    flat, off, _ = engine.pack_clouds(scans)
    engine.preprocess_device(engine.upload(flat), off, VOXELS, opts)
    T0d = engine.upload(np.ascontiguousarray(np.stack(inits)).reshape(24, 16))
    T, fit, rm, it, nc, st = (t.cpu().numpy() for t in engine.register_device([p[0] for p in pairs], [p[1] for p in pairs],
                                                                             [DISTS] * 24, [60] * 3, T0d, opts))
    engine.check()
    blocks = torch.cuda.get_device_properties(0).multi_processor_count
    seen = set()
    for b in (0, 13):
        s_id, t_id = pairs[b]
        Tc = inits[b]
        for s in range(3):
            sp, sn = (engine.get_stage(s_id, s, w, len(scans[s_id])) for w in (L.STAGE_ICP_POINTS, L.STAGE_ICP_NORMALS))
            tp, tn = (engine.get_stage(t_id, s, w, len(scans[t_id])) for w in (L.STAGE_ICP_POINTS, L.STAGE_ICP_NORMALS))
            V = int(max(1, min(16, (len(sp) * blocks + 4096 * 24 // 2) // (4096 * 24))))
            seen.add(V)
            ref = oracle.gicp_engine_order(sp, sn, tp, tn, DISTS[s], Tc, 60, cl=V, loss="l1")
            Tc = ref.transformation
            assert it[b, s] == ref.iterations[0], (b, s, V, it[b], ref.iterations)
        assert np.abs(T[b] - Tc).max() < 1e-12 and abs(fit[b] - ref.fitness) < FIT_TOL and abs(rm[b] - ref.inlier_rmse) < FIT_TOL
    assert len(seen) > 1, seen          # the scales really ran with different chunk counts


def test_task_mode_many_pairs_auto(pkg, engine):
    """automatic mode picks task scheduling for a batch: same results as one block per pair when V resolves to 1,
    and repeatable run to run"""
    scans, inits, truths = pkg.synthetic.make_sequence(41, azimuth_steps=120, seed=9)
    pairs = [(i + 1, i) for i in range(40)] * 16                      # 640 pairs >= 4 x 148: V = 1
    T0 = np.stack(list(inits) * 16)
    a = pkg.multiscale_gicp_batch(scans, pairs, VOXELS, DISTS, 30, T0, engine=engine, loss="l1")
    b = pkg.multiscale_gicp_batch(scans, pairs, VOXELS, DISTS, 30, T0, engine=engine, loss="l1", ctas_per_pair=1)
    assert np.array_equal(a.transformation, b.transformation) and np.array_equal(a.stats, b.stats)
    assert np.array_equal(a.transformation[:40], a.transformation[40:80])


def test_nclt_fixture_against_oracle_and_golden(pkg, oracle, engine):
    """real NCLT clouds (5 mm lattice => exact ties): report-level check, loose bound"""
    import os
    g = os.path.join(os.path.dirname(__file__), "golden", "nclt")
    for i in (0, 17):
        tgt = pkg.pcd_io.read_pcd_xyz(os.path.join(g, f"s{i}.pcd"))
        src = pkg.pcd_io.read_pcd_xyz(os.path.join(g, f"s{i + 1}.pcd"))
        T0 = pkg.pcd_io.read_pose(os.path.join(g, f"fgr_pose_{i + 1}_{i}.txt"))
        G = pkg.pcd_io.read_pose(os.path.join(g, f"golden_pose_{i + 1}_{i}.txt"))
        got = pkg.Multiscale_GICP(src, tgt, 5, 100, T0, engine=engine)
        ref = oracle.Multiscale_GICP(src, tgt, 5, 100, T0)
        rot_o, tr_o = pkg.synthetic.pose_error(got.transformation, ref.transformation)
        rot_g, tr_g = pkg.synthetic.pose_error(got.transformation, G)
        print(f"NCLT {i + 1}->{i}: vs oracle {tr_o:.2e} m {rot_o:.2e} rad; vs shipped golden {tr_g:.2e} m {rot_g:.2e} rad")
        assert tr_o < 5e-3 and rot_o < 1e-3
        assert tr_g < 1.5e-2 and rot_g < 3e-3


def test_edge_cases(pkg, engine, pair_small):
    src, tgt, T_init, _ = pair_small
    with pytest.raises(RuntimeError):
        pkg.multiscale_gicp(src, tgt, [0.0], [1.0], 5, T_init, engine=engine)           # voxel_size <= 0
    with pytest.raises(RuntimeError):
        pkg.multiscale_gicp(src, tgt, [0.5], [0.0], 5, T_init, engine=engine)           # max_correspondence_distance <= 0
    # max_iteration = 0: only the initial correspondence pass, pose unchanged
    r = pkg.multiscale_gicp(src, tgt, [0.5], [1.0], 0, T_init, engine=engine)
    assert np.array_equal(r.transformation, T_init) and r.iterations == [0] and r.fitness > 0
    # no overlap at all: fitness 0, rmse 0, identity updates
    far = tgt + 1000.0
    r = pkg.multiscale_gicp(src, far, [0.5], [1.0], 5, np.eye(4), engine=engine)
    assert r.fitness == 0.0 and r.inlier_rmse == 0.0 and np.array_equal(r.transformation, np.eye(4))
    # tiny clouds (fewer points than k) and an empty source
    r = pkg.multiscale_gicp(src[:7], tgt[:9], [0.5], [1.0], 3, T_init, engine=engine)
    assert np.isfinite(r.fitness)
    r = pkg.multiscale_gicp(np.zeros((0, 3)), tgt, [0.5], [1.0], 3, T_init, engine=engine)
    assert r.fitness == 0.0 and np.array_equal(r.transformation, T_init)


# ---- evaluate_registration / get_information_matrix_from_point_clouds on the clouds as given -------
def test_evaluate_registration_and_information_matrix(pkg, oracle, engine, pair30k):
    src, tgt, T_init, T_true = pair30k
    for T, d in ((T_true, 0.25), (T_init, 0.5), (np.eye(4), 1.0)):
        ref = oracle.evaluate_registration(src, tgt, d, T, want_corr=True, want_gtg=True)
        got = engine.evaluate_clouds([src, tgt], [(0, 1)], [d], T.reshape(1, 4, 4), want_corr=True)
        assert int(got["K"][0]) == ref.num_correspondences                      # integer work: exact
        assert np.array_equal(got["corr"][0], ref.correspondence)              # same nearest neighbour for every point
        assert got["fitness"][0] == ref.fitness
        assert abs(got["rmse"][0] - ref.inlier_rmse) <= 1e-12 * max(1.0, ref.inlier_rmse)      # summation order only
        assert np.allclose(got["information"][0], ref.information, rtol=1e-12, atol=1e-9)
        r = pkg.evaluate_registration(src, tgt, d, T, engine=engine)
        assert r.fitness == ref.fitness and abs(r.inlier_rmse - ref.inlier_rmse) < 1e-12
        G = pkg.get_information_matrix_from_point_clouds(src, tgt, d, T, engine=engine)
        assert np.allclose(G, ref.information, rtol=1e-12, atol=1e-9) and np.array_equal(G, G.T)


def test_calculate_RMSE_and_fitness_circuit(pkg, oracle, engine):
    """AF:801-824: open circuit (n-1 poses) and closed circuit (n poses, the last one cloud 0 -> cloud n-1), float32 clouds"""
    scans, inits, truths = pkg.synthetic.make_sequence(5, azimuth_steps=200, seed=2)
    scans = [s.astype(np.float32) for s in scans]
    rm, fit = pkg.calculate_RMSE_and_fitness(scans, truths, 0.3, engine=engine)
    assert len(rm) == 4
    for i in range(4):
        ref = oracle.evaluate_registration(scans[i + 1], scans[i], 0.3, truths[i])
        assert fit[i] == ref.fitness and abs(rm[i] - ref.inlier_rmse) < 1e-12
    loop = np.linalg.inv(np.linalg.multi_dot(truths))                 # cloud 0 -> cloud 4
    rm, fit = pkg.calculate_RMSE_and_fitness(scans, list(truths) + [loop], 0.3, engine=engine)
    ref = oracle.evaluate_registration(scans[0], scans[4], 0.3, loop)
    assert len(rm) == 5 and fit[4] == ref.fitness and abs(rm[4] - ref.inlier_rmse) < 1e-12
    assert pkg.calculate_RMSE_and_fitness(scans, truths[:2], 0.3, engine=engine) == ([], [])


def test_evaluate_edge_cases(pkg, engine, pair_small):
    src, tgt, T_init, _ = pair_small
    with pytest.raises(RuntimeError):
        pkg.evaluate_registration(src, tgt, 0.0, T_init, engine=engine)
    r = pkg.evaluate_registration(src, tgt + 1000.0, 0.5, np.eye(4), engine=engine)
    assert r.fitness == 0.0 and r.inlier_rmse == 0.0
    assert not pkg.get_information_matrix_from_point_clouds(src, tgt + 1000.0, 0.5, np.eye(4), engine=engine).any()
    r = pkg.evaluate_registration(np.zeros((0, 3)), tgt, 0.5, np.eye(4), engine=engine)
    assert r.fitness == 0.0
    r = pkg.evaluate_registration(src, np.zeros((0, 3)), 0.5, np.eye(4), engine=engine)
    assert r.fitness == 0.0
    # a huge radius (the all_functions schedule's ~40 m): every source point matches
    r = pkg.evaluate_registration(src, tgt, 40.0, T_init, engine=engine)
    assert r.fitness == 1.0
    # after an evaluation the engine must refuse to register without a new preprocess, and a full run still works
    got = pkg.multiscale_gicp(src, tgt, VOXELS, DISTS, 5, T_init, engine=engine, loss="l2")
    assert np.isfinite(got.fitness)


# ---- BASELINE.json config 2 (the ~100k-point NCLT-shaped pair the metric is quoted on) at full size ----------------------
def test_config2_100k_pair_l1_strict_and_properties(pkg, oracle, engine):
    """~100k points per scan, voxels 1.0/0.5/0.25, the reference's L1 kernel, 100 iterations per scale, the gang the
    engine picks for a single pair (96 blocks): bit-for-bit against the engine-order oracle scale by scale, recovers the
    known motion, bit-reproducible, float32 input == float64 input, inputs untouched"""
    from mgicp_b200 import _lib as L
    src, tgt, T_init, T_true = pkg.synthetic.make_pair(3125, seed=3)
    assert 90_000 < len(src) < 110_000 and 90_000 < len(tgt) < 110_000
    src0, tgt0 = src.copy(), tgt.copy()
    a = pkg.multiscale_gicp(src, tgt, VOXELS, DISTS, 100, T_init, loss="l1", engine=engine, ctas_per_pair=96)
    # the stages of this run, then the oracle's loop in the kernel's reduction order
    Tc = T_init
    for s in range(3):
        sp, sn = (engine.get_stage(0, s, w, len(src)) for w in (L.STAGE_ICP_POINTS, L.STAGE_ICP_NORMALS))
        tp, tn = (engine.get_stage(1, s, w, len(tgt)) for w in (L.STAGE_ICP_POINTS, L.STAGE_ICP_NORMALS))
        ref = oracle.gicp_engine_order(sp, sn, tp, tn, DISTS[s], Tc, 100, cl=96, loss="l1")
        Tc = ref.transformation
        assert a.iterations[s] == ref.iterations[0], (s, a.iterations, ref.iterations)
    rot, tr = pkg.synthetic.pose_error(a.transformation, Tc)
    assert rot < ROT_TOL and tr < TRANS_TOL, (rot, tr)
    assert np.abs(a.transformation - Tc).max() < 1e-12                    # in practice bit-identical
    assert abs(a.fitness - ref.fitness) < FIT_TOL and abs(a.inlier_rmse - ref.inlier_rmse) < FIT_TOL
    b = pkg.multiscale_gicp(src, tgt, VOXELS, DISTS, 100, T_init, loss="l1", engine=engine, ctas_per_pair=96)
    assert np.array_equal(a.transformation, b.transformation) and a.fitness == b.fitness and a.inlier_rmse == b.inlier_rmse
    c = pkg.multiscale_gicp(src.astype(np.float32), tgt.astype(np.float32), VOXELS, DISTS, 100, T_init, loss="l1", engine=engine,
                            ctas_per_pair=96)
    if np.array_equal(src.astype(np.float32).astype(np.float64), src):   # the generator emits float32-representable clouds
        assert np.array_equal(a.transformation, c.transformation)
    assert np.array_equal(src, src0) and np.array_equal(tgt, tgt0)
    rot, tr = pkg.synthetic.pose_error(a.transformation, T_true)
    rot0, tr0 = pkg.synthetic.pose_error(T_init, T_true)
    print(f"100k pair: {tr0:.3f} m / {rot0:.4f} rad -> {tr:.2e} m / {rot:.2e} rad, iterations {a.iterations}, fitness {a.fitness:.3f}")
    assert tr < 0.01 and rot < 2e-3 and tr < 0.2 * tr0
    assert 0.5 < a.fitness <= 1.0 and 0.0 < a.inlier_rmse < 0.25


# ---- BASELINE.json config 4 (dense TLS-like pair, 4 scales) and config 5 (loop-closure sweep) as parity / stress cases -----
def test_config4_tls_pair_vs_oracle(pkg, oracle, engine):
    """Courtyard/Facade-shaped dense pair (300k points per cloud: 40k ... 190k points per scale after down-sampling),
    script-2 4-scale schedule (voxels 0.4/0.3/0.2/0.1, distances 1.2/0.75/0.4/0.1) against the faithful oracle"""
    src, tgt, T_init, T_true = pkg.synthetic.make_tls_pair(300_000, seed=1)
    ref = oracle.Multiscale_GICP(src, tgt, 4, 100, T_init, schedule="script2", loss="l2")
    got = pkg.Multiscale_GICP(src, tgt, 4, 100, T_init, loss="l2", engine=engine)
    _check(pkg, got, ref)
    assert got.iterations == ref.iterations
    rot, tr = pkg.synthetic.pose_error(got.transformation, T_true)
    assert rot < 2e-4 and tr < 2e-3
    # the finest scale's down-sampled cloud (the large-cloud hash stress): same point set, bit for bit
    from mgicp_b200 import _lib as L
    ds = engine.get_stage(0, 3, L.STAGE_DOWNSAMPLED, len(src))
    ref_ds = oracle.voxel_down_sample(src, pkg.create_scales_script2(4)[3])
    key = lambda a: a[np.lexsort((a[:, 2], a[:, 1], a[:, 0]))]
    assert np.array_equal(key(ds), key(np.asarray(ref_ds)))


def test_config4_tls_pair_2M_points(pkg, engine):
    """the full-size case: ~2M points per cloud, 4 scales, the reference's L1 kernel; checked through size-independent
    properties (recovers the known motion, bit-reproducible, sane statistics)"""
    src, tgt, T_init, T_true = pkg.synthetic.make_tls_pair(2_000_000, seed=2)
    src32, tgt32 = src.astype(np.float32), tgt.astype(np.float32)
    a = pkg.Multiscale_GICP(src32, tgt32, 4, 100, T_init, engine=engine)
    b = pkg.Multiscale_GICP(src32, tgt32, 4, 100, T_init, engine=engine)
    assert np.array_equal(a.transformation, b.transformation) and a.inlier_rmse == b.inlier_rmse
    rot, tr = pkg.synthetic.pose_error(a.transformation, T_true)
    rot0, tr0 = pkg.synthetic.pose_error(T_init, T_true)
    print(f"2M TLS pair: {tr0:.3f} m / {rot0:.4f} rad -> {tr:.2e} m / {rot:.2e} rad, iterations {a.iterations}, fitness {a.fitness:.3f}, "
          f"points per scale {a.stats[:, 0].astype(int).tolist()}")
    assert tr < 2e-3 and rot < 2e-4 and 0.3 < a.fitness <= 1.0 and a.inlier_rmse < 0.1
    assert a.stats[3, 0] > 400_000                                   # the finest scale really is a large cloud
    ev = pkg.evaluate_registration(src32, tgt32, 0.1, a.transformation, engine=engine)      # raw 2M x 2M evaluation
    ev0 = pkg.evaluate_registration(src32, tgt32, 0.1, T_init, engine=engine)
    assert ev.fitness > ev0.fitness and ev.inlier_rmse < ev0.inlier_rmse


def test_config5_loop_closure_sweep(pkg, oracle, engine):
    """non-consecutive pairs with large initial offsets (identity as the initial guess for scans 10+ apart): low fitness,
    iteration caps, big search boxes; the batch (task mode) must equal pair-at-a-time runs under the contractive L2
    kernel, and spot checks against the oracle hold"""
    n = 36
    scans, inits, truths = pkg.synthetic.make_sequence(n, azimuth_steps=200, seed=6)
    pairs = [(i, j) for i in range(n) for j in range(n) if i - j >= 10][:220]
    T0 = np.stack([np.eye(4)] * len(pairs))
    vox, dist, its = [1.0, 0.5, 0.25], [3.0, 1.0, 0.25], [30, 30, 30]
    r = pkg.multiscale_gicp_batch(scans, pairs, vox, dist, its, T0, engine=engine, loss="l2")
    assert np.isfinite(r.transformation).all() and (r.fitness >= 0).all() and (r.fitness <= 1).all()
    assert (r.iterations == 30).any() and (r.fitness < 0.4).any()           # caps are hit, many pairs fail (as expected)
    for b in (0, 57, 219):
        s, t = pairs[b]
        one = pkg.multiscale_gicp(scans[s], scans[t], vox, dist, its, np.eye(4), engine=engine, loss="l2", ctas_per_pair=1)
        ref = oracle.multiscale_gicp(scans[s], scans[t], vox, dist, its, np.eye(4), loss="l2")
        assert one.iterations == r.iterations[b].tolist() == ref.iterations
        for got in (one.transformation, r.transformation[b]):
            rot, tr = pkg.synthetic.pose_error(got, ref.transformation)
            assert rot < 1e-7 and tr < 1e-7, (b, rot, tr)
        assert abs(r.fitness[b] - ref.fitness) < FIT_TOL and abs(r.inlier_rmse[b] - ref.inlier_rmse) < FIT_TOL


def test_correspondence_set_of_the_result(pkg, oracle, engine, pair_small):
    """RegistrationResult.correspondence_set (read at ALL_FUNCTIONS.py:1064): K rows, consistent with fitness / inlier_rmse, and
    every row is the nearest target point of its source point at the returned pose"""
    from mgicp_b200 import _lib as L
    src, tgt, T_init, _ = pair_small
    r = pkg.multiscale_gicp(src, tgt, VOXELS, DISTS, 30, T_init, loss="l2", engine=engine)
    cs = r.correspondence_set
    sp = engine.get_stage(0, 2, L.STAGE_POINTS, len(src))
    tp = engine.get_stage(1, 2, L.STAGE_POINTS, len(tgt))
    assert cs.shape == (r.num_correspondences, 2) and cs.dtype == np.int32
    assert abs(len(cs) / len(sp) - r.fitness) < 1e-15
    assert np.all(np.diff(cs[:, 0]) > 0) and cs[:, 1].max() < len(tp)
    # the pose of the last correspondence pass is the one BEFORE the last update was applied only if the loop ended on the
    # iteration cap; this run converges, so the returned pose is the pose the correspondences were found at
    moved = sp[cs[:, 0]] @ r.transformation[:3, :3].T + r.transformation[:3, 3]
    d = np.linalg.norm(moved - tp[cs[:, 1]], axis=1)
    assert d.max() < DISTS[2]
    assert abs(np.sqrt(np.mean(d ** 2)) - r.inlier_rmse) < 1e-9
    from scipy.spatial import cKDTree
    dn, jn = cKDTree(tp).query(moved)
    assert np.allclose(dn, d, rtol=0, atol=1e-12)


@pytest.mark.parametrize("loss,k", [("huber", 0.05), ("cauchy", 0.1), ("gm", 0.1), ("tukey", 0.5)])
def test_other_robust_kernels(pkg, oracle, engine, pair_small, loss, k):
    """Open3D's other RobustKernels (HuberLoss, CauchyLoss, GMLoss, TukeyLoss; the reference uses L1Loss, AF:284): weights as in
    RobustKernel.cpp on both sides; single-pass normal equations to rounding, the loop within the north-star tolerances"""
    src, tgt, T_init, _ = pair_small
    ref = oracle.multiscale_gicp(src, tgt, VOXELS, DISTS, 30, T_init, loss=loss, loss_k=k)
    got = pkg.multiscale_gicp(src, tgt, VOXELS, DISTS, 30, T_init, loss=loss, loss_k=k, engine=engine)
    rot, tr = pkg.synthetic.pose_error(got.transformation, ref.transformation)
    print(f"{loss}(k={k}): vs oracle {rot:.2e} rad {tr:.2e} m, iterations {got.iterations} / {ref.iterations}, "
          f"dfitness {abs(got.fitness - ref.fitness):.1e} drmse {abs(got.inlier_rmse - ref.inlier_rmse):.1e}")
    _check(pkg, got, ref)


def test_warp_specialised_task_kernel(pkg, engine):
    """k_icp_tasks_ws (MGICP_WS=1: 16 search warps feed 8 linearise warps through shared-memory rings, setmaxnreg): the same
    per-virtual-thread columns in the same order, hence the results of k_icp_tasks bit for bit -- on a batch with mixed
    iteration counts, an empty source and a pair without overlap, fixed and adaptive chunking."""
    import os
    pairs, clouds, T0 = [], [], []
    for k in range(6):
        s, t, Ti, _ = pkg.synthetic.make_pair(300 + 40 * k, seed=40 + k)
        clouds += [s, t]
        pairs.append((2 * k, 2 * k + 1))
        T0.append(Ti)
    clouds.append(np.zeros((0, 3)))
    pairs.append((len(clouds) - 1, 1)); T0.append(np.eye(4))                       # empty source
    far = clouds[0] + np.array([500.0, 0.0, 0.0])
    clouds.append(far)
    pairs.append((len(clouds) - 1, 1)); T0.append(np.eye(4))                       # no overlap
    T0 = np.stack(T0)
    for cpp in (-1, -3, 0):
        opts = engine.make_opts(loss="l1", ctas_per_pair=cpp)
        res = {}
        for ws in ("0", "1"):
            os.environ["MGICP_WS"] = ws
            try:
                res[ws] = engine.run(clouds, pairs * (25 if cpp == 0 else 1), VOXELS, DISTS, 60, np.concatenate([T0] * (25 if cpp == 0 else 1)), opts)
            finally:
                os.environ.pop("MGICP_WS", None)
        a, b = res["0"], res["1"]
        assert np.array_equal(a.transformation, b.transformation) and np.array_equal(a.iterations, b.iterations)
        assert np.array_equal(a.fitness, b.fitness) and np.array_equal(a.inlier_rmse, b.inlier_rmse) and np.array_equal(a.stats, b.stats)
