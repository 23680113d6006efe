"""Ensemble parity under the reference's own setting (L1 kernel, ALL_FUNCTIONS.py:284): a distribution, not one pair.

The L1-IRLS loop is chaotic: re-associating the 27 sums of the CPU oracle moves ITS OWN result by
1e-6..1e-3 m.  So the GPU engine (another summation order) is held to this: over an ensemble of tie-free synthetic pairs,
its distance to the faithful oracle must be distributed like the oracle's distance to itself under re-association -- and the
fraction of pairs inside the north-star tolerances (1e-4 rad, 1e-4 m, 1e-5 fitness / RMSE) is printed for both.
Second leg: the initial poses rounded to 10 decimals, the way the reference reads them from its %.10f text files
(2_MGICP_refinement_in_NCLT_dataset.py:173): Open3D (and the oracle) carry the non-orthonormal rotation into the
covariances, the engine assumes an orthonormal one (known deviation)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

VOXELS, DISTS = [1.0, 0.5, 0.25], [3.0, 1.0, 0.25]
N_PAIRS = 32


def _deltas(pkg, A, B):
    """A, B: lists of results with .transformation / .fitness / .inlier_rmse"""
    rot, tr = np.array([pkg.synthetic.pose_error(a.transformation, b.transformation) for a, b in zip(A, B)]).T
    df = np.array([abs(a.fitness - b.fitness) for a, b in zip(A, B)])
    dr = np.array([abs(a.inlier_rmse - b.inlier_rmse) for a, b in zip(A, B)])
    inside = (rot < 1e-4) & (tr < 1e-4) & (df < 1e-5) & (dr < 1e-5)
    return dict(rot=rot, tr=tr, dfit=df, drmse=dr, inside=float(inside.mean()))


def _fmt(d):
    q = lambda x: f"median {np.median(x):.2e} p90 {np.quantile(x, 0.9):.2e} max {x.max():.2e}"
    return (f"rot [{q(d['rot'])}] rad; trans [{q(d['tr'])}] m; dfitness [{q(d['dfit'])}]; drmse [{q(d['drmse'])}]; "
            f"inside 1e-4/1e-4/1e-5: {100 * d['inside']:.0f} %")


def _assert_like_envelope(gpu, env):
    for key, floor in (("rot", 1e-7), ("tr", 1e-6), ("dfit", 1e-6), ("drmse", 1e-6)):
        assert np.median(gpu[key]) <= 3.0 * np.median(env[key]) + floor, key
        assert np.quantile(gpu[key], 0.9) <= 3.0 * np.quantile(env[key], 0.9) + 10 * floor, key
        assert gpu[key].max() <= 10.0 * env[key].max() + 100 * floor, key
    assert gpu["inside"] >= env["inside"] - 0.25
    # absolute ceilings: nothing anywhere near a wrong basin
    assert gpu["rot"].max() < 2e-4 and gpu["tr"].max() < 1e-3 and gpu["drmse"].max() < 1e-3 and gpu["dfit"].max() < 1e-3


class _R:
    def __init__(self, T, f, r):
        self.transformation, self.fitness, self.inlier_rmse = T, f, r


@pytest.fixture(scope="module")
def ensemble(pkg):
    return [pkg.synthetic.make_pair(1000, seed=100 + k) for k in range(N_PAIRS)]


def _oracle_all(oracle, ens, inits, chunk=1024):
    oracle.set_sum_chunk(chunk)
    try:
        return [oracle.multiscale_gicp(s, t, VOXELS, DISTS, 100, T0, loss="l1") for (s, t, _, _), T0 in zip(ens, inits)]
    finally:
        oracle.set_sum_chunk(1024)


def _gpu_all(pkg, engine, ens, inits):
    clouds, pairs = [], []
    for k, (s, t, _, _) in enumerate(ens):
        clouds += [s, t]
        pairs.append((2 * k, 2 * k + 1))
    r = pkg.multiscale_gicp_batch(clouds, pairs, VOXELS, DISTS, 100, np.stack(inits), engine=engine, loss="l1")
    return [_R(r.transformation[b], r.fitness[b], r.inlier_rmse[b]) for b in range(len(ens))]


def test_l1_ensemble_against_the_oracles_own_envelope(pkg, oracle, engine, ensemble):
    inits = [e[2] for e in ensemble]
    ref = _oracle_all(oracle, ensemble, inits)
    ref2 = _oracle_all(oracle, ensemble, inits, chunk=333)          # the oracle against itself, sums re-associated
    got = _gpu_all(pkg, engine, ensemble, inits)
    env, gpu = _deltas(pkg, ref2, ref), _deltas(pkg, got, ref)
    print(f"\nL1 ensemble, {N_PAIRS} synthetic 30k pairs, 3 scales\n  oracle vs itself (chunk 333 vs 1024): {_fmt(env)}\n"
          f"  GPU vs oracle:                        {_fmt(gpu)}")
    truth = np.array([pkg.synthetic.pose_error(g.transformation, e[3])[1] for g, e in zip(got, ensemble)])
    truth_o = np.array([pkg.synthetic.pose_error(g.transformation, e[3])[1] for g, e in zip(ref, ensemble)])
    print(f"  distance to the true motion: GPU median {np.median(truth):.2e} m, oracle median {np.median(truth_o):.2e} m")
    # the GPU's scatter around the oracle is the oracle's own scatter: medians and the 90 % quantiles within a small factor; the
    # maxima of 32 samples of a heavy-tailed (chaotic) quantity only within an order of magnitude.  First B200 run (round 2):
    # medians 9.7e-7 rad / 9.3e-6 m (oracle vs itself 5.2e-7 / 1.0e-5), p90 3.9e-6 / 2.1e-5 (4.2e-6 / 2.4e-5), 47 % of the pairs
    # inside all four north-star tolerances (oracle vs itself: 62 %; binomial noise at n = 32 is +-9 %)
    _assert_like_envelope(gpu, env)
    assert abs(np.median(truth) - np.median(truth_o)) < 1e-4


def test_l1_ensemble_with_percent_10f_initial_poses(pkg, oracle, engine, ensemble):
    """T_init as the reference reads it: np.savetxt(fmt='%.10f') text (S1:176-177 -> S2:173), rotation orthonormal to ~1e-10"""
    inits = [np.array([[float(f"{v:.10f}") for v in row] for row in e[2]]) for e in ensemble]
    dev = max(np.abs(T[:3, :3] @ T[:3, :3].T - np.eye(3)).max() for T in inits)
    assert 1e-12 < dev < 1e-9                                   # genuinely non-orthonormal, as in the reference's files
    ref = _oracle_all(oracle, ensemble, inits)
    ref2 = _oracle_all(oracle, ensemble, inits, chunk=333)
    got = _gpu_all(pkg, engine, ensemble, inits)
    env, gpu = _deltas(pkg, ref2, ref), _deltas(pkg, got, ref)
    print(f"\nL1 ensemble with %.10f-rounded T_init (|R R^T - I| up to {dev:.1e})\n  oracle vs itself: {_fmt(env)}\n  GPU vs oracle:    {_fmt(gpu)}")
    _assert_like_envelope(gpu, env)


def test_l2_ensemble_is_far_inside_tolerance(pkg, oracle, engine, ensemble):
    """the contractive kernel on the same ensemble, %.10f initial poses included: no chaos to hide behind"""
    sub = ensemble[:8]
    inits = [np.array([[float(f"{v:.10f}") for v in row] for row in e[2]]) for e in sub]
    ref = [oracle.multiscale_gicp(s, t, VOXELS, DISTS, 100, T0, loss="l2") for (s, t, _, _), T0 in zip(sub, inits)]
    clouds, pairs = [], []
    for k, (s, t, _, _) in enumerate(sub):
        clouds += [s, t]
        pairs.append((2 * k, 2 * k + 1))
    r = pkg.multiscale_gicp_batch(clouds, pairs, VOXELS, DISTS, 100, np.stack(inits), engine=engine, loss="l2")
    got = [_R(r.transformation[b], r.fitness[b], r.inlier_rmse[b]) for b in range(len(sub))]
    d = _deltas(pkg, got, ref)
    print(f"\nL2 ensemble with %.10f-rounded T_init: {_fmt(d)}")
    assert d["rot"].max() < 1e-8 and d["tr"].max() < 1e-7 and d["dfit"].max() < 1e-9 and d["drmse"].max() < 1e-9
