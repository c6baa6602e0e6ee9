"""The shared small-matrix header (csrc/cm_math.h) must mean the same thing on the GPU and on the host, bit for bit."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _inputs(op, n, rng):
    if op in (0, 3, 4):
        B = rng.normal(size=(n, 12, 6)).astype(np.float32) * np.array([1, 1, 1, 30, 30, 30], np.float32)
        B[::7, :, 0] *= 1e-3                      # near-degenerate direction
        A = np.einsum("nkr,nkc->nrc", B, B).astype(np.float32).reshape(n, 36)
        if op == 0:
            return np.concatenate([A, rng.normal(size=(n, 6)).astype(np.float32) * 10], 1)
        return A
    if op == 1:
        nrm = rng.normal(size=(n, 3)); nrm[:, 2] += 2.0
        xy = rng.uniform(-20, 20, (n, 5, 2))
        z = (3.0 - nrm[:, None, 0] * xy[..., 0] - nrm[:, None, 1] * xy[..., 1]) / nrm[:, None, 2] + rng.normal(0, 0.05, (n, 5))
        A = np.concatenate([xy, z[..., None]], -1).astype(np.float32).reshape(n, 15)
        return np.concatenate([A, -np.ones((n, 5), np.float32)], 1)
    if op == 2:
        B = rng.normal(size=(n, 6, 3)).astype(np.float32) * rng.choice([1e-2, 1.0, 5.0], (n, 1, 1)).astype(np.float32)
        B[::5, :, 1] = 2 * B[::5, :, 0] + 1e-4 * B[::5, :, 1]     # near rank-1 (a line)
        C = np.einsum("nkr,nkc->nrc", B, B).astype(np.float32)
        return np.stack([C[:, 0, 0], C[:, 1, 0], C[:, 2, 0], C[:, 1, 1], C[:, 2, 1], C[:, 2, 2]], 1)
    if op == 5:
        return (rng.normal(size=(n, 36)) + 2 * np.eye(6).ravel()).astype(np.float32)
    if op == 6:
        return (rng.uniform(-1, 1, (n, 6)) * np.array([3.2, 1.6, 3.2, 100, 100, 100])).astype(np.float32)
    if op == 7:
        x = (rng.uniform(-1, 1, (n, 2)) * rng.choice([1e-3, 1.0, 120.0], (n, 2))).astype(np.float32)
        x[::97, 0] = 0.0; x[::89, 1] = 0.0; x[::101] = np.abs(x[::101])
        x[(x[:, 0] == 0) & (x[:, 1] == 0), 1] = 1.0        # 0 / 0: the NaN's sign bit is platform-defined
        return x
    raise ValueError(op)


@pytest.mark.parametrize("op", range(8))
def test_math_device_equals_host(ctx, oracle, op):
    x = _inputs(op, 4096, np.random.default_rng(100 + op))
    dev = ctx.debug_math(op, x)
    host = oracle.debug_math(op, x)
    assert np.array_equal(dev.view(np.uint32), host.view(np.uint32))


def test_warp_qr_equals_sequential_qr(ctx, oracle):
    """warp_qr_solve6 (the column-parallel 6x6 column-pivoted Householder QR of the Gauss-Newton step, one warp per system) against
    cm_math.h::colpiv_qr_solve<6, 6> on the host: bit for bit, for well-conditioned normal equations and for rank-deficient ones
    (zero columns, repeated columns, zero matrix, tiny pivots) where the pivot count stops short of 6."""
    rng = np.random.default_rng(77)
    x = _inputs(0, 6000, rng)
    A = x[:, :36].reshape(-1, 6, 6).copy()
    A[::11, :, 2] = 0.0; A[::11, 2, :] = 0.0                      # a zero column (and row)
    A[::13, :, 4] = A[::13, :, 1]; A[::13, 4, :] = A[::13, 1, :]  # two equal columns
    A[::17] = 0.0                                                 # nothing at all
    A[::19, :, 5] *= 1e-9; A[::19, 5, :] *= 1e-9                  # a pivot below the rank threshold
    A[5::23] = rng.normal(size=A[5::23].shape).astype(np.float32)  # general (unsymmetric) systems
    x[:, :36] = A.reshape(-1, 36)
    dev = ctx.debug_math(8, x)
    host = oracle.debug_math(0, x)
    # (a zero matrix divides 0 by 0 in the back substitution, as Eigen does: NaNs in the same places -- their sign bit is the platform's)
    nan = np.isnan(host)
    assert np.array_equal(np.isnan(dev), nan)
    bad = np.where((dev.view(np.uint32) != host.view(np.uint32)) & ~nan)[0]
    assert len(bad) == 0, (len(bad), sorted(set(bad.tolist()))[:20], dev[bad[:2]], host[bad[:2]])
    assert np.isfinite(host).mean() > 0.9


def test_eigen_solve_skip_never_hides_a_degenerate_system(ctx, oracle):
    """dev_min_eig_above6 (evaluation 0 skips SelfAdjointEigenSolver when A^T A is provably well-conditioned): whenever it says
    "skip", the restated Eigen solver's smallest eigenvalue IS at or above the threshold -- over matrices whose smallest eigenvalue
    sweeps through the threshold (100, and the odometry's 10), near-singular ones, and non-finite ones; and it does skip the
    clearly well-conditioned ones."""
    rng = np.random.default_rng(5)
    n = 20000
    Q, _ = np.linalg.qr(rng.normal(size=(n, 6, 6)))
    lam = np.sort(np.exp(rng.uniform(np.log(1e2), np.log(1e7), (n, 6))), axis=1)
    thr = np.where(np.arange(n) % 2 == 0, 100.0, 10.0)
    lam[:, 0] = thr * np.exp(rng.normal(0, 0.02, n))                 # smallest eigenvalue within a few per cent of the threshold
    lam[::5, 0] = thr[::5] * rng.uniform(2.0, 50.0, len(thr[::5]))   # clearly above
    lam[1::5, 0] = thr[1::5] * rng.uniform(0.0, 0.5, len(thr[1::5])) # clearly below
    A = np.einsum("nij,nj,nkj->nik", Q, lam, Q).astype(np.float32)
    A = 0.5 * (A + A.transpose(0, 2, 1))
    A[7::101, 0, 0] = np.nan; A[9::103] = 0.0; A[11::107, 3, 3] = np.inf
    x = np.concatenate([A.reshape(n, 36), thr[:, None].astype(np.float32)], 1)
    skip = ctx.debug_math(9, x)[:, 0] > 0
    E0 = oracle.debug_math(4, A.reshape(n, 36))[:, 0]
    assert not np.any(skip & ~(E0 >= thr)), "skipped a system the eigen-solver calls degenerate (or non-finite)"
    # its margin is 1e-4 trace(A): a system whose smallest eigenvalue clears the threshold by twice that is skipped
    fin = np.isfinite(A.reshape(n, 36)).all(1) & (np.abs(A.reshape(n, 36)).sum(1) > 0)
    clear = fin & (E0 > thr + 2e-4 * np.trace(A, axis1=1, axis2=2))
    assert clear.sum() > 500 and skip[clear].all()
    assert 0.02 < skip.mean() < 0.8
