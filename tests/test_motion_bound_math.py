"""The motion-box padding of the TLAS (kiraray_b200/csrc/bvh_build.cu chainMotionBound): a point of a moving instance is
at most V * h / 2 away from its position at the nearest of the sample times (spacing h), V being the speed bound computed
from the SRT keys.  This test restates the bound in float64 numpy, level by level as the kernel does, and checks the
property it rests on -- |p(t + dt) - p(t)| <= V dt for every point, time and small dt -- on random two-level chains,
including fast rotations, near-opposite key quaternions (the normalised blend then swings through a large angle in a
short time), multi-key nodes and non-uniform scales."""
import numpy as np

from kiraray_b200 import scenes


def level_matrix(keys, t0, t1, time):
    """T * R(q) * S of an SRT node at `time` (motion.cuh srtNodeXf semantics) as a 4x4, float64"""
    v = scenes.srt_lerp(keys, time, t0, t1)
    m = np.eye(4)
    m[:3, :3] = scenes.quat_to_mat(v[3:7]) * v[0:3][None, :]
    m[:3, 3] = v[7:10]
    return m


def level_bound(keys, t0, t1, B, V):
    """one level of chainMotionBound: (B, V) below the node -> (B, V) above it"""
    keys = np.asarray(keys, np.float64)
    n = len(keys)
    sigma = np.abs(keys[:, 0:3]).max()
    tmax = np.linalg.norm(keys[:, 7:10], axis=1).max()
    A = C = 0.0
    for k in range(n - 1):
        a, b = keys[k], keys[k + 1]
        ds = np.abs(b[0:3] - a[0:3]).max()
        dT = np.linalg.norm(b[7:10] - a[7:10])
        d = b[3:7] - a[3:7]
        dd, aa, ad, bb = d @ d, a[3:7] @ a[3:7], a[3:7] @ d, b[3:7] @ b[3:7]
        q2 = min(aa, bb)
        if dd > 0:
            f = -ad / dd
            if 0 < f < 1:
                q2 = min(q2, max(aa - ad * ad / dd, 0.0))
        omega = 2 * np.sqrt(dd) / max(np.sqrt(q2), 1e-30) if dd > 0 else 0.0
        A, C = max(A, omega * sigma + ds), max(C, dT)
    fp = (n - 1) / (t1 - t0)
    return sigma * B + tmax, fp * (A * B + C) + sigma * V


def random_keys(rng, n, spin, drift, scale_jitter, opposite=False):
    q0 = rng.normal(size=4)
    q0 /= np.linalg.norm(q0)
    keys = []
    for k in range(n):
        q = q0 + rng.normal(size=4) * spin * k
        if opposite and k == n - 1:
            q = -q0 + rng.normal(size=4) * 0.05  # nearly antipodal to the first key
        q /= np.linalg.norm(q)
        s = rng.uniform(0.5, 1.5) * (1 + scale_jitter * rng.normal(size=3))
        keys.append(np.concatenate([s, q, rng.normal(size=3) * drift * (k + 1)]))
    return np.array(keys)


def test_speed_bound_dominates_the_trajectory():
    rng = np.random.Generator(np.random.PCG64(20251018))
    worst = 0.0
    for case in range(60):
        n0, n1 = rng.integers(2, 5), rng.integers(2, 5)
        leaf = random_keys(rng, n0, spin=rng.choice([0.1, 1.0, 5.0]), drift=rng.choice([0.05, 1.0]), scale_jitter=0.2, opposite=case % 7 == 0)
        parent = random_keys(rng, n1, spin=rng.choice([0.1, 2.0]), drift=rng.choice([0.1, 3.0]), scale_jitter=0.1, opposite=case % 11 == 0)
        t0, t1 = 0.0, float(rng.uniform(0.5, 2.0))
        pts = rng.normal(size=(6, 3)) * 1.3
        B0 = np.linalg.norm(pts, axis=1).max()
        B1, V1 = level_bound(leaf, t0, t1, B0, 0.0)
        _, V = level_bound(parent, t0, t1, B1, V1)
        ts = np.sort(rng.uniform(t0 - 0.1, t1 + 0.1, 400))
        dt = 1e-4 * (t1 - t0)
        for t in ts:
            m_a = level_matrix(parent, t0, t1, t) @ level_matrix(leaf, t0, t1, t)
            m_b = level_matrix(parent, t0, t1, t + dt) @ level_matrix(leaf, t0, t1, t + dt)
            for p in pts:
                step = np.linalg.norm((m_b - m_a) @ np.append(p, 1.0))
                worst = max(worst, step / (V * dt))
                assert step <= V * dt * (1 + 1e-6) + 1e-12, (case, t, step, V * dt)
    # the bound is not vacuous either: some trajectory comes within an order of magnitude of it
    assert worst > 0.1, worst
    print(f"speed bound: largest |dp| / (V dt) over all cases {worst:.3f}")


def test_sample_hull_plus_pad_contains_every_time_of_the_window():
    """the statement the TLAS relies on: the positions at kMotionSamples + 1 sample times, padded by V h / 2, contain the
    position at EVERY time of the shutter window"""
    rng = np.random.Generator(np.random.PCG64(7))
    for case in range(20):
        leaf = random_keys(rng, 2, spin=4.0, drift=0.5, scale_jitter=0.0, opposite=case % 5 == 0)
        parent = random_keys(rng, 3, spin=1.0, drift=2.0, scale_jitter=0.0)
        t0, t1 = 0.0, 1.0
        w0, w1 = sorted(rng.uniform(0, 1, 2))
        p = rng.normal(size=3)
        B1, V1 = level_bound(leaf, t0, t1, np.linalg.norm(p), 0.0)
        _, V = level_bound(parent, t0, t1, B1, V1)
        steps = 8
        pos = lambda t: (level_matrix(parent, t0, t1, t) @ level_matrix(leaf, t0, t1, t) @ np.append(p, 1.0))[:3]
        samples = np.array([pos(w0 + (w1 - w0) * j / steps) for j in range(steps + 1)])
        pad = V * 0.5 * (w1 - w0) / steps
        lo, hi = samples.min(0) - pad, samples.max(0) + pad
        for t in rng.uniform(w0, w1, 300):
            q = pos(t)
            assert (q >= lo - 1e-9).all() and (q <= hi + 1e-9).all(), (case, t)
