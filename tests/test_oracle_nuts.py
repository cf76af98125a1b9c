"""CPU checks of the NUTS oracle (oracle/nuts_oracle.py) and of the host-side schedule of the product (mile_b200/nuts.py).
The reference holds no vectors for this branch and blackjax is not installable here (parity unpinned): the restatement
is pinned by the invariants the algorithm must satisfy."""
import numpy as np

from oracle import nuts_oracle as no


def gaussian(sig):
    return lambda th: (-0.5 * np.sum((th / sig) ** 2), -th / sig ** 2)


def test_velocity_verlet_is_reversible_and_second_order():
    sig = np.array([0.5, 1.0, 2.0])
    lg = gaussian(sig)
    rng = np.random.default_rng(0)
    th, p = rng.standard_normal(3), rng.standard_normal(3)
    imm = np.array([0.3, 1.0, 2.5])
    _, g = lg(th)
    t1, p1, l1, g1 = no.velocity_verlet(lg, th, p, g, 0.1, imm)
    t0, p0, _, _ = no.velocity_verlet(lg, t1, -p1, g1, 0.1, imm)
    np.testing.assert_allclose(t0, th, atol=1e-13)
    np.testing.assert_allclose(-p0, p, atol=1e-13)
    errs = []
    for eps in (0.1, 0.05):
        t, q, gg = th, p, g
        for _ in range(int(round(1.0 / eps))):
            t, q, l, gg = no.velocity_verlet(lg, t, q, gg, eps, imm)
        h0 = -lg(th)[0] + no.kinetic_energy(p, imm)
        errs.append(abs(-l + no.kinetic_energy(q, imm) - h0))
    assert 3.0 < errs[0] / errs[1] < 5.0          # energy error O(eps^2)


def test_checkpoint_indices_match_the_published_examples():
    # termination._leaf_idx_to_ckpt_idxs comments: set bits of n>>1: 6 -> 2, 7 -> 2, 13 -> 2; trailing ones: 6 -> 0, 7 -> 3, 13 -> 1
    assert no.leaf_idx_to_ckpt_idxs(6) == (2 - 0 + 1, 2)
    assert no.leaf_idx_to_ckpt_idxs(7) == (2 - 3 + 1, 2)
    assert no.leaf_idx_to_ckpt_idxs(13) == (2 - 1 + 1, 2)
    assert no.leaf_idx_to_ckpt_idxs(0) == (1, 0) and no.leaf_idx_to_ckpt_idxs(1) == (0, 0)
    # every odd leaf n closes sub-trees whose first leaves are n - 2^k + 1 (k = 1 .. trailing ones): the checkpoints read for
    # them must be the ones those (even) leaves wrote
    written = {}
    for n in range(64):
        lo, hi = no.leaf_idx_to_ckpt_idxs(n)
        if n % 2 == 0:
            written[hi] = n
        else:
            firsts = sorted(written[i] for i in range(lo, hi + 1))
            t, k = n, 0
            while t & 1:
                t >>= 1
                k += 1
            assert firsts == sorted(n - 2 ** j + 1 for j in range(1, k + 1))


def test_nuts_samples_a_gaussian_and_adapts_metric_and_step_size():
    sig = np.array([0.5, 1.0, 2.0, 3.0, 0.1])
    lg = gaussian(sig)
    rng = np.random.default_rng(0)
    D, d, n = 8, 5, 1000
    th, lp, g, eps, imm, infos = no.run_window_adaptation(lg, np.ones(d), rng.standard_normal((n, d)),
                                                          rng.random((n, no.uni_len(D))), max_num_doublings=D)
    np.testing.assert_allclose(imm, sig ** 2, rtol=0.25)
    assert abs(np.mean([i.acceptance_rate for i in infos[-250:]]) - 0.8) < 0.08
    n = 3000
    th, lp, g, pos, infos = no.run_nuts(lg, th, lp, g, eps, imm, rng.standard_normal((n, d)), rng.random((n, no.uni_len(D))), D)
    assert not any(i.is_divergent for i in infos)
    assert all(1 <= i.num_integration_steps <= 2 ** D - 1 + 2 ** (D - 1) for i in infos)
    np.testing.assert_allclose(pos.mean(0) / sig, 0, atol=0.1)
    np.testing.assert_allclose(pos.std(0) / sig, 1, atol=0.08)
    # fp32 twin follows the fp64 one on the first transitions
    a = no.nuts_step(lg, np.ones(d), *lg(np.ones(d)), 0.3, np.ones(d), np.full(d, 0.7), np.full(no.uni_len(4), 0.4), 4)
    b = no.nuts_step(lambda t: tuple(np.float32(v) for v in lg(t.astype(np.float64))), np.ones(d, np.float32),
                     np.float32(lg(np.ones(d))[0]), lg(np.ones(d))[1].astype(np.float32), 0.3, np.ones(d, np.float32),
                     np.full(d, 0.7, np.float32), np.full(no.uni_len(4), 0.4), 4)
    assert a[3][0] == b[3][0] and a[3][2] == b[3][2]
    np.testing.assert_allclose(a[0], b[0], rtol=1e-5)


def test_divergence_stops_the_trajectory_and_keeps_the_state():
    lg = gaussian(np.array([1e-3, 1e-3]))
    th = np.array([1.0, 1.0])
    lp, g = lg(th)
    t1, l1, g1, info = no.nuts_step(lg, th, lp, g, 5.0, np.ones(2), np.array([0.3, -0.2]), np.full(no.uni_len(5), 0.5), 5)
    assert info.is_divergent and info.num_trajectory_expansions == 1 and info.acceptance_rate < 1e-6
    np.testing.assert_array_equal(t1, th)


def test_schedule_structure_and_product_mirror():
    from mile_b200.nuts import build_schedule
    for n in (10, 19, 20, 100, 150, 333, 1000, 2000):
        s = no.build_schedule(n)
        assert len(s) == n and build_schedule(n) == s
    s = no.build_schedule(1000)
    assert all(st == (0, False) for st in s[:75]) and all(st == (0, False) for st in s[950:])
    ends = [i for i, st in enumerate(s) if st[1]]
    assert ends == [99, 149, 249, 449, 949]          # windows of 25, 50, 100, 200, 500 slow steps
    assert all(st[0] == 1 for st in s[75:950])
