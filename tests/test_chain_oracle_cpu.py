"""The chain-objective oracle (oracle/chain_oracle.py) on the CPU: its analytic gradient against finite differences of
its own loss -- the check the reference makes in internal/nnet/backward_test.go:24-140 (linear-chain numerator, epsilon
1e-4, tolerance 1e-3) -- plus the closed forms a linear chain allows."""
import numpy as np

from oracle import chain_oracle as CO


def test_linear_chain_closed_form():
    """one path: num_logprob = sum of the path's network outputs, numerator posteriors = 1 on the path"""
    T, P = 10, 20
    nnet = (np.sin(np.arange(T * P) * 0.1) * 0.5).astype(np.float16).astype(np.float32).reshape(T, P)   # backward_test.go:60-64
    num = CO.linear_chain_fst(T, P)
    a, b, tot = CO.forward_backward(nnet, num)
    want = sum(float(nnet[t, t % P]) for t in range(T))
    assert abs(tot - want) < 1e-9
    post = CO.posteriors(nnet, num, a, b, tot)
    assert np.allclose(post[np.arange(T), np.arange(T) % P], 1.0) and abs(post.sum() - T) < 1e-9


def test_denominator_posteriors_sum_to_one_per_frame():
    rng = np.random.default_rng(3)
    T, P = 12, 30
    den = CO.random_ergodic_fst(rng, 16, 4, P)
    nnet = (rng.standard_normal((T, P)) * 0.5).astype(np.float32)
    a, b, tot = CO.forward_backward(nnet, den)
    post = CO.posteriors(nnet, den, a, b, tot)
    assert np.allclose(post.sum(1), 1.0, atol=1e-9)
    # alpha/beta consistency: every frame gives the same total
    for t in range(T + 1):
        v = a[t] + b[t]
        v = v[(a[t] > CO.LOG_ZERO) & (b[t] > CO.LOG_ZERO)]
        assert abs(np.logaddexp.reduce(v) - tot) < 1e-9


def test_gradient_matches_finite_differences():
    rng = np.random.default_rng(11)
    T, P, eps = 10, 20, 1e-4
    den = CO.random_ergodic_fst(rng, 12, 3, P)
    num = CO.linear_chain_fst(T, P, offset=3)
    nnet = (np.sin(np.arange(T * P) * 0.1) * 0.5).reshape(T, P)

    def loss(x):
        _, _, tn = CO.forward_backward(x, num)
        _, _, td = CO.forward_backward(x, den)
        return -(tn - td)

    an, bn, tn = CO.forward_backward(nnet, num)
    ad, bd, td = CO.forward_backward(nnet, den)
    grad = CO.posteriors(nnet, den, ad, bd, td) - CO.posteriors(nnet, num, an, bn, tn)
    for _ in range(50):
        t, p = int(rng.integers(T)), int(rng.integers(P))
        xp, xm = nnet.copy(), nnet.copy()
        xp[t, p] += eps
        xm[t, p] -= eps
        num_grad = (loss(xp) - loss(xm)) / (2 * eps)
        assert abs(num_grad - grad[t, p]) <= 1e-3 * max(1.0, abs(grad[t, p])), (t, p, num_grad, grad[t, p])


def test_batch_rows_follow_the_subsampling_grid():
    rng = np.random.default_rng(5)
    n_seq, L, P, frames, sub, left = 3, 20, 16, 6, 3, 1
    out = (rng.standard_normal((n_seq * L, P)) * 0.3).astype(np.float16).astype(np.float32)
    den = CO.random_ergodic_fst(rng, 8, 3, P)
    nums = [CO.linear_chain_fst(frames, P, offset=s) for s in range(n_seq)]
    res, grad = CO.chain_loss_batch(out, nums, den, n_seq, L, frames, sub, left)
    used = np.zeros(n_seq * L, bool)
    for s in range(n_seq):
        used[s * L + left + np.arange(frames) * sub] = True
    assert not grad[~used].any() and np.abs(grad[used]).max() > 0
    assert np.allclose(res[:, 2], -(res[:, 0] - res[:, 1]))
