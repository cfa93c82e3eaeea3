"""GPU parity of the OPQ / PQ warm-up (run_warmup.py:85-132) against its numpy restatement
(oracle/warmup_np.py), step by step from identical states, plus the properties the algorithm guarantees.
Codes and cluster sizes are compared bit-exactly; means / errors within 1e-4 relative."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _dev(a):
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


def _mixture(n, D, ncl, seed, scale=0.35):
    r = np.random.default_rng(seed)
    centres = r.standard_normal((ncl, D)).astype(np.float32)
    return (centres[r.integers(0, ncl, n)] + scale * r.standard_normal((n, D))).astype(np.float32)


@pytest.mark.parametrize("K", [16, 256])
def test_code_histogram_matches_bincount(K):
    from repconc_b200 import warmup
    r = np.random.default_rng(0)
    B, M = 10007, 6
    codes = r.integers(0, K, (B, M))
    want = np.stack([np.bincount(codes[:, m], minlength=K) for m in range(M)])
    mb = _dev(codes.T.copy()).t()                      # the (B,M) view of an (M,B) buffer quantize returns
    assert np.array_equal(warmup.code_histogram(mb, K).cpu().numpy(), want)
    assert np.array_equal(warmup.code_histogram(_dev(codes), K).cpu().numpy(), want)
    if K == 256:
        assert np.array_equal(warmup.code_histogram(_dev(codes.astype(np.uint8)), K).cpu().numpy(), want)
    assert warmup.code_histogram(_dev(codes[:0]), K).sum().item() == 0     # empty input


@pytest.mark.parametrize("shape", [(4096, 32, 4, 16), (8192, 64, 4, 256), (3000, 48, 6, 256)])
def test_lloyd_steps_match_oracle(shape, oracle):
    from oracle import warmup_np as W
    from repconc_b200 import warmup
    n, D, M, K = shape
    x = _mixture(n, D, 40, 1)
    xd = _dev(x)
    c = warmup.initial_centroids(xd, M, K, seed=7).cpu().numpy()
    assert c.shape == (M, K, D // M)
    for it in range(4):
        want_c, want_obj, want_codes, want_counts = W.lloyd_step(x, c)
        got_c, got_obj, got_codes, got_counts = warmup.lloyd_step(xd, _dev(c))
        assert np.array_equal(got_codes.cpu().numpy(), want_codes), f"iteration {it}"
        assert np.array_equal(got_counts.cpu().numpy(), want_counts)
        np.testing.assert_allclose(got_obj.item(), want_obj, rtol=1e-4)
        np.testing.assert_allclose(got_c.cpu().numpy(), want_c, rtol=1e-4, atol=1e-6)
        c = want_c                                       # both sides continue from the oracle's state


def test_empty_clusters_are_split_like_the_oracle(oracle):
    from oracle import warmup_np as W
    from repconc_b200 import warmup
    x = _mixture(2048, 32, 10, 2)
    c = warmup.initial_centroids(_dev(x), 4, 16, seed=3).cpu().numpy()
    c[:, 3, :] += 100.0                                  # two unreachable centroids per sub-vector
    c[:, 11, :] -= 100.0
    want_c, _, _, want_counts = W.lloyd_step(x, c)
    got_c, _, _, got_counts = warmup.lloyd_step(_dev(x), _dev(c))
    assert (want_counts[:, 3] == 0).all() and (want_counts[:, 11] == 0).all()
    assert np.array_equal(got_counts.cpu().numpy(), want_counts)
    np.testing.assert_allclose(got_c.cpu().numpy(), want_c, rtol=1e-4, atol=1e-6)
    assert np.abs(got_c.cpu().numpy()[:, 3, :]).max() < 50.0       # re-seeded next to a populated cluster


def test_train_pq_objective_is_monotone_and_matches_oracle(oracle):
    from oracle import warmup_np as W
    from repconc_b200 import warmup
    x = _mixture(8192, 64, 300, 4)
    xd = _dev(x)
    c0 = warmup.initial_centroids(xd, 4, 256, seed=1234)
    c, objs = warmup.train_pq(xd, 4, 256, niter=12, seed=1234)
    assert all(b <= a * (1 + 1e-6) for a, b in zip(objs, objs[1:])), objs
    want_c, want_objs = W.train_pq(x, c0.cpu().numpy(), 12)
    np.testing.assert_allclose(objs, want_objs, rtol=2e-4)
    # same fixed point up to the few assignments that flip on fp32-vs-fp64 means
    assert np.mean(np.abs(c.cpu().numpy() - want_c) < 1e-3) > 0.99


def test_opq_alternation_matches_oracle_and_helps(oracle):
    from oracle import warmup_np as W
    from repconc_b200 import warmup
    r = np.random.default_rng(5)
    n, D, M, K = 4096, 32, 4, 16
    scales = np.geomspace(4.0, 0.05, D).astype(np.float32)
    mix = np.linalg.qr(r.standard_normal((D, D)))[0].astype(np.float32)
    x = ((_mixture(n, D, 30, 6) * scales) @ mix).astype(np.float32)      # strongly unbalanced sub-vectors
    xd = _dev(x)
    A0 = warmup.random_rotation(D, 11, "cuda")
    np.testing.assert_allclose((A0 @ A0.t()).cpu().numpy(), np.eye(D), atol=1e-5)
    # one alternation from identical state
    c0 = warmup.initial_centroids((xd @ A0.t()).contiguous(), M, K, seed=11)
    want_A, want_c, want_err = W.opq_alternation(x, A0.cpu().numpy(), c0.cpu().numpy(), 6)
    xproj = (xd @ A0.t()).contiguous()
    c1, _ = warmup.train_pq(xproj, M, K, 6, seed=11, init=c0, max_points_per_centroid=1000)
    from repconc_b200 import ops
    recons = ops.decode_forward(ops.nn_assign(xproj, c1), c1)
    got_err = ((recons - xproj) ** 2).sum(-1).mean().item()
    got_A = warmup.procrustes(xd, recons)
    np.testing.assert_allclose(got_err, want_err, rtol=1e-3)
    np.testing.assert_allclose((got_A @ got_A.t()).cpu().numpy(), np.eye(D), atol=1e-4)
    np.testing.assert_allclose(got_A.cpu().numpy(), want_A, atol=2e-2)
    # the full loop: orthogonal rotation, error below plain PQ on the unrotated vectors
    A, c, errs = warmup.train_opq(xd, M, K, niter=8, niter_pq=4, niter_pq_0=10, seed=11)
    np.testing.assert_allclose((A @ A.t()).cpu().numpy(), np.eye(D), atol=1e-4)
    _, objs_plain = warmup.train_pq(xd, M, K, niter=40, seed=11)
    assert errs[-1] < errs[0]
    assert errs[-1] < objs_plain[-1], (errs, objs_plain[-1])


def test_warmup_from_embeds_module_and_index(oracle):
    """run_warmup.warmup_from_embeds: rotation / centroids land in the module, the index holds the NN codes
    of the rotated corpus and can be searched."""
    from transformers import PretrainedConfig
    from repconc_b200 import RepCONC, warmup
    n, D, M = 5000, 64, 4
    x = _mixture(n, D, 200, 8)
    cfg = PretrainedConfig(hidden_size=D)
    cfg.MCQ_M, cfg.MCQ_K, cfg.similarity_metric = M, 256, "METRIC_IP"

    class Enc(torch.nn.Module):
        def __init__(self):
            super().__init__()
            self.config = cfg

    model = RepCONC(cfg, Enc(), False, None, None).cuda()
    model, index = warmup.warmup_from_embeds(x, model, opq_niter=3, pq_niter=5)
    A = model.rotation.cpu().numpy()
    c = model.centroids.detach().cpu().numpy()
    np.testing.assert_allclose(A @ A.T, np.eye(D), atol=1e-4)
    assert index.ntotal == n and index.pq.M == M
    np.testing.assert_array_equal(index.pq.centroid_array(), c)
    xrot = (_dev(x) @ model.rotation.t()).contiguous().cpu().numpy()
    want = oracle.nn_assign(xrot, c).astype(np.uint8)
    assert np.array_equal(index.code_array(), want)
    # quantisation error of the learnt PQ beats random centroids by a wide margin
    from oracle import warmup_np as W
    err = ((W.decode(want.astype(np.int64), c) - xrot) ** 2).sum(-1).mean()
    assert err < 0.5 * (xrot ** 2).sum(-1).mean()
    # and the index answers queries (a document's own rotated embedding ranks first for most documents)
    s, i = index.search(xrot[:64].astype(np.float32), 5)
    assert s.shape == (64, 5) and (i[:, 0] >= 0).all()

    cfg2 = PretrainedConfig(hidden_size=D)
    cfg2.MCQ_M, cfg2.MCQ_K, cfg2.similarity_metric = M, 256, "METRIC_CENTROID_COS"
    model2 = RepCONC(cfg2, Enc(), False, None, None).cuda()
    model2, _ = warmup.warmup_from_embeds(x, model2, opq_niter=1, pq_niter=2)
    norms = model2.centroids.detach().norm(dim=-1)
    np.testing.assert_allclose(norms.cpu().numpy(), 1.0, atol=1e-5)
