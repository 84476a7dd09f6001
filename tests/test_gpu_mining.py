"""GPU parity: all-pairs similarities (a11) and negative selection (a13) vs the oracle."""

import pytest
import torch

import oracle
from conftest import load_golden

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def M():
    from instance_search_b200 import mining
    return mining


def _check_negs(emb, lab, anchors, positives, semi, got_idx, got_sim):
    """got == oracle, except where fp32 noise in the oracle's S decides (adjudicate in fp64)."""
    S = oracle.mining.all_pairs_similarities(emb)
    couples = list(zip(anchors.tolist(), positives.tolist()))
    want = oracle.select_negatives(S, lab, couples, semi)
    got_idx = got_idx.cpu()
    bad = (want != got_idx).nonzero().flatten().tolist()
    S64 = emb.double() @ emb.double().t()
    for p in bad:
        a, b = couples[p]
        excl = (lab == lab[a])
        if semi:
            excl = excl | (S64[a] >= S64[a, b])
        if bool(excl.all()):
            assert int(got_idx[p]) == -1
            continue
        s = S64[a].clone()
        s[excl] = -2
        assert int(got_idx[p]) == int(s.argmax()), "couple %d: not the fp64 answer either" % p
        # and the oracle's disagreement is fp32 noise
        assert abs(float(S[a, want[p]]) - float(S[a, got_idx[p]])) < 2e-6 or int(want[p]) == -1 \
            or abs(float(S[a, b]) - float(S[a, want[p]])) < 2e-6
    ok = got_idx >= 0
    rows = torch.tensor([c[0] for c in couples])
    assert torch.allclose(got_sim.cpu()[ok], S[rows[ok], got_idx[ok]], rtol=1e-5, atol=1e-6)
    return len(bad)


def test_all_pairs_similarities(M):
    m = load_golden("mining_tiny")
    S = M.all_pairs_similarities(m["emb"].cuda())
    assert torch.allclose(S.cpu(), m["sim"], rtol=1e-5, atol=2e-6)
    g = torch.Generator().manual_seed(5)
    E = oracle.normalize_l2(torch.randn(700, 200, generator=g))
    S = M.all_pairs_similarities(E.cuda())
    assert torch.allclose(S.cpu(), oracle.mining.all_pairs_similarities(E), rtol=1e-5, atol=2e-6)
    assert torch.allclose(S.diagonal().cpu(), torch.ones(700), atol=1e-5)   # unit rows
    assert torch.allclose(S, S.t(), atol=2e-6)                               # symmetry


@pytest.mark.parametrize("semi", [False, True])
def test_select_negatives_golden(M, semi):
    m = load_golden("mining_tiny")
    idx = M.MiningIndex(m["emb"].cuda(), m["lab"])
    a, b = m["couples"][:, 0], m["couples"][:, 1]
    neg, nsim, psim = idx.select_negatives(a, b, semi)
    want = m["neg_semi" if semi else "neg_hard"]
    assert torch.equal(neg.cpu(), want)
    assert torch.allclose(psim.cpu(), m["sim"][a, b], rtol=1e-5, atol=1e-6)
    if semi:
        assert int(neg[-1]) == -1 and float(nsim[-1]) == -2.0   # all excluded -> caller falls back


@pytest.mark.parametrize("N,D,per,semi,terms", [
    (2048, 128, 16, False, 3),
    (2048, 128, 16, True, 3),
    (4096, 256, 16, True, 3),
    (1000, 72, 7, True, 3),       # ragged sizes
    (2048, 128, 16, True, 1),     # plain bf16 screen (the default): certificate + second line keep it exact
    (2048, 128, 16, False, 1),
    (20480, 128, 16, True, 1),    # CTA-pair screen, plain operands
    (20480, 128, 16, True, 3),    # large enough for the CTA-pair screen (masked epilogue, split operands)
    (20480, 136, 16, False, 3),   # the same, hard mode, D not a multiple of 64
])
def test_select_negatives_random(M, N, D, per, semi, terms):
    # SURVEY.md 8d cfg 3: E = normalize(center[label] + 0.5 randn), labels i // per
    g = torch.Generator().manual_seed(N + D)
    lab = torch.arange(N) // per
    centers = torch.randn(int(lab.max()) + 1, D, generator=g)
    E = oracle.normalize_l2(centers[lab] + 0.5 * torch.randn(N, D, generator=g))
    anchors = torch.randperm(N, generator=g)[:600]
    off = torch.randint(1, per, (600,), generator=g)
    positives = (anchors // per) * per + (anchors % per + off) % per
    positives = positives.clamp(max=N - 1)
    keep = (lab[anchors] == lab[positives]) & (anchors != positives)
    anchors, positives = anchors[keep], positives[keep]
    idx = M.MiningIndex(E.cuda(), lab, terms=terms)
    neg, nsim, psim = idx.select_negatives(anchors, positives, semi)
    n_adj = _check_negs(E, lab, anchors, positives, semi, neg, nsim)
    assert n_adj <= 3
    valid = neg.cpu() >= 0
    assert (lab[neg.cpu()[valid]] != lab[anchors[valid]]).all()
    if semi:
        assert (nsim.cpu()[valid] < psim.cpu()[valid]).all()
    assert int(idx.last_bruteforce) <= 2
    if terms == 1:
        assert idx.last_second_line <= len(anchors) // 20    # the plain screen certifies nearly every couple


def test_select_negatives_dense_neighbourhoods_fall_to_the_second_line(M):
    # negatives packed within bf16 noise of each other (all rows = one direction + 1e-4 noise, every
    # item its own label pair): the plain screen cannot certify, the split-operand second line /
    # the exhaustive pass answer -- and the answer is still the fp64 arg-max
    g = torch.Generator().manual_seed(77)
    N, D = 3000, 128
    base = torch.randn(1, D, generator=g)
    E = oracle.normalize_l2(base + 3e-3 * torch.randn(N, D, generator=g))
    lab = torch.arange(N) // 2
    anchors = torch.arange(0, 200, 2)
    positives = anchors + 1
    idx = M.MiningIndex(E.cuda(), lab)
    neg, nsim, psim = idx.select_negatives(anchors, positives, False)
    S64 = E.double() @ E.double().t()
    for p, (a, b) in enumerate(zip(anchors.tolist(), positives.tolist())):
        s = S64[a].clone()
        s[lab == lab[a]] = -2
        assert int(neg[p]) == int(s.argmax())
    assert idx.last_second_line > 0


@pytest.mark.parametrize("semi", [False, True])
def test_progressive_recheck_equals_the_full_recheck(M, semi):
    # only the candidates the current exact winner cannot exclude are gathered (option
    # mining_progressive, default on); the negatives must be those of the full re-check of all 16 --
    # on ordinary clustered rows, on rows whose best negatives are packed inside the screen noise
    # (many rounds), and in semi-hard mode where the best screen candidates sit at sim_pos (invalid:
    # the list has to be walked down)
    from instance_search_b200 import _lib
    g = torch.Generator().manual_seed(31 + semi)
    N, D, per = 6000, 192, 6
    lab = torch.arange(N) // per
    centers = torch.randn(N // per, D, generator=g)
    E = centers[lab] + 0.5 * torch.randn(N, D, generator=g)
    # a tight bundle: 300 rows of different labels within 2e-4 of one direction
    E[::20] = torch.randn(1, D, generator=g) * D ** 0.5 + 2e-3 * torch.randn(300, D, generator=g)
    E = oracle.normalize_l2(E)
    anchors = torch.arange(0, N, 3)[:900]
    positives = (anchors // per) * per + (anchors % per + 1) % per
    idx = M.MiningIndex(E.cuda(), lab)
    with _lib.options(mining_progressive=0):
        n0, s0, _ = idx.select_negatives(anchors, positives, semi)
        second0 = idx.last_second_line
    n1, s1, _ = idx.select_negatives(anchors, positives, semi)
    assert torch.equal(n0, n1) and torch.equal(s0, s1)
    assert idx.last_second_line == second0
    with _lib.options(mining_progressive=2):      # the one-CTA-per-couple kernel (what kc > 32 uses)
        n2, s2, _ = idx.select_negatives(anchors, positives, semi)
    assert torch.equal(n0, n2) and torch.equal(s0, s2)
    with _lib.options(mining_kc=40):              # ... and with a list longer than a warp
        n3, s3, _ = idx.select_negatives(anchors, positives, semi)
    assert torch.equal(n0, n3) and torch.equal(s0, s3)
    _check_negs(E, lab, anchors, positives, semi, n1, s1)


def test_lab_indicators_and_device_rule(M):
    ds = [(None, "b", "0"), (None, "a", "1"), (None, "b", "2")]
    ind = M.get_lab_indicators(ds, 0)
    assert set(ind) == {"a", "b"} and ind["b"].tolist() == [1, 0, 1] and ind["a"].dtype == torch.uint8
    o = oracle.get_lab_indicators(ds)
    assert all(torch.equal(ind[k].cpu(), o[k]) for k in o)

    class P(object):
        cuda_device, feature_dim, embeddings_cuda_size = 0, 2048, 2 ** 30

    class Net(object):
        feature_size = 2048
    assert M.embeddings_device_dim(P, Net, 16384, sim_matrix=True) == (0, 2048)
    assert M.embeddings_device_dim(P, Net, 16385, sim_matrix=True) == (-1, 2048)
