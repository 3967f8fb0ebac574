"""Exact inner-product top-k (BERT_DOT retrieval scoring) vs the oracle (flat_ip_search).  Ids must be bit-exact
under the common order (score desc, id asc) wherever the fp32 scores are separated by more than round-off."""
import numpy as np
import pytest
import torch

from conftest import assert_close_rel
from matchmaker_b200 import _lib, interaction
from oracle import interaction_oracle as O

pytestmark = pytest.mark.gpu
DEV = "cuda"


def _check(q, p, ids, k, got_s, got_i):
    ref_s, ref_i = O.flat_ip_search(q.float(), p, ids, k)
    got_s, got_i = got_s.cpu(), got_i.cpu()
    assert_close_rel(got_s, ref_s, what="scores")
    same = got_i == ref_i
    if not same.all():
        # a swap is only acceptable between entries whose scores agree to fp32 accumulation round-off
        bad = (~same).nonzero()
        for qi, j in bad.tolist():
            assert abs(got_s[qi, j].item() - ref_s[qi, j].item()) <= 2e-5 * max(1.0, abs(ref_s[qi, j].item())), \
                f"query {qi} rank {j}: id {got_i[qi, j]} vs {ref_i[qi, j]} with different scores"
        assert same.float().mean() > 0.99
    # as sets the results must agree except for boundary ties
    for qi in range(q.shape[0]):
        a, b = set(got_i[qi].tolist()), set(ref_i[qi].tolist())
        assert len(a ^ b) <= 2


@pytest.mark.parametrize("shape", [(7, 3000, 64, 10), (130, 70000, 128, 100), (64, 20000, 768, 100),
                                   (5, 300, 64, 128), (200, 5000, 64, 256), (3, 50, 64, 100),
                                   # 11 query blocks: clusters of two with one idle slot; 2 blocks, many ranges
                                   (1300, 6000, 64, 10), (256, 40000, 64, 100)])
@pytest.mark.parametrize("dt", [torch.float16, torch.bfloat16])
def test_seeded_vs_oracle(shape, dt):
    nq, n, dim, k = shape
    q, p = O.synth_dense_inputs(nq, n, dim, seed=nq + n, dtype=dt)
    ids = torch.randperm(n, generator=torch.Generator().manual_seed(1)) * 3 + 7
    s, i = interaction.flat_ip_topk(q.to(DEV), p.to(DEV), k, ids=ids.to(DEV))
    _check(q, p, ids, k, s, i)
    s2, i2 = interaction.flat_ip_topk(q.to(DEV), p.to(DEV), k, id_base=1000)   # implicit ids
    _check(q, p, torch.arange(n) + 1000, k, s2, i2)


def test_exact_ties_resolved_by_id():
    q, p = O.synth_dense_inputs(4, 4000, 64, seed=3)
    p[1000:1300] = p[10]        # 301 identical passages: more ties than k
    ids = torch.arange(4000).flip(0)  # descending ids: the smallest ids sit at the END of the shard
    k = 50
    s, i = interaction.flat_ip_topk(q.to(DEV), p.to(DEV), k, ids=ids.to(DEV))
    ref_s, ref_i = O.flat_ip_search(q.float(), p, ids, k)
    assert torch.equal(i.cpu(), ref_i)
    assert_close_rel(s, ref_s, what="scores")


def test_topk_merge_kernel():
    g = torch.Generator().manual_seed(5)
    s = torch.randn(9, 700, generator=g)
    s[:, 100:110] = s[:, 5:6]  # ties
    ids = torch.stack([torch.randperm(100000, generator=g)[:700] for _ in range(9)])
    ids[:, 650:] = -1
    ms, mi = interaction.topk_merge(s.to(DEV), ids.to(DEV), 64)
    for r in range(9):
        valid = ids[r] >= 0
        rs, ri = O.rank_desc_stable(s[r][valid], ids[r][valid], 64)
        assert torch.equal(mi[r].cpu(), ri) and torch.equal(ms[r].cpu(), rs)


def test_baseline_cfg4_slab_properties():
    """One slab of BASELINE config 4 (dim 768, k=100): order invariance, sharded-and-merged == unsharded,
    oracle on a few queries."""
    nq, n, dim, k = 256, 200000, 768, 100
    q, p = O.synth_dense_inputs(nq, n, dim, seed=1238)
    cq, cp = q.to(DEV), p.to(DEV)
    s, i = interaction.flat_ip_topk(cq, cp, k)
    assert (s[:, :-1] >= s[:, 1:]).all(), "scores must be sorted descending"
    # "fake 8 shards" on one GPU: per-slab top-k then the merge kernel must equal the unsharded result
    parts_s, parts_i = [], []
    for r in range(8):
        lo, hi = r * n // 8, (r + 1) * n // 8
        ps, pi = interaction.flat_ip_topk(cq, cp[lo:hi], k, id_base=lo)
        parts_s.append(ps)
        parts_i.append(pi)
    ms, mi = interaction.topk_merge(torch.cat(parts_s, 1), torch.cat(parts_i, 1), k)
    assert torch.equal(mi, i) and torch.equal(ms, s)
    # permuting the passages (with their ids) does not change the answer
    perm = torch.randperm(n, generator=torch.Generator().manual_seed(9))
    s2, i2 = interaction.flat_ip_topk(cq, cp[perm.to(DEV)], k, ids=perm.to(DEV))
    assert torch.equal(i2, i) and torch.equal(s2, s)
    _check(q[:8], p, torch.arange(n), k, s[:8], i[:8])


def test_indexer_dropin_api():
    from matchmaker_b200.retrieval import FlatIPIndexer
    cfg = {"token_dim": 64, "faiss_use_gpu": True, "token_dtype": "float16"}
    idx = FlatIPIndexer(cfg)
    q, p = O.synth_dense_inputs(5, 2500, 64, seed=8)
    ids = np.arange(2500, dtype=np.int64) * 2
    chunks = [p[:1000].numpy(), p[1000:].numpy()]
    idx.prepare(chunks)
    idx.index([ids[:1000], ids[1000:]], chunks)
    s, i = idx.search(q.float().numpy(), 10)
    assert s.shape == (5, 10) and i.dtype == np.int64 and s.dtype == np.float32
    ref_s, ref_i = O.flat_ip_search(q.float(), p, torch.from_numpy(ids), 10)
    assert np.array_equal(i, ref_i.numpy())
    s1, i1 = idx.search(q[0].float().numpy(), 10)   # 1-d query
    assert np.array_equal(i1[0], i[0])
    with pytest.raises(_lib.MatchmakerB200Error):
        FlatIPIndexer({"token_dim": 64, "faiss_use_gpu": False, "token_dtype": "float16"})
