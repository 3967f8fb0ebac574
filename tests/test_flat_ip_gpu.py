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
    """Scores within 1e-3 relative of the fp32 oracle; ids BIT-EXACT wherever the fp64 ranking separates neighbouring
    ranks by more than the fp32 accumulation bound (oracle.flat_ip_check_exact states the bound); inside an fp64
    near-tie run an id may only move within that run."""
    ref_s, ref_i = O.flat_ip_search(q.float(), p, ids, k)
    assert_close_rel(got_s.cpu(), ref_s, what="scores")
    stats = O.flat_ip_check_exact(q, p, ids, got_s, got_i, k)
    assert stats["decided"] >= 0.97 * (stats["decided"] + stats["undecided"]), stats  # the bound must not be vacuous


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


def test_storage_blocks_reach_the_gpu_through_the_native_loader(tmp_path):
    """Encode folder in the reference's layout -> load_token_storage -> FlatIPIndexer.index: the memmap blocks are read
    by mmb200_storage_load (pread -> pinned staging -> cudaMemcpyAsync); rows, order and ids must be preserved, also for
    a row range that starts and ends inside blocks and with a staging buffer smaller than a block."""
    from matchmaker_b200.retrieval import FlatIPIndexer
    from matchmaker_b200.retrieval.token_storage import TokenStorageWriter, blocks_to_device, load_token_storage
    dim, n = 64, 2500
    q, p = O.synth_dense_inputs(6, n, dim, seed=77)
    w = TokenStorageWriter(str(tmp_path), token_dim=dim, token_block_size=700, token_dtype="float16")
    for i in range(n):
        w.add(f"p{i}", p[i].numpy())
    w.close()
    storage, id_mapping, seq_ids, _ = load_token_storage(str(tmp_path), dim, 700, "float16")
    assert len(storage) == 4 and all(isinstance(s, np.memmap) for s in storage)
    full = blocks_to_device(storage, 0, n, DEV, staging_bytes=4096 * 3)
    assert torch.equal(full.cpu(), p)
    part = blocks_to_device(storage, 650, 1999, DEV, staging_bytes=1 << 20)
    assert torch.equal(part.cpu(), p[650:1999])
    mixed = blocks_to_device([storage[0], np.asarray(storage[1]).copy(), storage[2], storage[3]], 100, 2400, DEV)
    assert torch.equal(mixed.cpu(), p[100:2400])          # an in-memory block between file-backed ones
    idx = FlatIPIndexer({"token_dim": dim, "faiss_use_gpu": True, "token_dtype": "float16"})
    idx.index(id_mapping, storage)
    s, i = idx.search(q.float().numpy(), 10)
    ref_s, ref_i = O.flat_ip_search(q.float(), p, torch.arange(n), 10)
    assert np.array_equal(i, ref_i.numpy())
    idx.save(str(tmp_path / "faiss.index"))
    idx2 = FlatIPIndexer({"token_dim": dim, "faiss_use_gpu": True, "token_dtype": "float16"})
    idx2.load(str(tmp_path / "faiss.index"))
    s2, i2 = idx2.search(q.float().numpy(), 10)
    assert np.array_equal(i2, i) and np.array_equal(s2, s)
