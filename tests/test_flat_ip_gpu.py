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
    ids = torch.stack([torch.randperm(100000, generator=g)[:700] for _ in range(9)]) - 50000   # negative user ids are ids
    s[:, 650:] = float("-inf")          # void candidates are marked by their SCORE (-inf / -FLT_MAX / NaN), not by the id
    s[:, 640:650] = -3.4028234663852886e38
    ms, mi = interaction.topk_merge(s.to(DEV), ids.to(DEV), 64)
    for r in range(9):
        valid = s[r] > -3.0e38
        rs, ri = O.rank_desc_stable(s[r][valid], ids[r][valid], 64)
        assert torch.equal(mi[r].cpu(), ri) and torch.equal(ms[r].cpu(), rs)


def test_topk_merge_many_candidates_in_passes():
    """More candidates per query than one shared-memory sort holds (8192): merged in passes, same total order."""
    g = torch.Generator().manual_seed(6)
    nq, L, k = 5, 30000, 300
    s = torch.randn(nq, L, generator=g)
    s[:, 7000:7050] = s[:, 3:4]          # ties across pass boundaries
    ids = torch.stack([torch.randperm(10 ** 6, generator=g)[:L] for _ in range(nq)]) - 500000
    ms, mi = interaction.topk_merge(s.to(DEV), ids.to(DEV), k)
    for r in range(nq):
        rs, ri = O.rank_desc_stable(s[r], ids[r], k)
        assert torch.equal(mi[r].cpu(), ri) and torch.equal(ms[r].cpu(), rs)


@pytest.mark.parametrize("shape", [(40, 30000, 128, 1000), (130, 9000, 64, 300), (3, 500, 64, 1024), (300, 50000, 64, 1000)])
def test_large_k(shape):
    """top_n beyond 256 (dense_retrieval.py:391 passes any top_n; index_hit_top_n raises it to 1000): 2048-entry lists."""
    nq, n, dim, k = shape
    q, p = O.synth_dense_inputs(nq, n, dim, seed=nq + n + k)
    ids = torch.randperm(n, generator=torch.Generator().manual_seed(2)) * 5 - 1000   # some negative ids
    s, i = interaction.flat_ip_topk(q.to(DEV), p.to(DEV), k, ids=ids.to(DEV))
    _check(q, p, ids, k, s, i)


@pytest.mark.parametrize("shape", [(9, 4000, 64, 10), (70, 30000, 128, 100), (16, 6000, 768, 100)])
def test_fp32_storage_split(shape):
    """token_dtype float32 (faiss_indices.py:65,72: no useFloat16): fp32 vectors held as fp16 hi/lo halves; the result
    must track the fp64 ranking of the FP32 inputs (ids exact outside fp64 near-ties) and the scores to 1e-5 relative."""
    nq, n, dim, k = shape
    q, p = O.synth_dense_inputs(nq, n, dim, seed=nq + n, dtype=torch.float32)
    p[5] *= 37.0                                  # wide dynamic range inside one shard
    p[6] *= 1e-3
    ids = torch.arange(n) * 2 + 1
    ps, scale = interaction.flat_ip_split_f32(p.to(DEV), "passages")
    assert ps.shape == (n, 2 * dim) and ps.dtype == torch.float16
    s, i = interaction.flat_ip_topk(q.to(DEV), ps, k, ids=ids.to(DEV), split_scale=scale)
    s64 = q.double() @ p.double().T
    ref_s, ref_pos = torch.topk(s64, k, dim=1)
    got_s = s.cpu().double()
    assert ((got_s - ref_s).abs() <= 1e-5 * ref_s.abs().clamp(min=1.0)).all(), (got_s - ref_s).abs().max()
    ref_i = ids[ref_pos]
    gaps = (ref_s[:, :-1] - ref_s[:, 1:])
    decided = torch.ones_like(ref_i, dtype=torch.bool)
    tol = 2e-5 * ref_s.abs().clamp(min=1.0)
    decided[:, 1:] &= gaps > tol[:, 1:]
    decided[:, :-1] &= gaps > tol[:, :-1]
    assert (i.cpu()[decided] == ref_i[decided]).all()
    assert decided.float().mean() > 0.95


def test_indexer_fp32_and_large_top_n():
    from matchmaker_b200.retrieval import FlatIPIndexer
    idx = FlatIPIndexer({"token_dim": 64, "faiss_use_gpu": True, "token_dtype": "float32"})
    q, p = O.synth_dense_inputs(7, 3000, 64, seed=12, dtype=torch.float32)
    ids = np.arange(3000, dtype=np.int64) - 1500
    idx.index([ids], [p.numpy()])
    s, i = idx.search(q.numpy(), 400)
    assert s.shape == (7, 400) and i.shape == (7, 400)
    s64 = q.double() @ p.double().T
    ref_s, ref_pos = torch.topk(s64, 400, dim=1)
    assert np.allclose(s, ref_s.numpy(), rtol=1e-5, atol=1e-5)
    assert (i == ids[ref_pos.numpy()]).mean() > 0.98


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
