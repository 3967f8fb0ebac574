"""N>1 host logic on CPU: world_size-2 gloo process group, passages sharded over ranks, per-rank top-k,
all-gather + merge == single-process top-k (bit-exact ids under the (score desc, id asc) order)."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from matchmaker_b200 import sharding
from oracle import interaction_oracle as O


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, nq, n_pass, dim, k, out_dir):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        q, p = O.synth_dense_inputs(nq, n_pass, dim, seed=77, dtype=torch.float32)
        p[5] = p[n_pass - 3]  # duplicate passages across shards -> exact score ties
        p[6] = p[n_pass - 3]
        lo, hi = sharding.shard_bounds(n_pass, rank, world)
        scores = q @ p[lo:hi].T
        s, i = sharding.topk_all_gather_merge(scores, k, id_base=lo)
        torch.save((s, i), os.path.join(out_dir, f"r{rank}.pt"))
    finally:
        dist.destroy_process_group()


def test_shard_bounds_cover_everything():
    for n, w in [(10, 3), (8, 8), (5, 8), (1000003, 8)]:
        spans = [sharding.shard_bounds(n, r, w) for r in range(w)]
        assert spans[0][0] == 0 and spans[-1][1] == n
        assert all(spans[r][1] == spans[r + 1][0] for r in range(w - 1))
        assert max(b - a for a, b in spans) - min(b - a for a, b in spans) <= 1


def test_rank_topk_tie_break():
    s = torch.tensor([[1.0, 3.0, 3.0, 2.0, 3.0]])
    ids = torch.tensor([50, 40, 10, 20, 30])
    ts, ti = sharding.rank_topk(s, ids, 4)
    assert ti.tolist() == [[10, 30, 40, 20]] and ts.tolist() == [[3.0, 3.0, 3.0, 2.0]]


@pytest.mark.timeout(120)
def test_gloo_world2_topk_merge_matches_single_process(tmp_path):
    nq, n_pass, dim, k, world = 6, 501, 16, 10, 2
    port = _free_port()
    mp.spawn(_worker, args=(world, port, nq, n_pass, dim, k, str(tmp_path)), nprocs=world, join=True)
    q, p = O.synth_dense_inputs(nq, n_pass, dim, seed=77, dtype=torch.float32)
    p[5] = p[n_pass - 3]
    p[6] = p[n_pass - 3]
    ref_s, ref_i = O.flat_ip_search(q, p, torch.arange(n_pass), k)
    for r in range(world):
        s, i = torch.load(os.path.join(str(tmp_path), f"r{r}.pt"))
        assert torch.equal(i, ref_i), f"rank {r}: ids differ"
        assert torch.allclose(s, ref_s, rtol=1e-5, atol=1e-5)


def _loop_worker(rank, world, port, out_dir):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from matchmaker_b200 import rerank_loop as RL
        batches = [{"query_id": [f"q{i % 3}"] * 4, "doc_id": [f"d{i}_{j}" for j in range(4)]} for i in range(7)]
        local = {}
        for b in RL.shard_batches(batches, rank, world):
            for q, d in zip(b["query_id"], b["doc_id"]):
                local.setdefault(q, []).append((d, float(len(d))))
        merged = RL.gather_results(local)
        torch.save(merged, os.path.join(out_dir, f"m{rank}.pt"))
        m = RL.wrap_data_parallel(torch.nn.Linear(3, 1), torch.device("cpu"))
        assert type(m).__name__ == "DistributedDataParallel"
        m(torch.ones(2, 3)).sum().backward()
    finally:
        dist.destroy_process_group()


@pytest.mark.timeout(120)
def test_gloo_world2_rerank_sharding(tmp_path):
    """shard_batches / gather_results / wrap_data_parallel of the one-process-per-GPU re-ranking path, on gloo."""
    world, port = 2, _free_port()
    mp.spawn(_loop_worker, args=(world, port, str(tmp_path)), nprocs=world, join=True)
    merged = [torch.load(os.path.join(str(tmp_path), f"m{r}.pt")) for r in range(world)]
    assert merged[0].keys() == merged[1].keys() == {"q0", "q1", "q2"}
    for q in merged[0]:
        assert sorted(merged[0][q]) == sorted(merged[1][q])
    assert sum(len(v) for v in merged[0].values()) == 28
