"""The re-ranking inference loop (rerank_loop.RerankLoop) against the reference's loop pattern (eval.py:82-196 restated in
rerank_loop.reference_style_loop): identical result dictionaries, one device->host copy."""
import pytest
import torch

from matchmaker_b200 import interaction
from matchmaker_b200.rerank_loop import RerankLoop, reference_style_loop
from oracle import interaction_oracle as O

pytestmark = pytest.mark.gpu
DEV = torch.device("cuda", 0)


def _batches(n_batches, B, seed):
    g = torch.Generator().manual_seed(seed)
    out = []
    for i in range(n_batches):
        b = B if i % 3 else B - 5          # ragged batch sizes
        q, d, qm, dm = O.synth_kernel_pool_inputs(b, 12, 48, 64, seed=seed + i)
        out.append({"query_id": [f"q{(i * 7 + j) % 11}" for j in range(b)], "doc_id": [f"d{i}_{j}" for j in range(b)],
                    "query_tokens": {"emb": q, "mask": qm}, "doc_tokens": {"emb": d, "mask": dm}})
    return out


def test_loop_matches_reference_pattern():
    mu = torch.tensor([1.0, 0.9, 0.7, 0.5, 0.3, 0.1, -0.1, -0.3, -0.5, -0.7, -0.9], device=DEV)
    sg = torch.full((11,), 0.1, device=DEV)
    w = torch.linspace(-0.1, 0.1, 11, device=DEV)

    def score_fn(b):
        return interaction.kernel_pool(b["query_tokens"]["emb"], b["doc_tokens"]["emb"], b["query_tokens"]["mask"],
                                       b["doc_tokens"]["mask"], mu, sg, w)["score"]

    batches = _batches(9, 40, seed=500)
    ref = reference_style_loop(score_fn, batches, DEV)
    loop = RerankLoop(score_fn, DEV, initial_capacity=64)   # forces the score buffer to grow
    got = loop.run(batches)
    assert got.keys() == ref.keys()
    for q in ref:
        assert [d for d, _ in got[q]] == [d for d, _ in ref[q]]
        assert [s for _, s in got[q]] == [s for _, s in ref[q]], "same kernels on the same inputs: bit-identical scores"
    # the caller's batches are untouched (the reference needs a deepcopy for that)
    assert all(not b["query_tokens"]["emb"].is_cuda for b in batches)
    scores, qids, dids = loop.run_flat(batches)
    assert scores.is_cuda and scores.numel() == sum(len(b["doc_id"]) for b in batches) == len(qids) == len(dids)
