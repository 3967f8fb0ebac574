"""Seeded synthetic inputs in the shapes of BASELINE.json's configs: GloVe-like
embeddings with planted exact matches, unit-norm ColBERT token vectors stored in fp16, TAS-B-like CLS vectors,
MSMARCO-shaped length distributions.  Pure data generation -- no scoring arithmetic -- shared by the tests,
bench.py and the golden-vector script so that every side sees identical tensors."""
from __future__ import annotations

from typing import List, Tuple

import torch

# ----------------------------------------------------------------------------


def _gen(seed: int) -> torch.Generator:
    return torch.Generator(device="cpu").manual_seed(seed)


def synth_lengths(n: int, mean: float, std: float, lo: int, hi: int, g: torch.Generator) -> torch.Tensor:
    return (torch.randn(n, generator=g) * std + mean).round().clamp(lo, hi).long()


def synth_kernel_pool_inputs(B: int, Lq: int, Ld: int, D: int, seed: int, full_q: bool = False,
                             copy_frac: float = 0.2):
    """GloVe-like embeddings (randn*0.4); a fraction of query rows is copied into
    random doc positions so the exact-match kernel fires; padded rows zero."""
    g = _gen(seed)
    q = torch.randn(B, Lq, D, generator=g) * 0.4
    d = torch.randn(B, Ld, D, generator=g) * 0.4
    q_len = torch.full((B,), Lq) if full_q else torch.randint(min(3, Lq), Lq + 1, (B,), generator=g)
    d_len = synth_lengths(B, min(75.0, Ld * 0.4), 30.0, min(10, Ld), Ld, g)
    for b in range(B):
        n_copy = max(1, int(copy_frac * int(q_len[b])))
        qi = torch.randint(0, int(q_len[b]), (n_copy,), generator=g)
        dj = torch.randint(0, int(d_len[b]), (n_copy,), generator=g)
        d[b, dj] = q[b, qi]
    q_mask = (torch.arange(Lq).unsqueeze(0) < q_len.unsqueeze(1)).float()
    d_mask = (torch.arange(Ld).unsqueeze(0) < d_len.unsqueeze(1)).float()
    q = q * q_mask.unsqueeze(-1)
    d = d * d_mask.unsqueeze(-1)
    return q, d, q_mask, d_mask


def synth_colbert_inputs(n_queries: int, docs_per_query: int, Lq: int, Ld: int, dim: int, seed: int,
                         dtype: torch.dtype = torch.float16, full_q: bool = True):
    """Unit-norm token vectors stored in ``dtype`` (reference storage dtype
    ``token_dtype: float16``); q_len = Lq (MASK-augmented queries), d_len ~
    clip(N(75,30),10,Ld); pad rows zero."""
    g = _gen(seed)
    n_docs = n_queries * docs_per_query
    q = torch.nn.functional.normalize(torch.randn(n_queries, Lq, dim, generator=g), dim=-1)
    d = torch.nn.functional.normalize(torch.randn(n_docs, Ld, dim, generator=g), dim=-1)
    q_len = torch.full((n_queries,), Lq) if full_q else torch.randint(min(2, Lq), Lq + 1, (n_queries,), generator=g)
    d_len = synth_lengths(n_docs, min(75.0, Ld * 0.42), 30.0, min(10, Ld), Ld, g)
    q_mask = (torch.arange(Lq).unsqueeze(0) < q_len.unsqueeze(1))
    d_mask = (torch.arange(Ld).unsqueeze(0) < d_len.unsqueeze(1))
    q = (q * q_mask.unsqueeze(-1)).to(dtype)
    d = (d * d_mask.unsqueeze(-1)).to(dtype)
    return q, d, q_mask.long(), d_mask.long()


def synth_dense_inputs(n_queries: int, n_passages: int, dim: int, seed: int,
                       dtype: torch.dtype = torch.float16, shard_id: int = 0):
    """TAS-B-like CLS vectors (unnormalised randn).  Passages are generated from
    ``seed + 1 + shard_id`` so every rank can build only its own slab."""
    gq = _gen(seed)
    q = torch.randn(n_queries, dim, generator=gq).to(dtype)
    gp = _gen(seed + 1 + shard_id)
    p = torch.randn(n_passages, dim, generator=gp).to(dtype)
    return q, p


def tk_21_kernels() -> Tuple[List[float], List[float]]:
    """BASELINE config 2 asks for 21 kernels: exact-match centre 1.0 plus 20
    evenly spaced centres from 0.95 to -0.95, sigma 0.05."""
    mus = [1.0] + [0.95 - 0.1 * i for i in range(20)]
    sig = [0.05] * 21
    return mus, sig
