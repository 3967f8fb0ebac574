"""ColBERT retrieval aggregation over a GPU-resident token store.

The reference sketches this step in dense_retrieval.py:398-412 -- for every candidate passage of a query, fetch its
token matrix from the CPU memmap (`doc_infos[seq_id] = (block, start, end)`, :259-265), upcast, and call
`forward_aggregation` (colbert.py:100-112) one (query, passage) at a time in a Python loop -- but the loop reads
`current_ids` before assignment and an undefined `curr_q`, so it never ran.  Here the passages' token matrices stay in
HBM (fp16, padded to a common length with a mask) and ALL candidates of ALL queries are scored by one max-sim kernel
launch through the pair-index indirection (`pair_q`, `pair_d`): the candidate gather costs nothing extra because the
TMA coordinate of each document tile is just its index.
"""
from __future__ import annotations

from typing import List, Optional, Tuple

import numpy
import torch

from .. import _lib, interaction


class ColBERTTokenIndex:
    def __init__(self, token_dim: int, max_doc_length: int, device: Optional[torch.device] = None,
                 dtype: torch.dtype = torch.float16):
        self.token_dim, self.max_len, self.dtype = token_dim, max_doc_length, dtype
        self.device = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        self.tokens: Optional[torch.Tensor] = None   # [n, max_len, dim]
        self.mask: Optional[torch.Tensor] = None     # [n, max_len] bool
        self.ids: Optional[torch.Tensor] = None      # [n] int64 external ids

    def index(self, ids: numpy.ndarray, token_matrices: List[numpy.ndarray]):
        """token_matrices[i]: [len_i, dim] (all-zero rows already stripped, dense_retrieval.py:244)."""
        n = len(token_matrices)
        tok = torch.zeros((n, self.max_len, self.token_dim), dtype=self.dtype)
        msk = torch.zeros((n, self.max_len), dtype=torch.bool)
        for i, m in enumerate(token_matrices):
            L = min(len(m), self.max_len)
            tok[i, :L] = torch.from_numpy(numpy.ascontiguousarray(m[:L])).to(self.dtype)
            msk[i, :L] = True
        self.tokens, self.mask = tok.to(self.device), msk.to(self.device)
        self.ids = torch.from_numpy(numpy.asarray(ids, dtype=numpy.int64)).to(self.device)

    def rerank(self, query_vecs: torch.Tensor, query_mask: Optional[torch.Tensor], candidates: torch.Tensor,
               top_n: Optional[int] = None) -> Tuple[torch.Tensor, torch.Tensor]:
        """query_vecs [Nq, Lq, dim]; candidates [Nq, C] positions into the store (-1 = no candidate).
        Returns (scores [Nq, C or top_n], external ids) sorted by score descending."""
        if self.tokens is None:
            raise _lib.MatchmakerB200Error("rerank() before index()")
        nq, C = candidates.shape
        cand = candidates.to(self.device)
        valid = cand >= 0
        pair_d = cand.clamp(min=0).reshape(-1).to(torch.int32)
        pair_q = torch.arange(nq, device=self.device, dtype=torch.int32).repeat_interleave(C)
        s = interaction.maxsim(query_vecs.to(self.device, self.dtype), self.tokens,
                               None if query_mask is None else query_mask.to(self.device), self.mask,
                               pair_q=pair_q, pair_d=pair_d).view(nq, C)
        s = s.masked_fill(~valid, float("-inf"))
        ext = self.ids[cand.clamp(min=0)].masked_fill(~valid, -1)
        k = C if top_n is None else min(top_n, C)
        ms, mi = interaction.topk_merge(s, ext, k)
        return ms, mi
