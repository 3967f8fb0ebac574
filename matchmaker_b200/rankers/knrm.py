"""KNRM with the cosine + kernel-pooling chain on the GPU kernel.  Mirrors matchmaker/models/knrm.py."""
from __future__ import annotations

from typing import List

import torch
import torch.nn as nn

from .. import autograd, interaction


def kernel_mus(n_kernels: int) -> List[float]:
    """Bin centres: exact-match 1.0, then the middles of n-1 equal bins over [-1, 1] (knrm.py:101-115)."""
    if n_kernels == 1:
        return [1.0]
    width = 2.0 / (n_kernels - 1)
    mus = [1.0, 1.0 - width / 2]
    while len(mus) < n_kernels:
        mus.append(mus[-1] - width)
    return mus


def kernel_sigmas(n_kernels: int) -> List[float]:
    """1e-4 for the exact-match kernel, half a bin width for the rest (knrm.py:117-131)."""
    if n_kernels == 1:
        return [0.0001]
    return [0.0001] + [0.5 * (2.0 / (n_kernels - 1))] * (n_kernels - 1)


class KNRM(nn.Module):
    """forward(query_embeddings [B,Lq,D], document_embeddings [B,Ld,D], query_pad_oov_mask [B,Lq],
    document_pad_oov_mask [B,Ld], output_secondary_output=False) -> score [B] (knrm.py:43-90).

    State dict: ``dense.weight`` only, as in the reference (mu / sigma are non-persistent buffers here;
    the reference keeps them as plain CUDA tensors, which is why it is not DataParallel-safe)."""

    @staticmethod
    def from_config(config, word_embeddings_out_dim):
        return KNRM(n_kernels=config["knrm_kernels"])

    def __init__(self, n_kernels: int):
        super().__init__()
        self.register_buffer("mu", torch.tensor(kernel_mus(n_kernels), dtype=torch.float32).view(1, 1, 1, n_kernels),
                             persistent=False)
        self.register_buffer("sigma", torch.tensor(kernel_sigmas(n_kernels), dtype=torch.float32).view(1, 1, 1, n_kernels),
                             persistent=False)
        self.dense = nn.Linear(n_kernels, 1, bias=False)
        torch.nn.init.uniform_(self.dense.weight, -0.014, 0.014)

    def forward(self, query_embeddings: torch.Tensor, document_embeddings: torch.Tensor,
                query_pad_oov_mask: torch.Tensor, document_pad_oov_mask: torch.Tensor,
                output_secondary_output: bool = False):
        # knrm.py:74 scales the log by 0.01; no alpha; both masks gate the kernels (a padded query row
        # contributes log(1e-10)*0.01*0 = 0 either way)
        score, per_kernel = autograd.kernel_pool(query_embeddings, document_embeddings, query_pad_oov_mask,
                                                 document_pad_oov_mask, self.mu, self.sigma, self.dense.weight,
                                                 None, 0.01)
        score = score.to(query_embeddings.dtype) if query_embeddings.dtype != torch.float32 else score
        if not output_secondary_output:
            return score
        with torch.no_grad():
            cos = interaction.kernel_pool(query_embeddings, document_embeddings, query_pad_oov_mask,
                                          document_pad_oov_mask, self.mu, self.sigma, self.dense.weight, None, 0.01,
                                          want_cosine=True)["cosine"]
        query_mean_vector = query_embeddings.sum(dim=1) / query_pad_oov_mask.sum(dim=1).unsqueeze(-1)
        return score, {"score": score, "per_kernel": per_kernel, "query_mean_vector": query_mean_vector,
                       "cosine_matrix_masked": cos}

    def forward_representation(self, sequence_embeddings: torch.Tensor, sequence_mask: torch.Tensor) -> torch.Tensor:
        return sequence_embeddings * sequence_mask.unsqueeze(-1)

    def get_param_stats(self):
        return "KNRM: linear weight: " + str(self.dense.weight.data)

    def get_param_secondary(self):
        return {"kernel_weight": self.dense.weight}
