"""Conv-KNRM (Dai et al., WSDM'18): n-gram convolutions in PyTorch, the n x n cross-match kernel pooling on the GPU
kernels.  Mirrors matchmaker/models/conv_knrm.py."""
from __future__ import annotations

import torch
import torch.nn as nn

from .. import autograd
from .knrm import kernel_mus, kernel_sigmas


class Conv_KNRM(nn.Module):
    """forward(query_embeddings [B,Lq,D], document_embeddings [B,Ld,D], query_pad_oov_mask, document_pad_oov_mask,
    output_secondary_output=False) -> score [B] (conv_knrm.py:57-141).

    State dict: ``convolutions.<i>.1.{weight,bias}`` and ``dense.weight`` [1, K*n*n], as in the reference."""

    @staticmethod
    def from_config(config, word_embeddings_out_dim):
        return Conv_KNRM(word_embeddings_out_dim=word_embeddings_out_dim, n_grams=config["conv_knrm_ngrams"],
                         n_kernels=config["conv_knrm_kernels"], conv_out_dim=config["conv_knrm_conv_out_dim"])

    def __init__(self, word_embeddings_out_dim: int, n_grams: int, n_kernels: int, conv_out_dim: int):
        super().__init__()
        self.n_grams, self.n_kernels = n_grams, n_kernels
        self.register_buffer("mu", torch.tensor(kernel_mus(n_kernels), dtype=torch.float32).view(1, 1, 1, n_kernels),
                             persistent=False)
        self.register_buffer("sigma", torch.tensor(kernel_sigmas(n_kernels), dtype=torch.float32).view(1, 1, 1, n_kernels),
                             persistent=False)
        self.convolutions = nn.ModuleList([
            nn.Sequential(nn.ConstantPad1d((0, i - 1), 0),
                          nn.Conv1d(kernel_size=i, in_channels=word_embeddings_out_dim, out_channels=conv_out_dim),
                          nn.ReLU())
            for i in range(1, n_grams + 1)])
        self.dense = nn.Linear(n_kernels * n_grams * n_grams, 1, bias=False)
        torch.nn.init.uniform_(self.dense.weight, -0.014, 0.014)

    def forward(self, query_embeddings: torch.Tensor, document_embeddings: torch.Tensor, query_pad_oov_mask: torch.Tensor,
                document_pad_oov_mask: torch.Tensor, output_secondary_output: bool = False):
        q_t, d_t = query_embeddings.transpose(1, 2), document_embeddings.transpose(1, 2)
        q_grams = [conv(q_t).transpose(1, 2).contiguous() for conv in self.convolutions]   # conv_knrm.py:115-122
        d_grams = [conv(d_t).transpose(1, 2).contiguous() for conv in self.convolutions]
        # :125-135: every (query n-gram, document n-gram) pair is kernel-pooled KNRM-style and the n*n per-kernel vectors
        # meet in dense(K*n*n -> 1).  Linear over a concatenation = sum of per-block linears, so each cross match is ONE
        # kernel launch that takes its own K-slice of dense.weight and returns its share of the score.
        w = self.dense.weight.view(self.n_grams * self.n_grams, self.n_kernels)
        score = None
        blk = 0
        for qg in q_grams:
            for dg in d_grams:
                s, _ = autograd.kernel_pool(qg, dg, query_pad_oov_mask, document_pad_oov_mask, self.mu, self.sigma,
                                            w[blk], None, 0.01)
                score = s if score is None else score + s
                blk += 1
        if output_secondary_output:
            return score, {}
        return score

    def get_param_stats(self):
        return "CONV-KNRM: linear weight: " + str(self.dense.weight.data)

    def get_param_secondary(self):
        return {"kernel_weight": self.dense.weight}
