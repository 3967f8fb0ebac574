"""BERT_DOT with the pair dot product on the GPU kernel.  Mirrors matchmaker/models/bert_dot.py."""
from __future__ import annotations

from dataclasses import dataclass
from typing import Dict, Union

import torch
from torch import nn

from .. import autograd


@dataclass
class BERT_Dot_Config:
    """The reference's BERT_Dot_Config fields (bert_dot.py:7-12) as a plain dataclass."""
    bert_model: Union[str, nn.Module] = "distilbert-base-uncased"
    trainable: bool = True
    compress_dim: int = -1
    return_vecs: bool = False
    model_type: str = "BERT_Dot"


class BERT_Dot(nn.Module):
    """forward(query, document, use_fp16=True, output_secondary_output=False) -> score [B]
    (bert_dot.py:51-70).  State-dict keys: ``bert_model.*`` (+ ``compressor.*`` when compress_dim > -1)."""

    @staticmethod
    def from_config(config):
        cfg = BERT_Dot_Config()
        cfg.bert_model = config["bert_pretrained_model"]
        cfg.trainable = config["bert_trainable"]
        cfg.return_vecs = config.get("in_batch_negatives", False)
        cfg.compress_dim = config.get("bert_dot_compress_dim", -1)
        return BERT_Dot(cfg)

    def __init__(self, cfg: BERT_Dot_Config) -> None:
        super().__init__()
        self.config = cfg
        if isinstance(cfg.bert_model, str):
            from transformers import AutoModel
            self.bert_model = AutoModel.from_pretrained(cfg.bert_model)
        else:
            self.bert_model = cfg.bert_model
        for p in self.bert_model.parameters():
            p.requires_grad = cfg.trainable
        self.use_compressor = cfg.compress_dim > -1
        if self.use_compressor:
            self.compressor = torch.nn.Linear(self.bert_model.config.hidden_size, cfg.compress_dim)
        self.return_vecs = cfg.return_vecs

    def forward(self, query: Dict[str, torch.LongTensor], document: Dict[str, torch.LongTensor],
                use_fp16: bool = True, output_secondary_output: bool = False):
        with torch.autocast("cuda", enabled=use_fp16):
            query_vecs = self.forward_representation(query)
            document_vecs = self.forward_representation(document)
        if query_vecs.dtype != document_vecs.dtype:
            document_vecs = document_vecs.to(query_vecs.dtype)
        score = autograd.dot_pairs(query_vecs, document_vecs)  # bert_dot.py:62
        if use_fp16:
            score = score.to(query_vecs.dtype)
        if self.training and self.return_vecs:
            score = (score, query_vecs, document_vecs)
        if output_secondary_output:
            return score, {}
        return score

    def forward_representation(self, tokens: Dict[str, torch.LongTensor], sequence_type="n/a") -> torch.Tensor:
        vectors = self.bert_model(**tokens)[0][:, 0, :]
        if self.use_compressor:
            vectors = self.compressor(vectors)
        return vectors

    def get_param_stats(self):
        return "BERT_dot: / "

    def get_param_secondary(self):
        return {}
