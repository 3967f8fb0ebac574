"""Mirror of matchmaker/retrieval/base_index.py: the interface dense_retrieval.py drives."""
from typing import List

import numpy


class BaseNNIndexer:
    """prepare(data_chunks) / index(ids, data_chunks) / search(query_vec, top_n) -> (scores, ids)."""

    def __init__(self, config):
        self.token_dim = config["token_dim"]
        self.use_gpu = config["faiss_use_gpu"]
        self.use_fp16 = config["token_dtype"] == "float16"

    def prepare(self, data_chunks: List[numpy.ndarray], subsample=-1):
        pass

    def index(self, ids: List[numpy.ndarray], data_chunks: List[numpy.ndarray]):
        pass

    def search(self, query_vec: numpy.ndarray, top_n: int):
        pass
