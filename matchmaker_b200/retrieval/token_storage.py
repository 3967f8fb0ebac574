"""Reader / writer of the reference's dense-retrieval storage layout.

The reference's encode loop (matchmaker/dense_retrieval.py:197-286) writes passage vectors into numpy memmaps
``token_reps_<n>.npy`` of shape ``[token_block_size, token_dim]`` in ``token_dtype`` (fp16 in the documented
config), strips all-zero rows of multi-vector models (:244), records ``doc_infos[seq_id] = (block, start, end)``
(:259-265) and saves ``doc_infos.npz`` with ``doc_infos``, ``id_mapping``, ``seq_ids``, ``storage_filled_to_index``
(:279-286, loaded back at :291-302).  This module reads and writes exactly that layout so an index built by the
reference can be served by ``FlatIPIndexer`` / ``ColBERTTokenIndex`` and vice versa.  Host-side file I/O only; the
scoring stays in the kernels.
"""
from __future__ import annotations

import glob
import os
from typing import Dict, List, Tuple

import numpy as np


class TokenStorageWriter:
    """Append encoded vectors block by block (dense_retrieval.py:201-265)."""

    def __init__(self, folder: str, token_dim: int, token_block_size: int, token_dtype: str = "float16"):
        os.makedirs(folder, exist_ok=True)
        self.folder, self.dim, self.block_size, self.dtype = folder, token_dim, token_block_size, np.dtype(token_dtype)
        self.storage: List[np.memmap] = []
        self.filled: List[int] = []
        self.doc_infos: Dict[str, Tuple[int, int, int]] = {}
        self.id_mapping: List[List[int]] = []
        self.seq_ids: List[str] = []
        self._new_block()

    def _new_block(self):
        path = os.path.join(self.folder, "token_reps_" + str(len(self.storage)) + ".npy")
        self.storage.append(np.memmap(path, dtype=self.dtype, mode="w+", shape=(self.block_size, self.dim)))
        self.filled.append(0)
        self.id_mapping.append([])

    def add(self, seq_id: str, vectors: np.ndarray):
        """vectors: [dim] (single-vector model) or [n_tokens, dim]; all-zero rows are dropped (:244)."""
        v = np.asarray(vectors)
        if v.ndim == 1:
            v = v[np.newaxis, :]
        else:
            v = v[np.abs(v).sum(-1) != 0]
        n = len(v)
        if self.filled[-1] + n > self.block_size:
            self._new_block()
        b = len(self.storage) - 1
        lo = self.filled[b]
        self.storage[b][lo:lo + n] = v.astype(self.dtype)
        self.filled[b] = lo + n
        self.doc_infos[seq_id] = (b, lo, lo + n)
        self.id_mapping[b].extend([len(self.seq_ids)] * n)
        self.seq_ids.append(seq_id)

    def close(self):
        for m in self.storage:
            m.flush()
        np.savez(os.path.join(self.folder, "doc_infos.npz"), doc_infos=np.array(self.doc_infos, dtype=object),
                 id_mapping=np.array([np.array(x, dtype=np.int64) for x in self.id_mapping], dtype=object),
                 seq_ids=np.array(self.seq_ids), storage_filled_to_index=np.array(self.filled),
                 token_block_size=self.block_size, token_dim=self.dim, token_dtype=str(self.dtype))


def load_token_storage(folder: str):
    """Returns (storage blocks [list of memmaps cut to their fill level], id_mapping [list of int64 arrays],
    seq_ids, doc_infos) -- the four things dense_retrieval.py:291-302 restores."""
    meta = np.load(os.path.join(folder, "doc_infos.npz"), allow_pickle=True)
    filled = meta["storage_filled_to_index"]
    dim, block, dt = int(meta["token_dim"]), int(meta["token_block_size"]), np.dtype(str(meta["token_dtype"]))
    n_blocks = len(glob.glob(os.path.join(folder, "token_reps_*")))
    storage = [np.memmap(os.path.join(folder, f"token_reps_{f}.npy"), dtype=dt, mode="r", shape=(block, dim))[:filled[f]]
               for f in range(n_blocks)]
    id_mapping = [np.asarray(x, dtype=np.int64) for x in meta["id_mapping"]]
    return storage, id_mapping, [str(s) for s in meta["seq_ids"]], meta["doc_infos"].item()
