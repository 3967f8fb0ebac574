"""Reader / writer / GPU loader of the reference's dense-retrieval storage layout.

The reference's encode loop (matchmaker/dense_retrieval.py:197-286) writes passage vectors into numpy memmaps
``token_reps_<n>.npy`` of shape ``[token_block_size, token_dim]`` in ``token_dtype`` (fp16 in the documented
config), strips all-zero rows of multi-vector models (:244), records ``doc_infos[seq_id] = (block, start, end)``
(:259-265) and saves ``doc_infos.npz`` -- through ``saveCompressed`` (utils/utils.py:196-203), an uncompressed zip
of ``.npy`` members -- with exactly four keys: ``doc_infos``, ``id_mapping``, ``seq_ids``,
``storage_filled_to_index`` (:284-285).  Block shape and dtype are NOT in the file: the reader takes them from the
config (``token_block_size``, ``token_dim``, ``token_dtype``; :299-300).

* :func:`load_token_storage` opens such a folder -- one written by the reference or by :class:`TokenStorageWriter`
  -- given those three config values.
* :class:`TokenStorageWriter` writes the same four keys (plus a ``storage_meta.json`` side file with the three config
  values, which the reference ignores and our reader uses only when the caller gives none).
* :func:`blocks_to_device` moves a row range of the blocks into one device tensor through the native loader
  (``mmb200_storage_load``: pread -> pinned staging -> cudaMemcpyAsync, no Python per-row work), which is how
  ``FlatIPIndexer.index`` and ``ColBERTTokenIndex.index_storage`` fill their HBM-resident stores.
"""
from __future__ import annotations

import ctypes
import glob
import json
import os
import zipfile
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np

META_FILE = "storage_meta.json"


def save_npz_stored(path: str, **arrays) -> None:
    """``saveCompressed`` of the reference (utils/utils.py:196-203): ZIP_STORED zip of ``<key>.npy`` members written
    with ``allow_pickle=True``; ``np.load(path, allow_pickle=True)`` reads it back."""
    with zipfile.ZipFile(path, mode="w", compression=zipfile.ZIP_STORED, allowZip64=True) as zf:
        for k, v in arrays.items():
            with zf.open(k + ".npy", "w", force_zip64=True) as buf:
                np.lib.format.write_array(buf, np.asanyarray(v), allow_pickle=True)


def _object_array(items: Sequence) -> np.ndarray:
    out = np.empty(len(items), dtype=object)
    for i, x in enumerate(items):
        out[i] = x
    return out


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
            v = v[np.abs(v).sum(-1) > 0]
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
        # exactly the reference's four keys (dense_retrieval.py:284-285); id_mapping is a list of per-block int64
        # arrays (ragged -> object array, what numpy.asanyarray made of it at the reference's pinned numpy version)
        save_npz_stored(os.path.join(self.folder, "doc_infos.npz"), doc_infos=self.doc_infos,
                        id_mapping=_object_array([np.array(x, dtype=np.int64) for x in self.id_mapping]),
                        seq_ids=self.seq_ids, storage_filled_to_index=self.filled)
        with open(os.path.join(self.folder, META_FILE), "w") as f:
            json.dump({"token_block_size": self.block_size, "token_dim": self.dim, "token_dtype": str(self.dtype)}, f)


def load_token_storage(folder: str, token_dim: Optional[int] = None, token_block_size: Optional[int] = None,
                       token_dtype: Optional[str] = None):
    """Open an encode folder (dense_retrieval.py:291-302).  ``token_dim`` / ``token_block_size`` / ``token_dtype`` come
    from the encode config, as in the reference; when omitted they are read from ``storage_meta.json`` (our writer's
    side file).  Returns (storage blocks [memmaps cut to their fill level], id_mapping [int64 array per block], seq_ids
    [list of str], doc_infos [dict seq_id -> (block, start, end)])."""
    meta = np.load(os.path.join(folder, "doc_infos.npz"), allow_pickle=True)
    side = {}
    if token_dim is None or token_block_size is None or token_dtype is None:
        sp = os.path.join(folder, META_FILE)
        if os.path.isfile(sp):
            side = json.load(open(sp))
        for k in ("token_dim", "token_block_size", "token_dtype"):  # round-1 folders carried them inside the npz
            if k not in side and k in meta.files:
                side[k] = meta[k].item() if meta[k].ndim == 0 else meta[k]
    dim = token_dim if token_dim is not None else side.get("token_dim")
    block = token_block_size if token_block_size is not None else side.get("token_block_size")
    dt = token_dtype if token_dtype is not None else side.get("token_dtype")
    if dim is None or block is None or dt is None:
        raise ValueError("load_token_storage: token_dim, token_block_size and token_dtype are not stored in doc_infos.npz "
                         "(the reference reads them from the encode config, dense_retrieval.py:299-300): pass them")
    dim, block, dt = int(dim), int(block), np.dtype(str(dt))
    filled = np.asarray(meta["storage_filled_to_index"]).reshape(-1)
    n_blocks = len(glob.glob(os.path.join(folder, "token_reps_*")))
    if n_blocks != len(filled):
        raise ValueError(f"{folder}: {n_blocks} token_reps_* files but storage_filled_to_index has {len(filled)} entries")
    storage = [np.memmap(os.path.join(folder, f"token_reps_{f}.npy"), dtype=dt, mode="r", shape=(block, dim))[:int(filled[f])]
               for f in range(n_blocks)]
    idm = meta["id_mapping"]
    if idm.dtype == object and idm.ndim == 0:
        idm = idm.item()
    id_mapping = [np.asarray(x, dtype=np.int64).reshape(-1) for x in idm]
    doc_infos = meta["doc_infos"]
    doc_infos = doc_infos.item() if doc_infos.ndim == 0 else dict(doc_infos)
    return storage, id_mapping, [str(s) for s in np.asarray(meta["seq_ids"]).reshape(-1)], doc_infos


def _file_segment(arr: np.ndarray, lo: int, hi: int):
    """(path, byte offset, n bytes) of rows [lo, hi) if `arr` is a C-contiguous row slice of a file-backed memmap."""
    base = arr
    while not isinstance(base, np.memmap) and getattr(base, "base", None) is not None:
        base = base.base
    mm = arr if isinstance(arr, np.memmap) else base
    if not isinstance(mm, np.memmap) or not getattr(mm, "filename", None) or not arr.flags["C_CONTIGUOUS"]:
        return None
    root = mm
    while isinstance(getattr(root, "base", None), np.memmap):
        root = root.base
    # address arithmetic against the mapping's first byte gives the file offset of arr[0]
    a0 = arr.__array_interface__["data"][0]
    r0 = root.__array_interface__["data"][0]
    off = int(root.offset) + (a0 - r0)
    row_bytes = arr.strides[0] if arr.ndim > 1 else arr.itemsize
    return str(mm.filename), off + lo * row_bytes, (hi - lo) * row_bytes


def blocks_to_device(blocks: Sequence[np.ndarray], lo: int, hi: int, device, staging_bytes: int = 32 << 20):
    """Rows [lo, hi) of the concatenation of `blocks` ([n_i, dim] arrays of one dtype, e.g. the memmaps of
    :func:`load_token_storage`) as ONE device tensor [hi-lo, dim] in the storage dtype.

    File-backed memmaps go through the native loader (pread -> pinned staging -> cudaMemcpyAsync on the current
    stream); in-memory arrays are staged through a pinned torch buffer (same double-buffered pattern).  Neither path
    touches rows outside [lo, hi)."""
    import torch
    from .. import _lib
    device = torch.device(device)
    if not blocks:
        raise ValueError("blocks_to_device: no blocks")
    np_dt = blocks[0].dtype
    dim = blocks[0].shape[1]
    t_dt = {np.dtype("float16"): torch.float16, np.dtype("float32"): torch.float32}.get(np.dtype(np_dt))
    if t_dt is None:
        raise ValueError(f"unsupported storage dtype {np_dt}")
    out = torch.empty((max(0, hi - lo), dim), dtype=t_dt, device=device)
    if hi <= lo:
        return out
    row_bytes = dim * np.dtype(np_dt).itemsize
    lib = _lib.load()
    off, written = 0, 0
    pend_paths, pend_off, pend_n, pend_dst = [], [], [], None

    def flush():
        nonlocal pend_paths, pend_off, pend_n, pend_dst
        if not pend_paths:
            return
        n = len(pend_paths)
        c_paths = (ctypes.c_char_p * n)(*[p.encode() for p in pend_paths])
        c_off = (ctypes.c_int64 * n)(*pend_off)
        c_n = (ctypes.c_int64 * n)(*pend_n)
        with torch.cuda.device(device):
            rc = lib.mmb200_storage_load(c_paths, c_off, c_n, n, pend_dst, staging_bytes,
                                         torch.cuda.current_stream(device).cuda_stream)
        _lib.check(rc, "mmb200_storage_load")
        pend_paths, pend_off, pend_n, pend_dst = [], [], [], None

    for blk in blocks:
        a, b = max(lo, off), min(hi, off + len(blk))
        if a < b:
            seg = _file_segment(blk, a - off, b - off)
            dst_ptr = out.data_ptr() + written * row_bytes
            if seg is not None:
                if pend_dst is None:
                    pend_dst = dst_ptr
                pend_paths.append(seg[0]); pend_off.append(seg[1]); pend_n.append(seg[2])
            else:
                flush()
                src = torch.from_numpy(np.ascontiguousarray(blk[a - off:b - off]))
                step = max(1, staging_bytes // row_bytes)
                for s0 in range(0, b - a, step):
                    piece = src[s0:s0 + step].pin_memory()
                    out[written + s0:written + s0 + len(piece)].copy_(piece, non_blocking=True)
            written += b - a
        off += len(blk)
    flush()
    if written != hi - lo:
        raise ValueError(f"blocks_to_device: rows [{lo},{hi}) exceed the {off} stored rows")
    return out
