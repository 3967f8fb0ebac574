"""Re-ranking inference loop and data-parallel plumbing around the interaction kernels.

The reference's validation / test loop (matchmaker/eval.py:82-204) does, per batch: ``copy.deepcopy(batch)`` ->
``move_to_device`` (pageable host memory, synchronous) -> ``model.forward`` -> ``output.cpu()`` (a full device
synchronisation) -> a Python loop that appends ``(doc_id, float(output[i]))`` per sample.  Once the scoring op takes
tens of microseconds, every one of those steps is a stall: the GPU idles while the host copies, the host idles while the
GPU scores, and the ``.cpu()`` drains the pipeline 1 000 times per 256 000 pairs.

:class:`RerankLoop` keeps the same contract -- an iterable of batches in, ``{query_id: [(doc_id, score), ...]}`` out --
with none of the round trips:

* batches are staged through a small ring of PINNED host buffers and uploaded on a copy stream while the previous batch
  is being scored (``cudaMemcpyAsync`` overlap; no deepcopy -- nothing mutates the caller's batch);
* scores are written into ONE preallocated device buffer at a running offset; ids stay on the host as Python lists;
* there is ONE device->host copy per flush (default: at the end), and the result dictionary is built from it in one pass.

Multi-GPU: the reference wraps the model in ``nn.DataParallel`` (train.py:193-202: one process, a GIL-bound scatter /
replicate / gather per step).  :func:`shard_batches` + :func:`gather_results` run the loop one process per GPU over
``torch.distributed`` (NCCL): rank r scores batches r, r + world, ...; the only collective is the all-gather of the score
vectors at the end.  :func:`wrap_data_parallel` is the training-side replacement (DistributedDataParallel over NCCL).
"""
from __future__ import annotations

from typing import Any, Callable, Dict, Iterable, Iterator, List, Optional, Tuple

import torch


def _map_tensors(obj: Any, fn: Callable[[torch.Tensor], torch.Tensor]) -> Any:
    if isinstance(obj, torch.Tensor):
        return fn(obj)
    if isinstance(obj, dict):
        return {k: _map_tensors(v, fn) for k, v in obj.items()}
    if isinstance(obj, (list, tuple)) and obj and isinstance(obj[0], (torch.Tensor, dict)):
        return type(obj)(_map_tensors(v, fn) for v in obj)
    return obj


def _tensor_bytes(obj: Any) -> int:
    total = 0

    def acc(t):
        nonlocal total
        total += (t.numel() * t.element_size() + 255) // 256 * 256
        return t
    _map_tensors(obj, acc)
    return total


class _PinnedRing:
    """`depth` pinned host arenas; a batch's tensors are packed into one arena back to back (256-byte aligned) so that the
    upload is a few large async copies from page-locked memory.  An arena is reused only after the copies that read it
    have completed (event per arena)."""

    def __init__(self, depth: int, device: torch.device):
        self.depth, self.device = depth, device
        self.arenas: List[Optional[torch.Tensor]] = [None] * depth
        self.events: List[Optional[torch.cuda.Event]] = [None] * depth
        self.slot = 0

    def stage(self, batch: Any, copy_stream: torch.cuda.Stream) -> Any:
        i = self.slot
        self.slot = (self.slot + 1) % self.depth
        need = _tensor_bytes(batch)
        if self.events[i] is not None:
            self.events[i].synchronize()
        if self.arenas[i] is None or self.arenas[i].numel() < need:
            self.arenas[i] = torch.empty(max(need, 1), dtype=torch.uint8, pin_memory=True)
        arena, off = self.arenas[i], 0

        def up(t: torch.Tensor) -> torch.Tensor:
            nonlocal off
            if t.is_cuda:
                return t
            nbytes = t.numel() * t.element_size()
            view = arena[off:off + nbytes].view(t.dtype).view(t.shape)
            view.copy_(t)   # pageable -> pinned, on the host cores
            off += (nbytes + 255) // 256 * 256
            return view.to(self.device, non_blocking=True)

        with torch.cuda.stream(copy_stream):
            dev_batch = _map_tensors(batch, up)
            ev = torch.cuda.Event()
            ev.record(copy_stream)
        self.events[i] = ev
        return dev_batch, ev


class RerankLoop:
    """``score_fn(device_batch) -> scores [b]`` (any callable: a drop-in ranker's forward, or the interaction op alone).

    ``run(batches)`` consumes an iterable of dicts; keys ``id_keys`` (default ``("query_id", "doc_id")``) hold per-sample
    ids (lists / numpy arrays, left on the host), every tensor anywhere else in the dict is uploaded."""

    def __init__(self, score_fn: Callable[[Dict[str, Any]], torch.Tensor], device: Optional[torch.device] = None,
                 id_keys: Tuple[str, str] = ("query_id", "doc_id"), prefetch: int = 3, initial_capacity: int = 1 << 16):
        self.score_fn = score_fn
        self.device = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        self.id_keys = id_keys
        self.prefetch = max(2, prefetch)
        self.capacity = initial_capacity

    def _scores_and_ids(self, batches: Iterable[Dict[str, Any]]):
        dev = self.device
        copy_stream = torch.cuda.Stream(device=dev)
        ring = _PinnedRing(self.prefetch, dev)
        compute = torch.cuda.current_stream(dev)
        out = torch.empty(self.capacity, dtype=torch.float32, device=dev)
        n = 0
        qids: List[Any] = []
        dids: List[Any] = []
        it: Iterator[Dict[str, Any]] = iter(batches)
        pending = None   # (device batch, event, ids) uploaded ahead of the batch being scored

        def fetch():
            try:
                b = next(it)
            except StopIteration:
                return None
            ids = (list(b[self.id_keys[0]]), list(b[self.id_keys[1]]))
            payload = {k: v for k, v in b.items() if k not in self.id_keys}
            dev_batch, ev = ring.stage(payload, copy_stream)
            return dev_batch, ev, ids

        pending = fetch()
        with torch.no_grad():
            while pending is not None:
                dev_batch, ev, ids = pending
                pending = fetch()          # upload of batch i+1 is enqueued before batch i is scored
                compute.wait_event(ev)
                s = self.score_fn(dev_batch).reshape(-1).float()
                b = s.numel()
                if n + b > out.numel():
                    grown = torch.empty(max(2 * out.numel(), n + b), dtype=torch.float32, device=dev)
                    grown[:n] = out[:n]
                    out = grown
                out[n:n + b] = s
                # the arena of this batch may be refilled only after its tensors have been consumed by the kernels
                _map_tensors(dev_batch, lambda t: (t.record_stream(compute), t)[1])
                n += b
                qids.extend(ids[0])
                dids.extend(ids[1])
        self.capacity = max(self.capacity, out.numel())
        return out[:n], qids, dids

    def run(self, batches: Iterable[Dict[str, Any]]) -> Dict[Any, List[Tuple[Any, float]]]:
        """The reference's ``validation_results`` (eval.py:190-192): query id -> list of (doc id, score) in input order."""
        scores, qids, dids = self._scores_and_ids(batches)
        host = torch.empty(scores.numel(), dtype=torch.float32, pin_memory=True)
        host.copy_(scores, non_blocking=True)        # the ONE device->host copy
        torch.cuda.current_stream(self.device).synchronize()
        vals = host.tolist()
        results: Dict[Any, List[Tuple[Any, float]]] = {}
        for q, d, v in zip(qids, dids, vals):
            results.setdefault(q, []).append((d, v))
        return results

    def run_flat(self, batches: Iterable[Dict[str, Any]]):
        """(scores [n] device tensor, query ids, doc ids) without the final host copy (for callers that keep ranking on
        the GPU, e.g. a top-k per query)."""
        return self._scores_and_ids(batches)


def reference_style_loop(score_fn, batches: Iterable[Dict[str, Any]], device, id_keys=("query_id", "doc_id")):
    """The reference's loop pattern (eval.py:82-196) restated for A/B timing: deepcopy, synchronous move_to_device,
    forward, ``.cpu()`` per batch, per-sample Python appends.  Not used by the product path."""
    import copy
    results: Dict[Any, List[Tuple[Any, float]]] = {}
    with torch.no_grad():
        for batch_orig in batches:
            batch = _map_tensors(copy.deepcopy({k: v for k, v in batch_orig.items() if k not in id_keys}),
                                 lambda t: t.to(device))
            output = score_fn(batch).cpu()
            for i, q in enumerate(batch_orig[id_keys[0]]):
                results.setdefault(q, []).append((batch_orig[id_keys[1]][i], float(output[i])))
    return results


# ------------------------------------------------------------------------------------------------------------------
# one process per GPU
# ------------------------------------------------------------------------------------------------------------------
def shard_batches(batches: Iterable[Dict[str, Any]], rank: int, world: int) -> Iterator[Dict[str, Any]]:
    """Batches rank, rank + world, ... (re-ranking pairs are independent units: no data-path collective)."""
    for i, b in enumerate(batches):
        if i % world == rank:
            yield b


def gather_results(local: Dict[Any, List[Tuple[Any, float]]], group=None) -> Dict[Any, List[Tuple[Any, float]]]:
    """Union of the per-rank result dictionaries on every rank (all_gather_object: ids are arbitrary Python values; the
    scores of one 256 000-pair validation set are 1 MB)."""
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return local
    parts: List[Optional[dict]] = [None] * dist.get_world_size(group)
    dist.all_gather_object(parts, local, group=group)
    merged: Dict[Any, List[Tuple[Any, float]]] = {}
    for p in parts:
        for q, lst in p.items():
            merged.setdefault(q, []).extend(lst)
    return merged


def wrap_data_parallel(model: torch.nn.Module, device: torch.device, group=None) -> torch.nn.Module:
    """Training-side replacement of ``nn.DataParallel(model)`` (train.py:193-202): with an initialised process group the
    model is wrapped in DistributedDataParallel (one process per GPU, NCCL all-reduce of the gradients overlapped with
    backward); in a single-process run it is returned unchanged."""
    import torch.distributed as dist
    model = model.to(device)
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        from torch.nn.parallel import DistributedDataParallel
        return DistributedDataParallel(model, device_ids=[device.index] if device.type == "cuda" else None,
                                       process_group=group, broadcast_buffers=False)
    return model
