"""Run under torchrun on N GPUs: sharded FlatIPIndexer + NCCL top-k exchange vs the single-process oracle,
and the sharded max-sim re-ranking exchange."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np
import torch
import torch.distributed as dist

rank, world, lr = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(lr)
dev = torch.device("cuda", lr)
os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
dist.init_process_group("nccl", device_id=dev)
from matchmaker_b200 import interaction, sharding, synthetic as S
from matchmaker_b200.retrieval import FlatIPIndexer

# ---- BERT_DOT: passages sharded over ranks, one all-gather of per-query top-k, merge ----
nq, n, dim, k = 50, 60001, 128, 100
q, p = S.synth_dense_inputs(nq, n, dim, seed=11)
ids = np.arange(n, dtype=np.int64) * 2 + 1
idx = FlatIPIndexer({"token_dim": dim, "faiss_use_gpu": True, "token_dtype": "float16"}, device=dev)
idx.index([ids[:20000], ids[20000:]], [p[:20000].numpy(), p[20000:].numpy()])
lo, hi = sharding.shard_bounds(n, rank, world)
assert idx.passages.shape[0] == hi - lo
s, i = idx.search(q.float().numpy(), k)
if rank == 0:
    from oracle import interaction_oracle as O
    rs, ri = O.flat_ip_search(q.float(), p, torch.from_numpy(ids), k)
    match = (torch.from_numpy(i) == ri).float().mean().item()
    err = (torch.from_numpy(s) - rs).abs().max().item()
    print(f"[world {world}] FlatIPIndexer sharded search: id match {match:.4f}, max score err {err:.2e}", flush=True)
    assert match > 0.995 and err < 1e-2

# ---- ColBERT re-ranking: documents sharded, local max-sim, top-k exchange ----
n_q, dpq = 8, 400
qv, dv, qm, dm = S.synth_colbert_inputs(n_q, dpq * world, 32, 180, 128, seed=12)
# rank r owns documents [r*dpq, (r+1)*dpq) of every query
own = torch.cat([torch.arange(qi * dpq * world + rank * dpq, qi * dpq * world + (rank + 1) * dpq) for qi in range(n_q)])
sc = interaction.maxsim(qv.to(dev), dv[own].to(dev), qm.to(dev), dm[own].to(dev), docs_per_query=dpq).view(n_q, dpq)
gid = (torch.arange(dpq, device=dev) + rank * dpq).unsqueeze(0).expand(n_q, -1)  # id within the query's candidate list
ls, li = interaction.topk_merge(sc, gid.contiguous(), 50)
ms, mi = sharding.all_gather_merge(ls, li, 50)
if rank == 0:
    from oracle import interaction_oracle as O
    full = O.maxsim_one_query_many_docs(qv.float(), dv.float(), qm, dm, dpq * world).view(n_q, dpq * world)
    ts, ti = torch.topk(full, 50, dim=1)
    print(f"[world {world}] sharded max-sim rerank: id match {(mi.cpu() == ti).float().mean().item():.4f}, "
          f"max score err {(ms.cpu() - ts).abs().max().item():.2e}", flush=True)
    assert (mi.cpu() == ti).float().mean().item() > 0.99
dist.barrier()
dist.destroy_process_group()
