"""Per-role wait profile of the tensor-core kernel-pooling backward at the bench shape (needs a --prof build:
MMB200_LIB=scripts/ab_prof_libmatchmaker_b200.so MMB200_KPB_PROF=1)."""
import os
import sys

import torch

sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..", ".."))
from matchmaker_b200 import interaction, synthetic as O  # noqa: E402

DEV = "cuda"
B, Lq, Ld, D, K = int(os.environ.get("KPB_B", "1024")), 30, 200, 300, 21
mu, sg = O.tk_21_kernels()
mu, sg = torch.tensor(mu).to(DEV), torch.tensor(sg).to(DEV)
w = torch.linspace(-0.3, 0.3, K).to(DEV)
q, d, qm, dm = [t.to(DEV) for t in O.synth_kernel_pool_inputs(B, Lq, Ld, D, seed=1)]
gout = torch.ones(B, device=DEV)
tr = interaction.kernel_pool(q, d, qm, dm, mu, sg, w, save_for_backward=True)
for _ in range(3):
    interaction.kernel_pool_bwd(q, d, qm, dm, mu, sg, w, None, tr["per_kernel_query"], gout, 1.0, saved=tr["saved"])
torch.cuda.synchronize()
if not os.environ.get("MMB200_KPB_PROF"):
    def step():
        return interaction.kernel_pool_bwd(q, d, qm, dm, mu, sg, w, None, tr["per_kernel_query"], gout, 1.0, saved=tr["saved"])

    def timeit(fn, n=20):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        e0.record()
        for _ in range(n):
            fn()
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / n
    ms_eager = timeit(step)
    side = torch.cuda.Stream()
    with torch.cuda.stream(side):
        step()
    torch.cuda.synchronize()
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.graph(graph):
        out = step()
    graph.replay()
    timeit(graph.replay, 300)   # let the clocks settle
    ms = timeit(graph.replay, 200)
    print("backward tcgen05 [boxes %s stages %s]: graph replay %.4f ms, eager %.4f ms per %d pairs (%.0f GB/s of 558 KB/pair)"
          % (os.environ.get("MMB200_KPB_BOXES", "-"), os.environ.get("MMB200_KPB_STAGES", "-"), ms, ms_eager, B, B * 557964 / ms / 1e6), flush=True)
