"""Randomised stress of the tensor-core kernel-pooling training pair: many shapes (tile counts, odd feature-box counts,
pairs per CTA, mask patterns, with / without document gate), tensor-core backward against the FFMA backward on the same
saved S, twice per shape (bit-identical runs).  Test infrastructure."""
import os
import sys
import time

import torch

sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..", ".."))
from matchmaker_b200 import interaction, synthetic as O  # noqa: E402

DEV = "cuda"
N = int(os.environ.get("KPB_STRESS_N", "150"))
g = torch.Generator().manual_seed(int(os.environ.get("KPB_STRESS_SEED", "7")))


def ri(lo, hi):
    return int(torch.randint(lo, hi + 1, (1,), generator=g))


worst = 0.0
n_outliers = 0
t0 = time.time()
for it in range(N):
    Lq = ri(1, 32)
    Ld = [ri(1, 40), ri(100, 140), ri(120, 260), ri(250, 520), 128, 256, 129][ri(0, 6)]
    D = 4 * [ri(1, 8), ri(8, 24), ri(24, 80), 75, 80, 16][ri(0, 5)]
    K = [11, 21, ri(1, 32)][ri(0, 2)]
    B = [ri(1, 6), ri(100, 200), ri(140, 160), ri(290, 600)][ri(0, 3)]
    if B * (Lq + Ld) * D * 4 > 1.5e9:
        B = max(1, int(1.5e9 / ((Lq + Ld) * D * 4)))
    use_gate = ri(0, 3) == 0
    mu = torch.linspace(1.0, -0.9, K) if K > 1 else torch.tensor([0.3])
    sg = torch.full((K,), 0.1 + 0.05 * ri(0, 3))
    w = (torch.rand(K, generator=g) - 0.5) * 0.5
    alpha = torch.rand(K, generator=g) + 0.5
    gout = torch.randn(B, generator=g)
    q, d, qm, dm = O.synth_kernel_pool_inputs(B, Lq, Ld, D, seed=1000 + it)
    if ri(0, 4) == 0:   # holes in the document mask, an empty document
        dm = dm * (torch.rand(dm.shape, generator=g) > 0.3).float()
        dm[0] = 0
    gate = (torch.rand(B, Ld, generator=g) * 1.5 * dm) if use_gate else None
    args = [t.to(DEV) for t in (q, d, qm, dm, mu, sg, w)]
    cg = None if gate is None else gate.to(DEV)
    tr = interaction.kernel_pool(*args, alpha=alpha.to(DEV), save_for_backward=True, want_per_kernel=True, doc_gate=cg)
    res1 = interaction.kernel_pool_bwd(*args, alpha.to(DEV), tr["per_kernel_query"], gout.to(DEV), 1.0, saved=tr["saved"], doc_gate=cg)
    res2 = interaction.kernel_pool_bwd(*args, alpha.to(DEV), tr["per_kernel_query"], gout.to(DEV), 1.0, saved=tr["saved"], doc_gate=cg)
    ref = interaction.kernel_pool_bwd(*args, alpha.to(DEV), tr["per_kernel_query"], gout.to(DEV), 1.0, doc_gate=cg)
    torch.cuda.synchronize()
    names = ["dq", "dd", "dalpha", "dw"] + (["dgate"] if use_gate else [])
    for nm, a, b, r in zip(names, res1, res2, ref):
        if not torch.equal(a, b):
            print("NON-DETERMINISTIC", nm, (B, Lq, Ld, D, K, use_gate), flush=True)
            sys.exit(1)
        scale = r.abs().max().item()
        err = (a - r).abs().max().item()
        rel = err / max(scale, 1e-30)
        if not torch.isfinite(a).all():
            print("NON-FINITE", nm, (B, Lq, Ld, D, K, use_gate), flush=True)
            sys.exit(1)
        if rel > 1e-3 and scale > 1e-20 and nm in ("dq", "dd"):
            # how large are the terms the entry is summed from?  (tf32 operand error is relative to THEM)
            idx = tuple(torch.nonzero((a - r).abs() == (a - r).abs().max())[0].tolist())
            print("OUTLIER", nm, (B, Lq, Ld, D, K, use_gate), "rel %.3e scale %.3e at %s value %.3e; valid doc rows %d, valid query rows %d, sigma %.2f"
                  % (rel, scale, idx, r[idx].item(), int(dm[idx[0]].sum()), int(qm[idx[0]].sum()), sg[0].item()), flush=True)
            n_outliers += 1
        if nm in ("dq", "dd"):
            worst = max(worst, rel)
print("stress done: %d shapes in %.0f s, worst dq/dd deviation from the FFMA backward %.2e of the largest entry, %d entries above 1e-3" % (N, time.time() - t0, worst, n_outliers), flush=True)
