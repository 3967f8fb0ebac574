"""Where is the tensor-core backward's error for one shape?  Per feature box and per document tile, against fp64."""
import os
import sys

import torch

sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..", ".."))
from matchmaker_b200 import interaction, synthetic as O  # noqa: E402

DEV = "cuda"
B, Lq, Ld, D, K = [int(x) for x in os.environ.get("KPB_SHAPE", "1,26,296,316,21").split(",")]
for seed in range(int(os.environ.get("KPB_SEEDS", "6"))):
    g = torch.Generator().manual_seed(100 + seed)
    mu = torch.linspace(1.0, -0.9, K)
    sg = torch.full((K,), 0.1 + 0.05 * (seed % 4))
    w = (torch.rand(K, generator=g) - 0.5) * 0.5
    alpha = torch.rand(K, generator=g) + 0.5
    gout = torch.randn(B, generator=g)
    q, d, qm, dm = O.synth_kernel_pool_inputs(B, Lq, Ld, D, seed=2000 + seed)
    q64, d64 = q.double().requires_grad_(True), d.double().requires_grad_(True)
    qn = q64 / (q64.norm(dim=-1, keepdim=True) + 1e-13)
    dn = d64 / (d64.norm(dim=-1, keepdim=True) + 1e-13)
    cos = torch.bmm(qn, dn.transpose(-1, -2))
    raw = torch.exp(-torch.pow(cos.unsqueeze(-1) - mu.double().view(1, 1, 1, -1), 2) / (2 * sg.double().view(1, 1, 1, -1) ** 2))
    S = (raw * dm.double().unsqueeze(1).unsqueeze(-1)).sum(2)
    L = torch.log(torch.clamp(S * alpha.double().view(1, 1, -1), min=1e-10)) * qm.double().unsqueeze(-1)
    (L.sum(1) @ w.double()).backward(gout.double())
    args = [t.to(DEV) for t in (q, d, qm, dm, mu, sg, w)]
    tr = interaction.kernel_pool(*args, alpha=alpha.to(DEV), save_for_backward=True)
    res = interaction.kernel_pool_bwd(*args, alpha.to(DEV), tr["per_kernel_query"], gout.to(DEV), 1.0, saved=tr["saved"])
    ref = interaction.kernel_pool_bwd(*args, alpha.to(DEV), tr["per_kernel_query"], gout.to(DEV), 1.0)
    torch.cuda.synchronize()
    for nm, a, r, t64 in (("dq", res[0], ref[0], q64.grad), ("dd", res[1], ref[1], d64.grad)):
        a, r = a.double().cpu(), r.double().cpu()
        scale = t64.abs().max().item()
        e = (a - t64).abs()
        es = (r - t64).abs()
        per_box = [e[..., c:c + 32].max().item() / scale for c in range(0, D, 32)]
        per_tile = [e[:, t:t + 128].max().item() / scale for t in range(0, a.shape[1], 128)]
        idx = torch.nonzero(e == e.max())[0].tolist()
        print("seed %d sigma %.2f %s: tc %.2e ffma %.2e of max %.2e | worst at %s (value %.3e) | per box %s | per tile %s"
              % (seed, sg[0].item(), nm, e.max().item() / scale, es.max().item() / scale, scale, idx, t64[tuple(idx)].item(),
                 " ".join("%.1e" % x for x in per_box), " ".join("%.1e" % x for x in per_tile)), flush=True)
