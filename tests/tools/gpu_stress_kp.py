import sys, os, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from matchmaker_b200 import interaction, synthetic as O
mu, sg = O.tk_21_kernels(); mu, sg = torch.tensor(mu).cuda(), torch.tensor(sg).cuda()
w = torch.linspace(-0.014, 0.014, 21).cuda(); alpha = torch.linspace(0.5, 1.5, 21).cuda()
for B in (256, 1000):
    q, d, qm, dm = [t.cuda() for t in O.synth_kernel_pool_inputs(B, 30, 200, 300, seed=1236)]
    base = interaction.kernel_pool(q, d, qm, dm, mu, sg, w, alpha=alpha, want_per_kernel=True, want_per_kernel_query=True, impl="tcgen05")
    torch.cuda.synchronize()
    for it in range(12):
        o = interaction.kernel_pool(q, d, qm, dm, mu, sg, w, alpha=alpha, want_per_kernel=True, want_per_kernel_query=True, impl="tcgen05")
        torch.cuda.synchronize()
        ds = (o["score"] != base["score"]).nonzero().flatten().tolist()
        dS = (o["per_kernel_query"] != base["per_kernel_query"])
        if ds or dS.any():
            idx = dS.nonzero()
            print(f"B={B} iter {it}: {len(ds)} scores differ (pairs {ds[:6]}), S diffs {int(dS.sum())}; first S idx {idx[:4].tolist()} "
                  f"max |dS| {(o['per_kernel_query'] - base['per_kernel_query']).abs().max().item():.3e}")
        else:
            print(f"B={B} iter {it}: identical")
