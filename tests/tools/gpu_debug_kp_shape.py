"""Debug helper: one kernel-pooling shape through the tcgen05 kernels vs the oracle, with error statistics."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
from matchmaker_b200 import interaction
from oracle import interaction_oracle as O

B, Lq, Ld, D = [int(v) for v in sys.argv[1:5]]
mu, sg = O.tk_21_kernels()
mu, sg = torch.tensor(mu), torch.tensor(sg)
g = torch.Generator().manual_seed(B + Lq + Ld)
w = (torch.rand(len(mu), generator=g) - 0.5) * 0.03
alpha = torch.rand(len(mu), generator=g) + 0.5
q, d, qm, dm = O.synth_kernel_pool_inputs(B, Lq, Ld, D, seed=31 + Lq + Ld)
ref, sec = O.kernel_pool_tk(q, d, qm, dm, mu, sg, alpha, w)
for variant in ("ts", "ss"):
    os.environ["MMB200_KP_VARIANT"] = variant
    outs = []
    for rep in range(3):
        out = interaction.kernel_pool(q.cuda(), d.cuda(), qm.cuda(), dm.cuda(), mu.cuda(), sg.cuda(), w.cuda(), alpha=alpha.cuda(),
                                      log_scale=1.0, want_per_kernel=True, impl="tcgen05")
        outs.append(out["score"].cpu())
    err = (outs[0] - ref).abs() / ref.abs().clamp_min(1e-6)
    bad = (err > 1e-3).nonzero().flatten()
    print(variant, "max rel err", float(err.max()), "bad pairs", bad.tolist()[:20], "n_bad", len(bad),
          "run-to-run equal", bool((outs[0] == outs[1]).all() and (outs[1] == outs[2]).all()),
          "nan", int(torch.isnan(outs[0]).sum()))
    if len(bad):
        i = int(bad[0]); print("  pair", i, "got", float(outs[0][i]), "ref", float(ref[i]), "doc len", int(dm[i].sum()), "q len", int(qm[i].sum()))
