"""Bring-up of the tensor-core kernel-pooling backward on a B200 (run under `timeout`): one pair first, then the shapes
of the parity test, error statistics with / without the truncation compensation, one- vs two-box stages, timing at the
bench shape.  Test infrastructure (imports oracle/)."""
import os
import sys
import time

import torch

sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..", ".."))
sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))
from matchmaker_b200 import interaction  # noqa: E402
from oracle import interaction_oracle as O  # noqa: E402

DEV = "cuda"


def fp64_grads(q, d, qm, dm, mu, sg, alpha, w, ls, gout):
    q64 = q.double().requires_grad_(True)
    d64 = d.double().requires_grad_(True)
    qn = q64 / (q64.norm(dim=-1, keepdim=True) + 1e-13)
    dn = d64 / (d64.norm(dim=-1, keepdim=True) + 1e-13)
    cos = torch.bmm(qn, dn.transpose(-1, -2))
    raw = torch.exp(-torch.pow(cos.unsqueeze(-1) - mu.double().view(1, 1, 1, -1), 2) / (2 * sg.double().view(1, 1, 1, -1) ** 2))
    S = (raw * dm.double().unsqueeze(1).unsqueeze(-1)).sum(2)
    L = torch.log(torch.clamp(S * alpha.double().view(1, 1, -1), min=1e-10)) * ls * qm.double().unsqueeze(-1)
    score = L.sum(1) @ w.double()
    score.backward(gout.double())
    return q64.grad, d64.grad


def run(B, Lq, Ld, D, K=21, tag=""):
    if K == 21:
        mu, sg = O.tk_21_kernels()
    else:
        mu, sg = [1.0, 0.9, 0.7, 0.5, 0.3, 0.1, -0.1, -0.3, -0.5, -0.7, -0.9], [0.1] * 11
    mu, sg = torch.tensor(mu), torch.tensor(sg)
    g = torch.Generator().manual_seed(3)
    w = (torch.rand(K, generator=g) - 0.5) * 0.5
    alpha = torch.rand(K, generator=g) + 0.5
    gout = torch.randn(B, generator=g)
    q, d, qm, dm = O.synth_kernel_pool_inputs(B, Lq, Ld, D, seed=9)
    gq, gd = fp64_grads(q, d, qm, dm, mu, sg, alpha, w, 1.0, gout)
    args = [t.to(DEV) for t in (q, d, qm, dm, mu, sg, w)]
    tr = interaction.kernel_pool(*args, alpha=alpha.to(DEV), save_for_backward=True, want_per_kernel=True)
    torch.cuda.synchronize()
    res = interaction.kernel_pool_bwd(*args, alpha.to(DEV), tr["per_kernel_query"], gout.to(DEV), 1.0, saved=tr["saved"])
    torch.cuda.synchronize()
    ref = interaction.kernel_pool_bwd(*args, alpha.to(DEV), tr["per_kernel_query"], gout.to(DEV), 1.0)
    torch.cuda.synchronize()

    def rel(a, b):
        a, b = a.double().cpu(), b.double().cpu()
        return ((a - b).abs().max() / b.abs().max()).item(), (((a - b) * b).sum() / (b * b).sum()).item()
    print("%-28s B %d Lq %d Ld %d D %d K %d | dq tc %.2e (bias %+.2e) simt %.2e | dd tc %.2e (bias %+.2e) simt %.2e | ga %.1e gw %.1e"
          % (tag, B, Lq, Ld, D, K, *rel(res[0], gq), rel(ref[0], gq)[0], *rel(res[1], gd), rel(ref[1], gd)[0],
             rel(res[2], ref[2])[0], rel(res[3], ref[3])[0]), flush=True)


def main():
    print(torch.cuda.get_device_name(0), flush=True)
    run(1, 30, 100, 64, 11, "one pair, one tile")
    run(1, 30, 200, 300, 21, "one pair, cfg2 shape")
    os.environ["MMB200_KPB_BOXES"] = "1"
    run(3, 30, 200, 300, 21, "one box per stage")
    del os.environ["MMB200_KPB_BOXES"]
    os.environ["MMB200_KPB_COMP"] = "1.0"
    run(8, 30, 200, 300, 21, "no truncation compensation")
    del os.environ["MMB200_KPB_COMP"]
    run(8, 30, 200, 300, 21, "default")
    run(400, 8, 20, 32, 11, "many pairs per CTA")
    run(6, 30, 300, 100, 21, "three tiles")
    run(3, 9, 129, 96, 11, "odd boxes")
    # timing at the bench shape
    B, Lq, Ld, D, K = 1024, 30, 200, 300, 21
    mu, sg = O.tk_21_kernels()
    mu, sg = torch.tensor(mu).to(DEV), torch.tensor(sg).to(DEV)
    w = torch.linspace(-0.3, 0.3, K).to(DEV)
    q, d, qm, dm = [t.to(DEV) for t in O.synth_kernel_pool_inputs(B, Lq, Ld, D, seed=1)]
    gout = torch.ones(B, device=DEV)
    tr = interaction.kernel_pool(q, d, qm, dm, mu, sg, w, save_for_backward=True)
    for name, kw in (("tcgen05", dict(saved=tr["saved"])), ("simt", {})):
        for stages in ((None, 2, 3) if name == "tcgen05" else (None,)):
            if stages is None:
                os.environ.pop("MMB200_KPB_STAGES", None)
            else:
                os.environ["MMB200_KPB_STAGES"] = str(stages)
            for _ in range(3):
                interaction.kernel_pool_bwd(q, d, qm, dm, mu, sg, w, None, tr["per_kernel_query"], gout, 1.0, **kw)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(10):
                interaction.kernel_pool_bwd(q, d, qm, dm, mu, sg, w, None, tr["per_kernel_query"], gout, 1.0, **kw)
            e1.record()
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / 10
            print("backward %s stages %s: %.3f ms per %d pairs (%.0f GB/s of 558 KB/pair)" % (name, stages, ms, B, B * 557964 / ms / 1e6), flush=True)
    os.environ.pop("MMB200_KPB_STAGES", None)
    for name, kw in (("fwd plain", {}), ("fwd train", dict(save_for_backward=True))):
        for _ in range(3):
            interaction.kernel_pool(q, d, qm, dm, mu, sg, w, **kw)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10):
            interaction.kernel_pool(q, d, qm, dm, mu, sg, w, **kw)
        e1.record()
        torch.cuda.synchronize()
        print("%s: %.3f ms per %d pairs" % (name, e0.elapsed_time(e1) / 10, B), flush=True)


if __name__ == "__main__":
    main()
