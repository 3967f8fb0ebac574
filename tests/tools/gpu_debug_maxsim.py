"""GPU bring-up script for the max-sim kernels: staged checks, each in its own subprocess with a timeout
(a watchdog trap poisons the CUDA context), then a quick timing.  Development tool, not a test.

    python tests/tools/gpu_debug_maxsim.py            # all stages
    python tests/tools/gpu_debug_maxsim.py stage N    # one stage (internal)
"""
import os
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

STAGES = [
    # n_q, dpq, Lq, Ld, dim, dtype, masks
    (1, 1, 32, 128, 64, "f16", False),
    (1, 1, 32, 128, 128, "f16", False),
    (1, 1, 32, 180, 128, "f16", True),
    (2, 3, 32, 180, 128, "f16", True),
    (3, 500, 32, 180, 128, "f16", True),
    (4, 2, 64, 57, 256, "bf16", True),
    (2, 5, 128, 129, 192, "f16", True),
    (3, 7, 17, 300, 64, "f16", True),
    (2, 4, 32, 255, 128, "bf16", True),
    (2, 4, 32, 256, 128, "f16", True),
    (2, 3, 32, 700, 128, "f16", True),
]


def run_stage(i):
    import torch
    from matchmaker_b200 import interaction
    from oracle import interaction_oracle as O
    n_q, dpq, Lq, Ld, dim, dt, masks = STAGES[i]
    dtype = {"f16": torch.float16, "bf16": torch.bfloat16}[dt]
    q, d, qm, dm = O.synth_colbert_inputs(n_q, dpq, Lq, Ld, dim, seed=100 + i, dtype=dtype, full_q=False)
    if not masks:
        qm = dm = None
    ref = O.maxsim_one_query_many_docs(q.float(), d.float(), qm, dm, dpq)
    dev = "cuda"
    args = [t.to(dev) if t is not None else None for t in (q, d, qm, dm)]
    for impl in ("simt", "tcgen05_docm", "tcgen05"):
        got = interaction.maxsim(*args, docs_per_query=dpq, impl=impl)
        torch.cuda.synchronize()
        err = (got.cpu() - ref).abs().max().item()
        rel = err / max(ref.abs().max().item(), 1e-9)
        print(f"stage {i} {STAGES[i]} {impl}: max abs err {err:.3e} rel {rel:.3e}", flush=True)
        if rel > 1e-3:
            print("   got", got.cpu()[:8].tolist())
            print("   ref", ref[:8].tolist())


def timing():
    import torch
    from matchmaker_b200 import interaction
    from oracle import interaction_oracle as O
    n_q, dpq = 64, 1000
    q, d, qm, dm = O.synth_colbert_inputs(n_q, dpq, 32, 180, 128, seed=1237)
    args = [t.cuda() for t in (q, d, qm.bool(), dm.bool())]
    bytes_per_pair = 180 * 128 * 2 + 180 + 4
    for impl in ("tcgen05", "tcgen05_docm", "simt"):
        for _ in range(3):
            interaction.maxsim(*args, docs_per_query=dpq, impl=impl)
        torch.cuda.synchronize()
        n = 20 if impl != "simt" else 3
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(n):
            interaction.maxsim(*args, docs_per_query=dpq, impl=impl)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / n
        pairs = n_q * dpq
        print(f"timing {impl}: {ms:.3f} ms / {pairs} pairs -> {pairs / ms * 1e3 / 1e6:.2f} M pairs/s, "
              f"{pairs * bytes_per_pair / ms * 1e3 / 1e9:.1f} GB/s", flush=True)


def main():
    if len(sys.argv) >= 3 and sys.argv[1] == "stage":
        run_stage(int(sys.argv[2]))
        return
    if len(sys.argv) >= 2 and sys.argv[1] == "timing":
        timing()
        return
    for i in range(len(STAGES)):
        t0 = time.time()
        try:
            r = subprocess.run([sys.executable, __file__, "stage", str(i)], timeout=180, capture_output=True, text=True)
            print(r.stdout.strip())
            if r.returncode != 0:
                print(f"stage {i} FAILED rc={r.returncode}\n{r.stderr[-2000:]}")
        except subprocess.TimeoutExpired as e:
            print(f"stage {i} TIMEOUT after {time.time() - t0:.0f}s\n{(e.stdout or b'')[-1500:]}")
    try:
        r = subprocess.run([sys.executable, __file__, "timing"], timeout=300, capture_output=True, text=True)
        print(r.stdout.strip())
        if r.returncode != 0:
            print("timing FAILED", r.stderr[-2000:])
    except subprocess.TimeoutExpired:
        print("timing TIMEOUT")


if __name__ == "__main__":
    main()
