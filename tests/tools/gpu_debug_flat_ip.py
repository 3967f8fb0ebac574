"""Bring-up of the flat-IP top-k kernel: staged, each stage in a subprocess with a timeout."""
import os, subprocess, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
STAGES = [(7, 3000, 64, 10), (130, 70000, 128, 100), (64, 20000, 768, 100), (200, 5000, 64, 256)]

def stage(i):
    import torch
    from matchmaker_b200 import interaction
    from oracle import interaction_oracle as O
    nq, n, dim, k = STAGES[i]
    q, p = O.synth_dense_inputs(nq, n, dim, seed=nq + n)
    s, ids = interaction.flat_ip_topk(q.cuda(), p.cuda(), k)
    torch.cuda.synchronize()
    rs, ri = O.flat_ip_search(q.float(), p, torch.arange(n), k)
    print(f"stage {STAGES[i]}: id match {(ids.cpu() == ri).float().mean().item():.4f} max score err {(s.cpu() - rs).abs().max().item():.3e}", flush=True)

def timing():
    import torch
    from matchmaker_b200 import interaction
    from oracle import interaction_oracle as O
    for nq, n in [(6400, 1100000), (1024, 1100000), (128, 1100000)]:
        q, p = O.synth_dense_inputs(nq, n, 768, seed=1)
        cq, cp = q.cuda(), p.cuda()
        for _ in range(2): interaction.flat_ip_topk(cq, cp, 100)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(3): interaction.flat_ip_topk(cq, cp, 100)
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 3
        print(f"timing nq={nq} n={n} dim=768 k=100: {ms:.2f} ms -> {nq * n / ms * 1e3 / 1e9:.1f} G pairs/s, {2 * nq * n * 768 / ms * 1e3 / 1e12:.1f} TFLOP/s", flush=True)

if __name__ == "__main__":
    if len(sys.argv) > 2 and sys.argv[1] == "stage": stage(int(sys.argv[2])); sys.exit(0)
    if len(sys.argv) > 1 and sys.argv[1] == "timing": timing(); sys.exit(0)
    if len(sys.argv) > 1 and sys.argv[1] == "timing1":
        import torch
        from matchmaker_b200 import interaction, synthetic as O
        q, p = O.synth_dense_inputs(6400, 1100000, 768, seed=1)
        cq, cp = q.cuda(), p.cuda()
        for _ in range(3): interaction.flat_ip_topk(cq, cp, 100)
        torch.cuda.synchronize(); sys.exit(0)
    for i in range(len(STAGES)):
        try:
            r = subprocess.run([sys.executable, __file__, "stage", str(i)], timeout=200, capture_output=True, text=True)
            print(r.stdout.strip()); 
            if r.returncode: print("FAILED", r.stderr[-1500:])
        except subprocess.TimeoutExpired: print("stage", i, "TIMEOUT")
    try:
        r = subprocess.run([sys.executable, __file__, "timing"], timeout=400, capture_output=True, text=True)
        print(r.stdout.strip())
        if r.returncode: print("timing FAILED", r.stderr[-1500:])
    except subprocess.TimeoutExpired: print("timing TIMEOUT")
