import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from matchmaker_b200 import interaction
B, Lq, Ld, D, K = 4096, 30, 200, 300, 21
import numpy as np
mu = torch.linspace(-1, 1, K).cuda(); sg = torch.full((K,), 0.1).cuda()
w = torch.linspace(-0.014, 0.014, K).cuda()
q = torch.randn(B, Lq, D, device="cuda"); d = torch.randn(B, Ld, D, device="cuda")
qm = torch.ones(B, Lq, device="cuda"); dm = torch.ones(B, Ld, device="cuda")
for _ in range(3):
    interaction.kernel_pool(q, d, qm, dm, mu, sg, w, impl="tcgen05")
torch.cuda.synchronize()
