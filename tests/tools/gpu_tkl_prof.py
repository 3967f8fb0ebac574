"""Per-role wait profile of tkl_ts_kernel (needs a library built with -DMMB200_ENABLE_PROF: MMB200_LIB=... and
MMB200_TKL_TS_PROF=1).  Test tooling."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import bench  # noqa: E402
from matchmaker_b200 import interaction  # noqa: E402

dev = torch.device("cuda", 0)
wl = bench.TklWorkload(0, dev)
wl.to_device()
c = wl.c
for i in range(4):
    interaction.tkl_window_scores(c["q"], c["qm"], c["ch"], c["cm"], c["pk"], wl.pieces, c["mu"], c["sg"], c["dw"], "embedding",
                                  c["sat"], c["red"], impl="tcgen05")
    torch.cuda.synchronize()
