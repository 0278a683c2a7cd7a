"""Phase timestamps inside k_tile (debug build: `make -C pc-gnn_b200/csrc trace`, run with PCG_LIB_VARIANT=trace)."""
import ctypes as C
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.abspath(os.path.join(os.path.dirname(__file__), "..")))
import bench  # noqa: E402
from pcgnn_b200 import _lib  # noqa: E402
from pcgnn_b200.synth import make_graph  # noqa: E402
from pcgnn_b200.testing import build_cuda_pcgnn  # noqa: E402

wl = sys.argv[1] if len(sys.argv) > 1 else "yelp"
spec, batch, embed, desc = bench.WORKLOADS[wl]
data = make_graph(spec, seed=72)
params = bench.init_params(data.feat.shape[1], embed, 3, 72)
model = build_cuda_pcgnn(data.feat, data.graph, sorted(data.train_pos), params, device="cuda")
batches = bench.make_batches(data, 4, batch, 72)
L = _lib.lib()
L.pcg_debug_set_tile_trace.argtypes = [C.c_void_p]
buf = torch.zeros(64, dtype=torch.int64, device="cuda")
L.pcg_debug_set_tile_trace(buf.data_ptr())
names = ["start", "dep wait", "rows loaded", "relations done", "combine done", "out/cat stored", "heads done", "dH done",
         "dh stored"]
for n, l in batches:
    buf.zero_()
    loss = model.loss(n.tolist(), torch.from_numpy(l).cuda())
    torch.cuda.synchronize()
    t = buf.cpu().numpy()
    t0 = t[0]
    print("tile 0:", ", ".join(f"{names[i]} +{(t[i] - t0) / 1e3:.2f}" for i in range(1, 9)),
          "| phase A warp 0: wait-begin +%.2f data +%.2f fma-done +%.2f scratch-written +%.2f synced +%.2f reduced +%.2f"
          % tuple((t[i] - t0) / 1e3 for i in (20, 21, 22, 23, 24, 25)))
