import sys, os, cProfile, pstats, io
sys.path.insert(0, "graph-transformer_b200")
import torch, u2gnn_b200 as U
from u2gnn_b200.synthetic import make_batch
from u2gnn_b200.trainer import SupTrainer
b = make_batch(72, 8, 7, 2, avg_graph=18, seed=1, device="cuda")
torch.manual_seed(0)
m = U.TransformerU2GNN(7, 1024, 2, 3, 0.5, 1, attn_axis="nodes").cuda()
tr = SupTrainer(m, lr=5e-4, precision="fp32")
for _ in range(20): tr.step(b["input_x"], b["rowptr"], b["X"], b["labels"])
torch.cuda.synchronize()
import time
t=time.perf_counter()
for _ in range(200): tr.step(b["input_x"], b["rowptr"], b["X"], b["labels"])
torch.cuda.synchronize()
print("ms/step", (time.perf_counter()-t)/200*1e3, "launches/step", U.LIB.launches/220)
pr = cProfile.Profile(); pr.enable()
for _ in range(200): tr.step(b["input_x"], b["rowptr"], b["X"], b["labels"])
torch.cuda.synchronize()
pr.disable()
s = io.StringIO(); pstats.Stats(pr, stream=s).sort_stats("tottime").print_stats(18); print(s.getvalue()[:3500])
