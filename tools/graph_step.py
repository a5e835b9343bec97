"""Experiment: capture the whole training step (forward + loss + backward + clip + AdamW) in one CUDA graph."""
import os, sys, time, traceback
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench
from partdistillation_b200 import compat, presets
from partdistillation_b200.engine import DataParallelTrainer

device = torch.device("cuda", 0)
torch.cuda.set_device(0)
cfg = presets.make_cfg("ProposalModel", "swin_b", bench.QUERIES, 10, bench.POINTS, 0.0, device=str(device))
torch.manual_seed(0)
model = compat.build_model(cfg); model.train()
trainer = DataParallelTrainer(model, freeze_keys=("backbone", "encoder"))
batch = bench.make_batch(0, bench.PER_GPU_BATCH, device=device)

def timed(f, n=10):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(n): f()
    torch.cuda.synchronize(); return (time.perf_counter() - t0) / n * 1e3

s = torch.cuda.Stream()
s.wait_stream(torch.cuda.current_stream())
with torch.cuda.stream(s):
    for _ in range(3):
        total, _ = trainer.step(batch)
torch.cuda.current_stream().wait_stream(s)
print("eager ms/step", timed(lambda: trainer.step(batch)))
g = torch.cuda.CUDAGraph()
try:
    with torch.cuda.graph(g):
        static_total, static_losses = trainer.step(batch)
    print("captured")
    for _ in range(3): g.replay()
    torch.cuda.synchronize()
    print("graph ms/step", timed(g.replay), "loss", float(static_total))
    g.replay(); torch.cuda.synchronize(); print("loss after more steps", float(static_total))
except Exception:
    traceback.print_exc()
