"""Where one training step spends its time: wall clock vs summed kernel time (torch.profiler / CUPTI, warm
caches, real overlap) and the top kernels.  Answers "is the step launch-bound or GPU-bound?".
Usage: python tools/step_profile.py [out.txt]"""
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402


def main(out_path, workload="c2"):
    import contextlib
    from partdistillation_b200 import compat, presets
    from partdistillation_b200.engine import DataParallelTrainer
    device = torch.device("cuda", 0)
    torch.cuda.set_device(0)
    spec = bench.SPECS[workload]
    cfg = presets.make_cfg(spec["arch"], "swin_b", bench.QUERIES, 10, bench.POINTS, 0.0, device=str(device), **spec["cfg"])
    torch.manual_seed(0)
    model = compat.build_model(cfg)
    model.train()
    real = DataParallelTrainer(model, freeze_keys=("backbone", "encoder"))
    batch = bench.make_batch(0, spec["batch"], device=device, spec=spec)
    amp = (lambda: torch.autocast("cuda", dtype=torch.bfloat16)) if spec["amp"] else contextlib.nullcontext

    class T:                                    # the trainer's step under the workload's autocast context
        def step(self, b):
            with amp():
                return real.step(b)
    trainer = T()
    for _ in range(3):
        trainer.step(batch)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    n = 5
    for _ in range(n):
        trainer.step(batch)
    torch.cuda.synchronize()
    wall = (time.perf_counter() - t0) / n * 1e3
    # stage split with events
    from torch.profiler import ProfilerActivity, profile
    stacks = bool(os.environ.get("PROFILE_STACK"))
    with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA], record_shapes=True, with_stack=stacks) as prof:
        for _ in range(2):
            trainer.step(batch)
        torch.cuda.synchronize()
    ev = prof.key_averages()
    rows = []
    total = 0.0
    for e in ev:
        t = getattr(e, "self_device_time_total", 0.0) or 0.0
        if t > 0 and e.device_type == torch.autograd.DeviceType.CUDA:
            rows.append((t / 2, e.count // 2, e.key))
            total += t / 2
    rows.sort(reverse=True)
    with open(out_path, "w") as w:
        w.write(f"# one training step ({spec['workload']}): wall {wall:.2f} ms/step; summed kernel time "
                f"{total / 1e3:.2f} ms/step over {sum(r[1] for r in rows)} launches (torch.profiler, warm)\n")
        w.write(f"{'us/step':>10} {'share':>7} {'count':>6}  kernel\n")
        for t, c, k in rows[:70]:
            w.write(f"{t:10.1f} {100 * t / total:6.1f}% {c:6d}  {k[:110]}\n")
        w.write("\n# ATen / autograd ops by device time (self + children), with input shapes\n")
        ops = []
        for e in prof.key_averages(group_by_input_shape=True):
            t = getattr(e, "device_time_total", 0.0) or 0.0
            if t > 0 and e.device_type == torch.autograd.DeviceType.CPU:
                ops.append((t / 2, e.count // 2, e.key, str(e.input_shapes)[:150]))
        ops.sort(reverse=True)
        for t, c, k, sh in ops[:60]:
            w.write(f"{t:10.1f} {c:6d}  {k[:40]:40s} {sh}\n")
        w.write("\n# ops by SELF device time (kernels the op launched itself), all shapes together\n")
        selfs = []
        for e in prof.key_averages():
            t = getattr(e, "self_device_time_total", 0.0) or 0.0
            if t > 0 and e.device_type == torch.autograd.DeviceType.CPU:
                selfs.append((t / 2, e.count // 2, e.key))
        selfs.sort(reverse=True)
        for t, c, k in selfs[:50]:
            w.write(f"{t:10.1f} {c:6d}  {k[:80]}\n")
        if stacks:
            w.write("\n# python call sites of the element-wise / copy ATen ops (PROFILE_STACK=1)\n")
            glue = ("aten::copy_", "aten::add", "aten::add_", "aten::cat", "aten::div", "aten::mul", "aten::fill_", "aten::zero_",
                    "aten::sum", "aten::index", "aten::gather", "aten::clone", "aten::sub", "aten::where", "aten::addcmul")
            sites = []
            for e in prof.key_averages(group_by_stack_n=12):
                t = getattr(e, "self_device_time_total", 0.0) or 0.0
                if t > 0 and e.key in glue:
                    here = [f for f in e.stack if "/partdistillation_b200/" in f or "/bench.py" in f][:3]
                    sites.append((t / 2, e.count // 2, e.key, " <- ".join(h.split("/partdistillation_b200/")[-1] for h in here)))
            sites.sort(reverse=True)
            for t, c, k, st in sites[:70]:
                w.write(f"{t:10.1f} {c:6d}  {k:14s} {st[:260]}\n")
    print(open(out_path).read()[:6000])


if __name__ == "__main__":
    os.makedirs("gpurun_out", exist_ok=True)
    main(sys.argv[1] if len(sys.argv) > 1 else "gpurun_out/step_profile.txt", sys.argv[2] if len(sys.argv) > 2 else "c2")
