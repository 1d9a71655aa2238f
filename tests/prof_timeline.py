#!/usr/bin/env python3
"""Kernel timeline of the benchmark training step (CUPTI through torch.profiler; not a timing source for bench numbers).

    python tests/prof_timeline.py [--steps 3] [--out gpurun_out/timeline.json]

Answers what the per-kernel CUDA-event sums cannot: how much of a step the GPU is IDLE between kernels on the main stream
(launch gaps, prologue/drain), how much side-stream work actually overlaps, and the span of one step.  Output: a JSON
summary (per stream: busy ms, kernel count; main-stream gap histogram; per-kernel-name totals) of the last profiled step."""
import argparse
import contextlib
import io
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "pytorch-kaldi-resnet_b200")
for p in (ROOT, PKG, os.path.join(PKG, "scripts")):
    if p not in sys.path:
        sys.path.insert(0, p)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--batch", type=int, default=256)
    ap.add_argument("--out", default=os.path.join(ROOT, "gpurun_out", "timeline.json"))
    ap.add_argument("--graph", action="store_true", help="profile the CUDA-graph replay of the step instead of eager launches")
    args = ap.parse_args()
    import torch
    from torch.profiler import ProfilerActivity, profile
    from model import NeuralSpeakerModel
    from svk.loss import CrossEntropyLoss
    from svk.optim import SGD
    torch.manual_seed(1234)
    with contextlib.redirect_stdout(io.StringIO()):
        net = NeuralSpeakerModel(spk_num=5994, feat_dim=40, pooling="mean+std", loss="AAM").cuda()
    crit = CrossEntropyLoss()
    opt = SGD(net.parameters(), 0.1, momentum=0.9, weight_decay=5e-4)
    x = torch.randn(args.batch, 40, 200, device="cuda")
    y = torch.randint(0, 5994, (args.batch,), device="cuda")
    net.train()

    def step():
        loss, _ = net.forward_loss(x, y)
        opt.zero_grad()
        loss.backward()
        opt.step()
    if args.graph:
        from svk.graph import GraphedTrainStep
        gs = GraphedTrainStep(net, crit, opt)
        step = lambda: gs(x, y)      # noqa: E731
    for _ in range(5):
        step()
    torch.cuda.synchronize()
    with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
        for _ in range(args.steps):
            step()
            torch.cuda.synchronize()
    trace = args.out + ".trace.json"
    os.makedirs(os.path.dirname(args.out), exist_ok=True)
    prof.export_chrome_trace(trace)
    with open(trace) as f:
        ev = json.load(f)["traceEvents"]
    os.remove(trace)
    ks = [(float(e["ts"]), float(e["ts"]) + float(e["dur"]), e["name"], e.get("args", {}).get("stream", 0))
          for e in ev if e.get("cat") == "kernel" and "dur" in e]
    ks.sort()
    # split into steps at the synchronize boundaries: the largest (steps - 1) idle intervals
    run_end, idle = ks[0][1], []
    for i in range(1, len(ks)):
        if ks[i][0] > run_end:
            idle.append((ks[i][0] - run_end, i))
        run_end = max(run_end, ks[i][1])
    cuts = sorted(i for _, i in sorted(idle, reverse=True)[:args.steps - 1])
    last = ks[cuts[-1] if cuts else 0:]
    t0, t1 = min(a for a, _, _, _ in last), max(b for _, b, _, _ in last)
    by_stream, by_name = {}, {}
    for a, b, name, sid in last:
        by_stream.setdefault(sid, []).append((a, b, name))
        n = by_name.setdefault(name.split("<")[0].split("(")[0], [0.0, 0])
        n[0] += (b - a) / 1e3
        n[1] += 1
    out = {"step_span_ms": (t1 - t0) / 1e3, "kernels": len(last), "streams": {}}
    union = []
    for sid, evs in by_stream.items():
        evs.sort()
        busy = sum(b - a for a, b, _ in evs) / 1e3
        g = [evs[i + 1][0] - evs[i][1] for i in range(len(evs) - 1)]
        gpos = [v for v in g if v > 0]
        out["streams"][str(sid)] = {"kernels": len(evs), "busy_ms": busy, "gap_sum_ms": sum(gpos) / 1e3,
                                    "gap_median_us": (sorted(gpos)[len(gpos) // 2] if gpos else 0.0),
                                    "gaps_over_20us": sum(1 for v in gpos if v > 20)}
        union += [(a, b) for a, b, _ in evs]
    union.sort()
    idle, cur_end = 0.0, union[0][1]
    for a, b in union[1:]:
        if a > cur_end:
            idle += a - cur_end
        cur_end = max(cur_end, b)
    out["gpu_idle_ms_all_streams"] = idle / 1e3
    main_sid = max(by_stream, key=lambda k: len(by_stream[k]))
    seq, prev_end = [], None
    for a, b, name in by_stream[main_sid]:
        seq.append([name.split("<")[0].split("(")[0][-40:], round(b - a, 1), round(a - prev_end, 1) if prev_end is not None else 0.0])
        prev_end = b
    out["main_stream_sequence_us"] = seq          # [kernel, duration, gap before it]
    out["by_kernel_ms"] = {k: [round(v[0], 4), v[1]] for k, v in sorted(by_name.items(), key=lambda kv: -kv[1][0])[:40]}
    with open(args.out, "w") as f:
        json.dump(out, f, indent=1)
    print(json.dumps({k: v for k, v in out.items() if k not in ("by_kernel_ms", "main_stream_sequence_us")}, indent=1))


if __name__ == "__main__":
    main()
