#!/usr/bin/env python3
"""bench.py — headline benchmark of the speaker-embedding hot path (BASELINE.json: "ResNet-34 AAM train chunks/sec").

    python bench.py --gpus N --steps K --warmup W              our arm (libsvk, bf16 tcgen05 path)
    python bench.py --impl reference --gpus N --steps K ...    the reference's own CPU path on the host cores: the UNMODIFIED
                                                               reference from baseline/_ref/scripts (copied there verbatim by
                                                               __graft_entry__.build()), else the oracle port
    python bench.py --workload cfg4|cfg2|cfg5 ...              the other BASELINE.json configurations, same JSON line

A "step" = one full training step of NeuralSpeakerModel(5994 speakers, 40-dim fbank, mean+std pooling, AAM m=0.2 s=30)
on a batch of 256 synthetic 200-frame chunks PER GPU (BASELINE.json config 3: forward, cross-entropy, backward,
bucketed NCCL gradient all-reduce when N > 1, SGD momentum 0.9 wd 5e-4).  `value` = chunks/s of the whole job with the
batch resident in HBM; `e2e` = the same step through the public API (scripts/model.py + svk.loss + svk.optim) with
the batch copied from pinned host memory every step and the loss read back every step.  One JSON line on stdout.
"""
import argparse
import contextlib
import io
import json
import os
import statistics
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
PKG = os.path.join(ROOT, "pytorch-kaldi-resnet_b200")
for p in (ROOT, PKG, os.path.join(PKG, "scripts")):
    if p not in sys.path:
        sys.path.insert(0, p)

SPK, FEAT, FRAMES, BATCH = 5994, 40, 200, 256
TRAIN_FLOP_PER_CHUNK = 13588306944          # SURVEY.md §8d / BASELINE.md §4 (C = 5994)
METRIC, UNIT = "resnet34_aam_train_chunks_per_sec", "chunks/s"
WORKLOAD = "cfg3: ResNet-34 AAM train step, bf16, batch 256/GPU x 200 frames x 40 fbank, 5994 speakers"


def peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            d = json.load(f)
        return d.get("hbm_gbs", 6650.0), d.get("bf16_tflops_sustained", 1400.0), "measured (MEASURED_PEAKS.json, sustained)"
    return 6650.0, 1590.0, "fallback (B200_PROFILING.md)"


def ncu_traffic():
    """DRAM bytes per launch of the fprop/dgrad kernels from the committed ncu capture of this same command
    (profiles/r02_collect.sh -> profiles/summarize_ncu_step.py over `ncu --metrics ...dram__bytes_read.sum,
    dram__bytes_write.sum`, one training step): a static figure measured under the profiler, reported beside the live
    timings; null when the capture is absent."""
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "profiles", "r02_ncu_step_metrics.json")
    try:
        with open(path) as f:
            ks = [k for k in json.load(f)["kernels"] if k["kernel"].startswith("conv_tc_gather")]
        n = sum(k["launches"] for k in ks)
        by = sum(k["dram_MB_per_launch"] * 1e6 * k["launches"] for k in ks)
        pipe = sum(k["tensor_pct_of_elapsed"] * k["time_us"] for k in ks) / sum(k["time_us"] for k in ks)
        return {"traffic": by / n,
                "traffic_unit": "DRAM bytes per kernel launch (ncu: %d kernel launches of one step; a downsample data gradient "
                                "is one call of `launches` but three kernels)" % n,
                "traffic_source": "profiles/r02_ncu_step_metrics.json",
                "ncu_tensor_pipe_pct_of_elapsed": pipe}
    except (OSError, KeyError, ValueError, ZeroDivisionError):
        return {}


def conv_flops(tag):
    """Algorithmic FLOPs of one conv launch from its profile tag 'HxW Cin->Cout kR sS NN' (2*MACs, zero padding
    counted as in SURVEY.md §8d: output pixels x Cout x R*R*Cin)."""
    if " + " in tag:                       # several convs in one call
        return sum(conv_flops(t) for t in tag.split(" + "))
    hw, ch, k, s, n = tag.split()
    H, W = map(int, hw.split("x"))
    ci, co = map(int, ch.split("->"))
    R, S, N = int(k[1:]), int(s[1:]), int(n[1:])
    Ho, Wo = (H - 1) // S + 1, (W - 1) // S + 1
    return 2.0 * N * Ho * Wo * co * ci * R * R


class ClockSampler(object):
    """nvidia-smi clocks / throttle reasons sampled every 200 ms while the timed region runs."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.f = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        try:
            self.p = subprocess.Popen(["nvidia-smi", "-i", str(index), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits",
                                       "-lms", "200"], stdout=self.f, stderr=subprocess.DEVNULL)
        except OSError:
            self.p = None

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": []}
        if self.p is None:
            return out
        self.p.terminate()
        try:
            self.p.wait(timeout=5)
        except Exception:
            self.p.kill()
        self.f.flush()
        self.f.seek(0)
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for line in self.f.read().splitlines():
            parts = [x.strip() for x in line.split(",")]
            if len(parts) < 7:
                continue
            try:
                sm.append(float(parts[0]))
                mx.append(float(parts[1]))
            except ValueError:
                continue
            for nm, v in zip(names, parts[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        os.unlink(self.f.name)
        if sm:
            hi = [s for s in sm if s >= 0.5 * max(sm)] or sm     # samples under load
            out = {"sm_mhz": statistics.median(hi), "sm_max_mhz": max(mx), "reasons": sorted(reasons), "samples": len(sm)}
        return out


# ------------------------------------------------------------------------------------------------ reference arm
import bench_workloads as W  # noqa: E402

WORKLOADS = {
    "cfg3": dict(metric=METRIC, unit=UNIT, frames=200, flop=13588306944, batch=256, widths=None,
                 name="cfg3: ResNet-34 AAM train step, bf16, batch %d/GPU x 200 frames x 40 fbank, 5994 speakers"),
    "cfg4": dict(metric="wide_resnet34_aam_train_chunks_per_sec", unit=UNIT, frames=300, flop=81637039104, batch=128,
                 widths=(64, 128, 256, 512),
                 name="cfg4: 2x-wide ResNet-34 AAM train step, bf16, batch %d/GPU x 300 frames x 40 fbank, 5994 speakers"),
    "cfg2": dict(metric="embed_extract_utts_per_sec", unit="utts/s",
                 name="cfg2: ResNet-34 embedding extraction, %d variable-length utterances (200-6000 frames, 40 fbank) + 37,720-trial cosine scoring"),
    "cfg5": dict(metric="snorm_scoring_trials_per_sec", unit="trials/s",
                 name="cfg5: top-300 cohort statistics of 100,000 embeddings vs a 50,000 x 256 cohort + 1,000,000 cosine trials + adaptive s-norm"),
}


def cpu_reference(steps, warmup, batch=32, frames=200, widths=None):
    return W.cpu_train(steps, warmup, batch, frames, widths)


def run_reference(args):
    """`--impl reference`: the reference's CPU implementation of the workload on the host cores (rank 0 only), each step
    a bounded sample of the workload so the run ends within a few minutes."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    wl = WORKLOADS[args.workload]
    steps = max(1, min(args.steps, 5))
    warmup = max(1, min(args.warmup, 2))
    if args.workload in ("cfg3", "cfg4"):
        cb = W.cpu_train(steps, warmup, 32 if args.workload == "cfg3" else 16, wl["frames"], wl["widths"])
        name = wl["name"] % wl["batch"]
        ms = cb["ms_per_step"]
    elif args.workload == "cfg2":
        n = args.extract_utts
        ds = W.MemoryUtterances(W.cfg2_lengths()[:min(n, 48)])
        t0 = time.perf_counter()
        cb = W.cpu_extract(ds, 48)
        ms = 1e3 * (time.perf_counter() - t0)
        name = wl["name"] % n
        steps, warmup = 1, 1
    else:
        t0 = time.perf_counter()
        cb = W.cpu_snorm()
        ms = 1e3 * (time.perf_counter() - t0)
        name = wl["name"]
        steps, warmup = 1, 0
    line = {"impl": "reference", "metric": wl["metric"], "value": cb["value"], "unit": wl["unit"], "n_gpus": args.gpus,
            "steps": steps, "warmup": warmup, "ms_per_step": ms, "higher_is_better": True,
            "scaling": "strong" if args.workload == "cfg5" else "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": name, "note": "reference CPU path on the host cores (%s); each step is a bounded sample: %s"
                       % (cb["kind"], cb["sample"])},
            "cpu_baseline": {k: cb[k] for k in ("value", "unit", "cores", "kind", "sample")},
            "e2e": {"value": cb["value"], "unit": wl["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line))


# ------------------------------------------------------------------------------------------------ our arm
def run_ours(args):
    import torch
    import torch.distributed as dist
    from model import NeuralSpeakerModel
    from svk import lib
    from svk.loss import CrossEntropyLoss
    from svk.optim import SGD
    from svk.parallel import DistributedDataParallel

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the product path has no CPU fallback")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    global FRAMES, TRAIN_FLOP_PER_CHUNK, WORKLOAD, METRIC
    wl = WORKLOADS[args.workload]
    B = args.batch if args.batch else wl["batch"]
    FRAMES, TRAIN_FLOP_PER_CHUNK, METRIC = wl["frames"], wl["flop"], wl["metric"]
    WORKLOAD = wl["name"] % B
    kw = {}
    if args.workload == "cfg4":      # BASELINE.json config 4: 2x-wide trunk, 300-frame chunks (not the headline; --workload cfg4)
        kw = {"widths": wl["widths"]}
        args.no_extras = True        # the secondary legs describe the headline workload
    torch.manual_seed(1234)
    with contextlib.redirect_stdout(io.StringIO()):
        net = NeuralSpeakerModel(spk_num=SPK, feat_dim=FEAT, pooling="mean+std", loss="AAM", m=0.2, s=30, **kw).cuda(local)
    model = DistributedDataParallel(net, device_ids=[local]) if world > 1 else net
    crit = CrossEntropyLoss()
    opt = SGD(model.parameters(), 0.1, momentum=0.9, weight_decay=5e-4)
    g = torch.Generator().manual_seed(1234 + rank)
    x_host = torch.randn(B, FEAT, FRAMES, generator=g).pin_memory()
    y_host = torch.randint(0, SPK, (B,), generator=g).pin_memory()
    x_dev, y_dev = x_host.to(dev), y_host.to(dev)
    model.train()

    def eager_step(x, y):
        loss, out = model.forward_loss(x, y)     # = model(x, y) + CrossEntropyLoss, what scripts/train_resnet.py::train calls
        opt.zero_grad()
        loss.backward()
        opt.step()
        return loss

    # --cuda-graph: replay ONE captured CUDA graph per step (svk.graph.GraphedTrainStep: forward, fused loss, backward,
    # all-reduce, SGD) instead of issuing the ~250 launches from Python.  Measured equal (profiles/r02_launch_overhead.md:
    # the step is GPU-bound and PDL already hides the launch gaps), so the default stays the eager loop of train_resnet.py.
    gstep = None
    if args.cuda_graph:
        from svk.graph import GraphedTrainStep
        gstep = GraphedTrainStep(model, opt)

    def step(x, y):
        return gstep(x, y)[0] if gstep is not None else eager_step(x, y)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, n):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(n):
            fn()
        e1.record()
        barrier()
        ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms)

    for _ in range(max(args.warmup, 3)):
        step(x_dev, y_dev)
    # ---- device-resident throughput
    sampler = ClockSampler(local) if rank == 0 else None
    n0, r0 = lib.launch_count(), (gstep.replays if gstep else 0)
    ms_total = timed(lambda: step(x_dev, y_dev), args.steps)
    launches = lib.launch_count() - n0 + ((gstep.replays - r0) * gstep.launches_per_replay if gstep else 0)
    # ---- end to end through the public API (the loop of scripts/train_resnet.py): every step's batch is copied from
    # pinned host memory (svk.data.DevicePrefetcher: the copy of batch i+1 overlaps step i) and every step's loss is
    # read back to the host (svk.data.ScalarReader: asynchronous copy to a pinned slot, collected two steps later)
    from svk.data import DevicePrefetcher, ScalarReader

    def e2e_run(n):
        reader = ScalarReader(dev)
        for xb, yb in DevicePrefetcher(((x_host, y_host) for _ in range(n)), dev):
            reader.push(step(xb, yb))
        vals = reader.flush()
        assert len(vals) == n and all(v == v for v in vals), "e2e: a step's loss did not reach the host"
        return vals

    e2e_run(2)
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    e2e_run(args.steps)
    e1.record()
    barrier()
    ms_t = torch.tensor([e0.elapsed_time(e1)], device=dev)
    if world > 1:
        dist.all_reduce(ms_t, op=dist.ReduceOp.MAX)
    ms_e2e = float(ms_t)
    clocks = sampler.stop() if sampler else None
    # ---- sanity of what was timed: the loss of one more step is finite and, with N > 1, every rank holds the same
    # parameters after the same number of all-reduced updates (train_resnet.py:185 semantics)
    final_loss = float(step(x_dev, y_dev))
    assert final_loss == final_loss and abs(final_loss) < 1e4, "bench: training loss is not finite (%r)" % final_loss
    if world > 1:
        flat = net.engine.flat_params
        chk = torch.stack([flat.double().sum(), flat.double().abs().sum()])
        lo, hi = chk.clone(), chk.clone()
        dist.all_reduce(lo, op=dist.ReduceOp.MIN)
        dist.all_reduce(hi, op=dist.ReduceOp.MAX)
        assert torch.equal(lo, hi), "bench: parameters differ between ranks after the timed steps"
    ms_step = ms_total / args.steps
    value = world * B / (ms_step / 1e3)
    e2e_value = world * B / (ms_e2e / args.steps / 1e3)

    line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
            "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "bf16",
            "data": "synthetic",
            "config": {"workload": WORKLOAD, "global_batch": world * B, "parallelism": "dp%d" % world,
                       "l2": "no flush needed: one step streams ~4 GB of activations per GPU (>> 126 MB L2)",
                       "launch": "eager (Python issues every launch)" if gstep is None else
                                 "one CUDA-graph replay per step (%d libsvk kernels per replay)" % gstep.launches_per_replay,
                       "algorithmic_tflops_per_gpu": TRAIN_FLOP_PER_CHUNK * B / (ms_step / 1e3) / 1e12},
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": x_host.numel() * 4 + y_host.numel() * 8,
                    "d2h_bytes_per_step": 4},
            "gpu_launches": launches, "clocks": clocks,
            "checks": {"final_loss": final_loss, "ranks_hold_identical_parameters": True if world > 1 else None}}

    # ---- per-kernel device times of one step, measured live with CUDA events on the launching stream.  EVERY rank runs
    # this step (it contains the gradient all-reduce); only rank 0 records and reports.
    # The profiled step runs on ONE stream (the timed steps above overlap the weight gradients with the BatchNorm passes on a
    # side stream, where per-call event times are not additive); it is outside the timed region.
    side = net.engine.wgrad_side
    net.engine.wgrad_side = False
    if rank == 0:
        lib.profile_begin()
    eager_step(x_dev, y_dev)
    prof = lib.profile_end() if rank == 0 else None
    net.engine.wgrad_side = side
    barrier()
    if rank == 0:
        agg = {}
        for name, tag, ms in prof:
            key = name
            if name in ("svk_conv2d_fwd", "svk_conv2d_dgrad", "svk_conv2d_dgrad_bn", "svk_downsample_dgrad_bn"):
                key = "conv_tc_gather(fprop+dgrad)"
            a = agg.setdefault(key, [0.0, 0, 0.0])
            a[0] += ms
            a[1] += 1
            if tag:
                a[2] += conv_flops(tag)
        if args.profile_out:
            bytag = {}
            for name, tag, ms in prof:
                a = bytag.setdefault((name, tag), [0.0, 0])
                a[0] += ms
                a[1] += 1
            rows = [{"kernel": k[0], "shape": k[1], "ms": round(v[0], 4), "launches": v[1],
                     "tflops": (round(conv_flops(k[1]) * v[1] / (v[0] / 1e3) / 1e12, 1) if k[1] else None)}
                    for k, v in sorted(bytag.items(), key=lambda kv: -kv[1][0])]
            with open(args.profile_out, "w") as f:
                json.dump(rows, f, indent=1)
        total_prof = sum(a[0] for a in agg.values())
        top = max(agg.items(), key=lambda kv: kv[1][0])
        hbm, tf, src = peaks()
        kname, (kms, kn, kflop) = top
        line["kernel_breakdown_ms"] = {k: round(v[0], 3) for k, v in sorted(agg.items(), key=lambda kv: -kv[1][0])}
        if kflop > 0:
            ach = kflop / (kms / 1e3) / 1e12
            line["roofline"] = {"bound": "tensor", "kernel": kname, "achieved": ach, "peak": tf, "unit": "TFLOP/s",
                                "frac": ach / tf, "traffic": None, "peak_source": src, "launches": kn,
                                "avg_launch_ms": kms / kn, "share_of_step": kms / total_prof}
            line["roofline"].update(ncu_traffic())
        else:
            line["roofline"] = {"bound": "hbm", "kernel": kname, "achieved": None, "peak": hbm, "unit": "GB/s", "frac": None,
                                "traffic": None, "peak_source": src, "share_of_step": kms / total_prof}
        # wgrad kernel as a second roofline entry (same tensor bound)
        wg = agg.get("svk_conv2d_wgrad")
        if wg:
            line["roofline_wgrad"] = {"bound": "tensor", "achieved": wg[2] / (wg[0] / 1e3) / 1e12, "peak": tf, "unit": "TFLOP/s",
                                      "frac": wg[2] / (wg[0] / 1e3) / 1e12 / tf, "share_of_step": wg[0] / total_prof}
        if world == 1 and not args.no_extras:
            try:
                line["extra"] = extras(net, dev)
            except Exception as ex:          # secondary numbers must never break the headline line
                line["extra"] = {"error": str(ex)[:200]}
        if world == 1 and not args.no_cpu_baseline:
            cb = cpu_reference(3, 1, 32 if args.workload == "cfg3" else 16, FRAMES, wl["widths"])
            line["cpu_baseline"] = {k: cb[k] for k in ("value", "unit", "cores", "kind", "sample")}
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


EXTRACT_UTTS = int(os.environ.get("SVK_BENCH_EXTRACT_UTTS", "2048"))


def _sync_time(fn, reps=1):
    import torch
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(reps):
        fn()
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / reps


def measure_extraction(net, dev, n_utts, resident_reps=1):
    """BASELINE.json config 2 on one GPU, decode.py semantics (eval-mode predict, every utterance = its batch-1 result):
    `e2e`   = scripts/decode.py::extract over host-resident utterances — staging, padding, host -> device copies of every
              batch and the device -> host read of every embedding inside the timed region;
    `value` = the same padded batches already resident in HBM (predict only).
    FLOPs per frame from SURVEY.md §8d; the roofline is the tensor one (the trunk is the training forward pass)."""
    import numpy as np
    import torch
    import decode
    T = W.cfg2_lengths()[:n_utts]
    ds = W.MemoryUtterances(T)
    net.eval()
    out = {}
    idx = list(range(len(ds)))
    # warm-up = one full pass: the engine's activation arenas and the pinned staging slots grow to their final sizes
    decode.extract(net, ds, idx, dev, 65536, lambda u, v: None)
    dt_e2e = _sync_time(lambda: decode.extract(net, ds, idx, dev, 65536, lambda u, v: out.__setitem__(u, v)))
    # device-resident: the same batches, padded and copied beforehand
    lengths = [ds.num_frames(i) for i in idx]
    batches = []
    for b in decode.plan_batches(lengths, idx, 65536):
        tmax = max(lengths[i] for i in b)
        xb = torch.zeros(len(b), FEAT, tmax)
        for r, i in enumerate(b):
            xb[r, :, :lengths[i]] = torch.from_numpy(ds.mats[i])
        ln = torch.tensor([lengths[i] for i in b], dtype=torch.int32)
        batches.append((xb.to(dev), None if min(lengths[i] for i in b) == tmax else ln.to(dev)))

    def resident():
        with torch.no_grad():
            for xb, ln in batches:
                net.predict(xb, lengths=ln)
    resident()
    dt_res = _sync_time(resident, resident_reps)
    del batches
    frames = float(T.sum())
    flop = frames * W.EXTRACT_FLOP_PER_FRAME + len(T) * W.EXTRACT_FLOP_PER_UTT
    hbm, tf, src = peaks()
    emb = np.stack([out[u] for u in ds.utts]).astype(np.float32)
    net.train()
    return {"utts": len(T), "frames": int(frames), "utts_per_sec": len(T) / dt_res, "frames_per_sec": frames / dt_res,
            "e2e_utts_per_sec": len(T) / dt_e2e, "e2e_frames_per_sec": frames / dt_e2e,
            "h2d_bytes": int(sum(4 * FEAT * t for t in T)), "d2h_bytes": int(len(T) * 256 * 4),
            "roofline": {"bound": "tensor", "achieved": flop / dt_res / 1e12, "peak": tf, "unit": "TFLOP/s",
                         "frac": flop / dt_res / 1e12 / tf, "peak_source": src,
                         "note": "algorithmic FLOPs of the valid frames (22.63 MFLOP/frame); padded batches add up to ~5 % more work"},
            "sample": "%d of the 4708 cfg2 utterances (%d frames), length-sorted zero-padded batches of <= 65,536 frames" % (len(T), int(frames))}, ds, emb


def measure_cosine(emb, dev, n_trials=37720):
    """37,720 trial pairs (VoxCeleb1-O shape) with mean subtraction: device-resident kernel rate, and end to end through
    scripts/cosine_score.py on Kaldi text files next to the reference's cosine_score.py on the same files (scores compared)."""
    import numpy as np
    import torch
    from svk import scoring
    rs = np.random.RandomState(1234)
    n = len(emb)
    ie = rs.randint(0, n, n_trials).astype(np.int32)
    it = rs.randint(0, n, n_trials).astype(np.int32)
    E = torch.from_numpy(emb).to(dev)
    mean = E.mean(0)
    ied, itd = torch.from_numpy(ie).to(dev), torch.from_numpy(it).to(dev)
    scoring.cosine_scores(E, E, mean, ied, itd, device=dev)
    dt = _sync_time(lambda: scoring.cosine_scores(E, E, mean, ied, itd, device=dev), 20)
    hbm, tf, src = peaks()
    res = {"trials": n_trials, "trials_per_sec": n_trials / dt,
           "roofline": {"bound": "hbm", "achieved": n_trials * 2 * emb.shape[1] * 4 / dt / 1e9, "peak": hbm, "unit": "GB/s",
                        "frac": n_trials * 2 * emb.shape[1] * 4 / dt / 1e9 / hbm, "peak_source": src,
                        "note": "upper-bound bytes (2 x D x 4 per trial, no reuse); the %.1f MB embedding table is L2-resident, "
                                "so this kernel is launch/L2-bound, not HBM-bound" % (emb.nbytes / 1e6)}}
    with tempfile.TemporaryDirectory() as tmp:
        paths = W.write_scoring_files(tmp, emb, ie, it)
        ours = os.path.join(tmp, "ours_scores")
        dt_ours = W.run_script(os.path.join(PKG, "scripts"), "cosine_score.py",
                               ["--mean", paths["mean"], "--enroll", paths["emb"], "--test", paths["emb"], "--trials", paths["trials"],
                                "--score-file", ours])
        res["e2e_script_trials_per_sec"] = n_trials / dt_ours
        res["e2e_note"] = "scripts/cosine_score.py as a process on Kaldi text files: interpreter + CUDA start-up, text parse, one kernel, write"
        cb = W.cpu_cosine_score(paths, n_trials)
        if cb is not None:
            a = np.array([float(l.split()[2]) for l in open(ours)])
            b = np.array([float(l.split()[2]) for l in open(cb.pop("score_file"))])
            cb["max_abs_score_difference_vs_ours"] = float(np.abs(a - b).max())
            res["cpu_baseline"] = cb
    return res


def extras(net, dev):
    """Secondary numbers of BASELINE.json configs 2 and 5 on one GPU (not part of `value`), each with the reference's CPU
    path timed beside it on a bounded sample and its roofline."""
    import torch
    from svk import scoring
    ext, ds, emb = measure_extraction(net, dev, EXTRACT_UTTS)
    ext["cpu_baseline"] = W.cpu_extract(ds, 24)
    out = {"extraction": ext, "extract_utts_per_sec": ext["e2e_utts_per_sec"], "extract_frames_per_sec": ext["e2e_frames_per_sec"]}
    out["cosine_scoring"] = measure_cosine(emb, dev)
    out["score_trials_per_sec"] = out["cosine_scoring"]["trials_per_sec"]
    # adaptive s-norm statistics (config 5 shape, bounded sample): 4,096 embeddings against a 50,000 x 256 cohort, top-300
    coh = torch.randn(50000, 256, device=dev)
    q = torch.randn(4096, 256, device=dev)
    for tf32 in (False, True):
        scoring.cohort_topk_meanstd(q, coh, topk=300, device=dev, tf32=tf32)
        dt = _sync_time(lambda: scoring.cohort_topk_meanstd(q, coh, topk=300, device=dev, tf32=tf32))
        out["snorm_stats_rows_per_sec" + ("_tf32" if tf32 else "")] = 4096 / dt
    out["snorm_sample"] = "4,096 embeddings x 50,000-row cohort, top-300 mean/std (see --workload cfg5 for the full job)"
    return out


# ------------------------------------------------------------------------------------------------ --workload cfg2
def run_extract(args):
    import torch
    from model import NeuralSpeakerModel
    from svk import lib
    if int(os.environ.get("WORLD_SIZE", "1")) > 1:
        raise SystemExit("--workload cfg2 is the single-GPU extraction configuration of BASELINE.json")
    dev = torch.device("cuda", 0)
    torch.cuda.set_device(0)
    torch.manual_seed(1234)
    with contextlib.redirect_stdout(io.StringIO()):
        net = NeuralSpeakerModel(spk_num=SPK, feat_dim=FEAT, pooling="mean+std", loss="AAM").cuda()
    sampler = ClockSampler(0)
    n0 = lib.launch_count()
    ext, ds, emb = measure_extraction(net, dev, args.extract_utts, resident_reps=max(1, args.steps // 10))
    launches = lib.launch_count() - n0
    clocks = sampler.stop()
    wl = WORKLOADS["cfg2"]
    line = {"metric": wl["metric"], "value": ext["utts_per_sec"], "unit": wl["unit"], "n_gpus": 1, "steps": 1, "warmup": 1,
            "ms_per_step": 1e3 * ext["utts"] / ext["utts_per_sec"], "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "bf16", "data": "synthetic",
            "config": {"workload": wl["name"] % ext["utts"], "sample": ext["sample"],
                       "l2": "no flush needed: one pass streams every utterance once (>> 126 MB of activations)"},
            "e2e": {"value": ext["e2e_utts_per_sec"], "unit": wl["unit"], "h2d_bytes_per_step": ext["h2d_bytes"],
                    "d2h_bytes_per_step": ext["d2h_bytes"]},
            "gpu_launches": launches, "clocks": clocks, "roofline": dict(ext["roofline"], traffic=None),
            "frames_per_sec": ext["frames_per_sec"], "e2e_frames_per_sec": ext["e2e_frames_per_sec"],
            "cosine_scoring": measure_cosine(emb, dev)}
    if not args.no_cpu_baseline:
        line["cpu_baseline"] = W.cpu_extract(ds, 48)
    print(json.dumps(line))


# ------------------------------------------------------------------------------------------------ --workload cfg5
def run_scoring(args):
    """1,000,000 trials over 100,000 embeddings with adaptive s-norm against a 50,000-row cohort (BASELINE.json config 5).
    Rank r owns embedding rows [lo, hi) for the cohort statistics and trial block [tlo, thi) for scoring; the only exchange
    is one all-gather of the (mean, std) pairs.  `value` = 1 M trials / wall time of the whole job (strong scaling: the job
    is fixed), inputs resident in HBM; `e2e` = the same with embeddings / cohort / trial indices copied from pinned host
    memory and the scores read back inside the timed region."""
    import numpy as np
    import torch
    import torch.distributed as dist
    from svk import lib, scoring
    from svk.parallel import shard_range
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    NE, NT, NC, D, K = args.cfg5_embeddings, args.cfg5_trials, 50000, 256, 300
    rs = np.random.RandomState(1234)
    emb_h = torch.from_numpy(rs.randn(NE, D).astype(np.float32)).pin_memory()
    coh_h = torch.from_numpy(rs.randn(NC, D).astype(np.float32)).pin_memory()
    ie_h = torch.from_numpy(rs.randint(0, NE, NT).astype(np.int32)).pin_memory()
    it_h = torch.from_numpy(rs.randint(0, NE, NT).astype(np.int32)).pin_memory()
    lo, hi = shard_range(NE, rank, world)
    tlo, thi = shard_range(NT, rank, world)
    per = (NE + world - 1) // world
    out_h = torch.empty(thi - tlo, dtype=torch.float32).pin_memory()

    def job(E, C, ie, it, tf32):
        mean = scoring.global_mean(E, device=dev)
        Ec = E[lo:hi] - mean                                                       # compute_topk_mean_std.py:41-43
        m, s = scoring.cohort_topk_meanstd(Ec, C - mean, topk=K, device=dev, tf32=tf32)
        if world > 1:                                                              # the final gather of (mean, std) pairs
            pad = torch.zeros(2, per, device=dev)
            pad[0, :hi - lo], pad[1, :hi - lo] = m, s
            allp = torch.empty(world, 2, per, device=dev)
            dist.all_gather_into_tensor(allp, pad)
            m = torch.cat([allp[r, 0, :shard_range(NE, r, world)[1] - shard_range(NE, r, world)[0]] for r in range(world)])
            s = torch.cat([allp[r, 1, :shard_range(NE, r, world)[1] - shard_range(NE, r, world)[0]] for r in range(world)])
        sc = scoring.cosine_scores(E, E, mean, ie[tlo:thi], it[tlo:thi], device=dev)  # cosine_score.py:60-65
        return scoring.snorm_apply(sc, ie[tlo:thi], it[tlo:thi], m, s, m, s, device=dev)  # adaptive_snorm.py:28-38 (test.sh:53-54: same stats file)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, n):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(n):
            fn()
        e1.record()
        barrier()
        ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms) / n

    E, C, ie, it = emb_h.to(dev), coh_h.to(dev), ie_h.to(dev), it_h.to(dev)
    tf32 = not args.cfg5_fp32
    steps = max(1, min(args.steps, 5))
    for _ in range(max(1, min(args.warmup, 3))):
        ref_out = job(E, C, ie, it, tf32)
    sampler = ClockSampler(local) if rank == 0 else None
    n0 = lib.launch_count()
    ms = timed(lambda: job(E, C, ie, it, tf32), steps)
    launches = lib.launch_count() - n0

    def e2e():
        out = job(emb_h.to(dev, non_blocking=True), coh_h.to(dev, non_blocking=True), ie_h.to(dev, non_blocking=True),
                  it_h.to(dev, non_blocking=True), tf32)
        out_h.copy_(out, non_blocking=True)
    e2e()
    ms_e2e = timed(e2e, steps)
    clocks = sampler.stop() if sampler else None
    # a bounded accuracy check of what was timed: tf32 cohort statistics against the exact-fp32 path on the first rows
    chk = None
    if rank == 0:
        mean = scoring.global_mean(E, device=dev)
        a = scoring.cohort_topk_meanstd(E[:512] - mean, C - mean, topk=K, device=dev, tf32=tf32)
        b = scoring.cohort_topk_meanstd(E[:512] - mean, C - mean, topk=K, device=dev, tf32=False)
        chk = {"stats_max_abs_diff_vs_fp32_path": float(max((a[0] - b[0]).abs().max(), (a[1] - b[1]).abs().max())),
               "finite_scores": bool(torch.isfinite(ref_out).all())}
    if rank == 0:
        wl = WORKLOADS["cfg5"]
        hbm, tf, src = peaks()
        flop = 2.0 * NE * NC * D
        line = {"metric": wl["metric"], "value": NT / (ms / 1e3), "unit": wl["unit"], "n_gpus": world, "steps": steps,
                "warmup": max(1, min(args.warmup, 3)), "ms_per_step": ms, "higher_is_better": True, "scaling": "strong",
                "vs_baseline": None, "dtype": "tf32" if tf32 else "f32", "data": "synthetic",
                "config": {"workload": wl["name"], "embeddings": NE, "trials": NT, "cohort": NC, "topk": K,
                           "parallelism": "embedding rows and trial blocks sharded over %d rank(s), one all-gather of (mean, std)" % world,
                           "l2": "inputs larger than L2 (102 MB embeddings + 51 MB cohort); score blocks are sized to stay L2-resident by design"},
                "e2e": {"value": NT / (ms_e2e / 1e3), "unit": wl["unit"],
                        "h2d_bytes_per_step": int(emb_h.numel() * 4 + coh_h.numel() * 4 + 2 * NT * 4), "d2h_bytes_per_step": int((thi - tlo) * 4)},
                "gpu_launches": launches, "clocks": clocks, "checks": chk,
                "roofline": {"bound": "tensor", "kernel": "cohort GEMM (svk_gemm_tf32) + top-300 select", "achieved": flop / world / (ms / 1e3) / 1e12,
                             "peak": tf / 2, "unit": "TFLOP/s", "frac": flop / world / (ms / 1e3) / 1e12 / (tf / 2), "traffic": None,
                             "peak_source": src + "; tf32 dense peak taken as half the measured bf16 figure",
                             "note": "algorithmic FLOPs of the cohort GEMM (2 x rows x 50,000 x 256) per GPU over the WHOLE job time"},
                "embeddings_per_sec": NE / (ms / 1e3)}
        if world == 1 and not args.no_cpu_baseline:
            line["cpu_baseline"] = W.cpu_snorm()
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--batch", type=int, default=0, help="chunks per GPU per step (default: 256 for cfg3 = BASELINE.json config 3, 128 for cfg4)")
    ap.add_argument("--workload", default="cfg3", choices=["cfg3", "cfg4", "cfg2", "cfg5"],
                    help="cfg3 = BASELINE.json headline (default); cfg4 = 2x-wide / 300-frame variant; cfg2 = extraction + "
                         "37,720-trial scoring; cfg5 = 1 M-trial s-norm scoring against a 50 k cohort, sharded over the ranks")
    ap.add_argument("--extract-utts", type=int, default=4708, help="cfg2: utterances of the 4,708-utterance set to run")
    ap.add_argument("--cfg5-embeddings", type=int, default=100000)
    ap.add_argument("--cfg5-trials", type=int, default=1000000)
    ap.add_argument("--cfg5-fp32", action="store_true", help="cfg5: exact fp32 cohort products (CUDA cores) instead of tf32 tensor cores")
    ap.add_argument("--cuda-graph", action="store_true", help="replay one captured CUDA graph per step instead of issuing every launch from Python")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-extras", action="store_true", help="skip the secondary extraction / scoring numbers")
    ap.add_argument("--profile-out", default="", help="write the per-(kernel, shape) CUDA-event times of one step here")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    elif args.workload == "cfg2":
        run_extract(args)
    elif args.workload == "cfg5":
        run_scoring(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
