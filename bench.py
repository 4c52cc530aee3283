#!/usr/bin/env python
"""Benchmark of the north-star path (BASELINE.json): train steps on synthetic 3 s @ 16 kHz utterances.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--model dccrn|fullsubnet]
                    [--perceptual PMSQE] [--batch B] [--no-extra] [--no-cpu-baseline]

One JSON line on rank 0 (contract in the task statement).  The headline workload is BASELINE configs[1]: DCCRN mask C,
SI-SNR, Adam, batch 32 per GPU.  `value` = utterances/s of K train steps (forward + loss + backward + [NCCL all-reduce] +
Adam) with inputs resident in HBM, timed with CUDA events, max over ranks.  `e2e` = the reference's own loop body
(trainer.py:27-37 / :97-112) driving the drop-in module through pinned host buffers (H2D of both waveforms and a D2H read of
the loss inside the timed region).  `roofline` = the dominant kernel category timed with CUDA events around its launches
(one extra profiled step) against the measured peaks (MEASURED_PEAKS.json; the TF32 dense rate is measured here with a
cuBLAS TF32 GEMM probe).  `cpu_baseline` = the reference's CPU path on this box's host cores (the unmodified reference from
baseline/_ref when present, else the oracle port).  `other_configs` (same line, unless --no-extra) carries the same
measurement of BASELINE configs[2] (FullSubNet, batch 64 per GPU) and configs[3] (DCCRN, (SI-SNR + PMSQE) / 2, batch 32 per GPU
= 256 over 8 GPUs).  `--impl reference` times the CPU arm alone.
"""
import argparse
import gc
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
PKG = os.path.join(ROOT, "dnn-based-speech-enhancement-in-the-frequency-domain_b200")
for p in (ROOT, PKG):
    if p not in sys.path:
        sys.path.insert(0, p)

import torch  # noqa: E402

L = 48000
# kernel categories of csrc/prof.cuh: (name, roofline that bounds it)
CATS = [("tapgemm_tc (conv/convT/linear fwd+dgrad, tcgen05 tf32)", "tensor"), ("wgrad_tc (weight gradients, tcgen05 tf32)", "tensor"),
        ("bn_prelu (fwd + 2-pass bwd)", "hbm"), ("lstm recurrence", None), ("stft / features / mask+istft / loss", "hbm"),
        ("pack / fold / reductions / dropout / adam", None), ("skinny 2-channel layers (CUDA cores)", "hbm")]
NCU_TRAFFIC = os.path.join(ROOT, "profiles", "ncu_traffic.json")   # per-kernel DRAM bytes per launch from the committed ncu capture

WORKLOADS = {
    "dccrn": dict(metric="utterances/sec DCCRN train step (3s@16kHz)", batch=32,
                  workload="DCCRN mask C, SI-SNR loss, Adam, 3s@16kHz, batch 32 per GPU (BASELINE configs[1])"),
    "dccrn_pmsqe": dict(metric="utterances/sec DCCRN SI-SNR+PMSQE train step (3s@16kHz)", batch=32,
                        workload="DCCRN mask C, (SI-SNR + PMSQE) / 2 perceptual step, Adam, 3s@16kHz, batch 32 per GPU "
                                 "(BASELINE configs[3]: 256 over 8 GPUs; PMSQE parity unpinned)"),
    "fullsubnet": dict(metric="utterances/sec FullSubNet train step (3s@16kHz)", batch=64,
                       workload="FullSubNet cIRM, MSE loss, train-mode inter-layer dropout 0.8, Adam, 3s@16kHz, batch 64 per GPU "
                                "(BASELINE configs[2])"),
}


def synthetic(B, seed, device=None, pin=False):
    g = torch.Generator().manual_seed(seed)
    noisy = (torch.rand(B, L, generator=g) * 2 - 1) * 0.1
    clean = (torch.rand(B, L, generator=g) * 2 - 1) * 0.1
    if pin:
        return noisy.pin_memory(), clean.pin_memory()
    return noisy.to(device), clean.to(device)


class ClockSampler:
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.rows, self.proc, self.index = [], None, index

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        sm = [int(r[0]) for r in self.rows if len(r) >= 6 and r[0].isdigit()]
        mx = [int(r[1]) for r in self.rows if len(r) >= 6 and r[1].isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for i, n in enumerate(names) if any(len(r) >= 6 and r[2 + i].lower().startswith("active") for r in self.rows)]
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": reasons, "samples": len(sm)}


# ---------------------------------------------------------------------------------------------------------------------
# CPU arm: the reference's own path on the host cores
# ---------------------------------------------------------------------------------------------------------------------
def cpu_step_rate(model, B, steps, warmup):
    """(utterances/s, s/step, threads, kind, description).  kind 'reference' = the UNMODIFIED reference modules from
    baseline/_ref (copied there by __graft_entry__.build() in the build container, git-ignored) running the loop body of
    trainer.model_train / fullsubnet_train with torch.optim.Adam; 'port' = the oracle restatement (oracle/)."""
    torch.set_num_threads(os.cpu_count())
    from baseline import refshim
    ref = refshim.load(model)
    if ref is not None:
        step, desc = ref.make_step(B), "unmodified reference (baseline/_ref) loop body + torch.optim.Adam"
        kind = "reference"
    elif model == "fullsubnet":
        from oracle import fullsubnet_oracle as FS
        sd = {k: v.clone().requires_grad_(True) for k, v in FS.init_state(0).items()}
        opt = torch.optim.Adam(list(sd.values()), lr=1e-3)
        noisy, clean = synthetic(B, 1234)

        def step():
            loss = FS.train_step_loss(sd, noisy, clean)
            opt.zero_grad()
            loss.backward()
            opt.step()
        desc, kind = "oracle/fullsubnet_oracle.py", "port"
    else:
        from oracle import dccrn_oracle as O
        tr = O.OracleTrainer(O.init_state(0), masking_mode="C", loss="SI-SNR")
        noisy, clean = O.synthetic_batch(B)

        def step():
            tr.step(noisy, clean)
        desc, kind = "oracle/dccrn_oracle.py", "port"
    for _ in range(warmup):
        step()
    t0 = time.perf_counter()
    for _ in range(steps):
        step()
    dt = (time.perf_counter() - t0) / steps
    return B / dt, dt, torch.get_num_threads(), kind, desc


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    model = "fullsubnet" if args.model == "fullsubnet" else "dccrn"
    B = 2 if model == "fullsubnet" else 4         # bounded sample: CPU throughput is flat in the batch size
    steps, warmup = max(1, args.steps), max(0, args.warmup)
    # keep the whole arm within a few minutes whatever K / W the driver passes: at most ~25 CPU steps in total
    if steps + warmup > 25:
        warmup = max(1, 25 - steps) if steps < 25 else 1
    rate, dt, cores, kind, desc = cpu_step_rate(model, B, steps, warmup)
    w = WORKLOADS[model]
    out = {
        "impl": "reference", "metric": w["metric"], "value": rate, "unit": "utterances/s", "n_gpus": args.gpus,
        "steps": steps, "warmup": warmup, "ms_per_step": dt * 1e3, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": w["workload"] + f"; CPU arm runs a bounded sample of {B} utterances per step"},
        "cpu_baseline": {"value": rate, "unit": "utterances/s", "cores": cores, "kind": kind,
                         "sample": f"{steps} train steps of {B} utterances after {warmup} warm-up, {desc}"},
        "e2e": {"value": rate, "unit": "utterances/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(out), flush=True)


# ---------------------------------------------------------------------------------------------------------------------
# measured peaks
# ---------------------------------------------------------------------------------------------------------------------
def tf32_probe(dev, seconds=1.5):
    """Dense TF32 rate of this GPU: cuBLAS fp32 GEMM 8192^3 with TF32 tensor cores (the same operand type our tcgen05
    kernels use), best of 10 (burst) and back to back for `seconds` (sustained, under the power cap)."""
    n = 8192
    old = torch.backends.cuda.matmul.allow_tf32
    torch.backends.cuda.matmul.allow_tf32 = True
    try:
        a = torch.randn(n, n, device=dev)
        b = torch.randn(n, n, device=dev)
        c = torch.empty(n, n, device=dev)
        for _ in range(3):
            torch.matmul(a, b, out=c)
        torch.cuda.synchronize()
        best = 1e9
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        for _ in range(10):
            e0.record()
            torch.matmul(a, b, out=c)
            e1.record()
            torch.cuda.synchronize()
            best = min(best, e0.elapsed_time(e1))
        reps = max(10, int(seconds * 1e3 / best))
        e0.record()
        for _ in range(reps):
            torch.matmul(a, b, out=c)
        e1.record()
        torch.cuda.synchronize()
        sus = e0.elapsed_time(e1) / reps
        fl = 2.0 * n ** 3 / 1e9
        return {"tf32_tflops": round(fl / best, 1), "tf32_tflops_sustained": round(fl / sus, 1),
                "how": f"torch.matmul fp32 {n}^3 with allow_tf32 (cuBLAS TF32 tensor-core GEMM): best of 10 and {reps} back to back"}
    finally:
        torch.backends.cuda.matmul.allow_tf32 = old
        del a, b, c
        torch.cuda.empty_cache()


# ---------------------------------------------------------------------------------------------------------------------
# one workload on the GPU(s)
# ---------------------------------------------------------------------------------------------------------------------
def run_workload(name, args, env, peaks, primary):
    import torch.distributed as dist
    import models
    import tools_for_model as tools
    from sefd import _lib
    from sefd.train import FlatAdam, FsnTrainStep, TrainStep
    lib = _lib.load()
    world, rank, dev = env["world"], env["rank"], env["dev"]
    w = WORKLOADS[name]
    B = args.batch if (primary and args.batch) else w["batch"]
    steps = args.steps if primary else max(3, min(args.steps, 10))
    warmup = args.warmup if primary else 3
    fsn = name == "fullsubnet"
    perceptual = "PMSQE" if name == "dccrn_pmsqe" else None
    torch.manual_seed(0)
    if fsn:
        models.cfg.loss = "MSE"
        model = models.FullSubNet().to(dev).train()
        ts = FsnTrainStep(model, lr=1e-3)
    else:
        models.cfg.loss = "SI-SNR"
        model = models.DCCRN(masking_mode="C").to(dev).train()
        ts = TrainStep(model, lr=1e-3, loss="SI-SNR", perceptual=perceptual, graph=not args.no_graph)
        models.cfg.perceptual = perceptual if perceptual else False
    noisy, clean = synthetic(B, 1234 + rank, dev)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---------------- device-resident throughput ----------------
    # the clock sampler (nvidia-smi -lms 100) needs a few hundred ms to come up: it starts before the warm-up (the same
    # step, so the same load) and runs until the end of the timed region
    sampler = ClockSampler(env["local"])
    if rank == 0:
        sampler.start()
    for _ in range(warmup):
        ts.step(noisy, clean)
    barrier()
    for _ in range(5 if fsn else 25):       # ~0.5 s of extra untimed steps on EVERY rank (same count: each holds an all-reduce)
        ts.step(noisy, clean)
    barrier()
    graphed = bool(getattr(ts, "graph", False))
    n0 = lib.sefd_launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        loss = ts.step(noisy, clean)
    e1.record()
    torch.cuda.synchronize()
    launches = lib.sefd_launch_count() - n0
    if graphed:                 # the kernels were launched by graph replays: count the nodes of one eagerly issued step
        n0 = lib.sefd_launch_count()
        ts._step_eager(noisy, clean)
        ts.steps += 1
        torch.cuda.synchronize()
        launches = (lib.sefd_launch_count() - n0) * steps
    ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
    barrier()
    clocks = sampler.stop() if rank == 0 else None
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    ms = float(ms)
    value = world * B * steps / (ms / 1e3)
    final_loss = float(loss)

    # ---------------- end to end through the drop-in module (reference loop body) ----------------
    opt = FlatAdam(model, lr=1e-3)
    h_noisy, h_clean = synthetic(B, 99 + rank, pin=True)
    # host -> device: every step's inputs are copied from pinned host memory inside the timed region, double-buffered on a
    # copy stream the way sefd.feed.WaveFeeder stages batches (the copy for step i+1 is issued when step i starts and
    # overlaps it; the loss read at the end of each step is the hand-over point that frees the other slot)
    copy_stream = torch.cuda.Stream(device=dev)
    slots = [(torch.empty(B, L, device=dev), torch.empty(B, L, device=dev)) for _ in range(2)]
    ready = [torch.cuda.Event() for _ in range(2)]
    state = {"i": 0}

    def stage(slot):
        with torch.cuda.stream(copy_stream):
            slots[slot][0].copy_(h_noisy, non_blocking=True)
            slots[slot][1].copy_(h_clean, non_blocking=True)
            ready[slot].record(copy_stream)

    def e2e_step():
        slot = state["i"] & 1
        state["i"] += 1
        stage(slot ^ 1)                        # next step's batch, in flight during this step
        torch.cuda.current_stream().wait_event(ready[slot])
        inputs, targets = slots[slot]
        if fsn:                                # trainer.fullsubnet_train, trainer.py:97-112
            noisy_complex = tools.stft(inputs)
            clean_complex = tools.stft(targets)
            noisy_mag, _ = tools.mag_phase(noisy_complex)
            cirm = tools.build_complex_ideal_ratio_mask(noisy_complex, clean_complex)
            lo = model.loss(cirm, model(noisy_mag))
        else:                                  # trainer.model_train / model_perceptual_train, trainer.py:27-37, 59-67
            real_spec, img_spec, outputs = model(inputs, targets)
            lo = model.loss(outputs, targets)
            if perceptual:
                lo = (lo + model.loss(outputs, targets, real_spec, img_spec, perceptual=True)) / 2
        opt.zero_grad()
        lo.backward()
        opt.step()
        return lo.item()                       # D2H read of the loss

    stage(0)
    for _ in range(2):
        e2e_step()
    barrier()
    e0.record()
    for _ in range(steps):
        e2e_step()
    e1.record()
    torch.cuda.synchronize()
    ms2 = torch.tensor([e0.elapsed_time(e1)], device=dev)
    barrier()
    if world > 1:
        dist.all_reduce(ms2, op=dist.ReduceOp.MAX)
    e2e_value = world * B * steps / (float(ms2) / 1e3)

    # ---------------- per-kernel-category CUDA-event timing (one extra profiled step) ----------------
    roofline, breakdown = None, None
    # every rank runs the extra step (it contains the gradient all-reduce); only rank 0 records events
    if rank == 0:
        lib.sefd_prof_reset()
        lib.sefd_prof_enable(1)
    if graphed:                      # the profiled step is issued kernel by kernel (events around every launch)
        ts._step_eager(noisy, clean)
        ts.steps += 1
    else:
        ts.step(noisy, clean)
    torch.cuda.synchronize()
    lib.sefd_prof_enable(0)
    if rank == 0:
        import ctypes as C
        if primary and os.path.isdir(os.path.join(ROOT, "gpurun_out")):
            lib.sefd_prof_dump(os.path.join(ROOT, "gpurun_out", "launch_profile.csv").encode())
        hbm_peak, tf32_peak, tf32_burst, src = peaks["hbm"], peaks["tf32"], peaks["tf32_burst"], peaks["src"]
        try:
            traffic = json.load(open(NCU_TRAFFIC))
        except OSError:
            traffic = {}
        breakdown = {}
        bounds = dict(CATS)
        bounds["lstm recurrence"] = "tensor" if fsn else None     # FullSubNet: fused tcgen05 step GEMMs; DCCRN: latency-bound
        for c, (cname, _) in enumerate(CATS):
            bound = bounds[cname]
            t, n, f, b = C.c_double(), C.c_longlong(), C.c_double(), C.c_double()
            lib.sefd_prof_get(c, C.byref(t), C.byref(n), C.byref(f), C.byref(b))
            if n.value == 0:
                continue
            e = {"ms": round(t.value, 3), "launches": n.value, "gflop": round(f.value / 1e9, 1),
                 "gbyte": round(b.value / 1e9, 3)}
            if t.value > 0 and bound == "tensor":
                e["tflops"] = round(f.value / 1e9 / t.value, 1)
                e["frac_of_tf32_peak"] = round(e["tflops"] / tf32_peak, 3)
            elif t.value > 0 and bound == "hbm":
                e["gbs"] = round(b.value / 1e6 / t.value, 1)
                e["frac_of_hbm_peak"] = round(e["gbs"] / hbm_peak, 3)
            breakdown[cname] = e
        lib.sefd_prof_reset()
        top = max(breakdown, key=lambda k: breakdown[k]["ms"])
        bt = breakdown[top]
        nl = max(bt["launches"], 1)
        tkey = "tapgemm_tc_kernel" if top.startswith("tapgemm_tc") else ("wgrad_tc_kernel" if top.startswith("wgrad_tc") else
                                                                       ("lstm_" if top.startswith("lstm") and fsn else None))
        tr = None
        if tkey:                      # launch-weighted mean over the template instances of the kernel
            inst = [v for k, v in traffic.items() if k.startswith(tkey)]
            if inst:
                nn_ = sum(v["launches"] for v in inst)
                tr = {"dram_bytes_per_launch": round(sum(v["dram_bytes_per_launch"] * v["launches"] for v in inst) / nn_),
                      "source": f"profiles/ncu_traffic.json ({nn_} launches of one step; {inst[0]['source']})"}
        if bounds[top] == "tensor":
            ach = bt["gflop"] / bt["ms"]                      # GFLOP / ms = TFLOP/s
            roofline = {"kernel": top, "bound": "tensor", "achieved": round(ach, 2), "peak": round(tf32_peak, 1),
                        "unit": "TFLOP/s", "frac": round(ach / tf32_peak, 4), "frac_of_burst_peak": round(ach / tf32_burst, 4),
                        "traffic": tr["dram_bytes_per_launch"] if tr else None,
                        "algorithmic_bytes_per_launch": round(bt["gbyte"] * 1e9 / nl),
                        "algorithmic_flops_per_launch": round(bt["gflop"] * 1e9 / nl),
                        "peak_source": src, "traffic_source": tr["source"] if tr else None,
                        "avg_launch_ms": round(bt["ms"] / nl, 4), "launches": nl}
        else:
            ach = bt["gbyte"] / bt["ms"] * 1e3
            roofline = {"kernel": top, "bound": "hbm", "achieved": round(ach, 1), "peak": hbm_peak, "unit": "GB/s",
                        "frac": round(ach / hbm_peak, 4), "traffic": None, "peak_source": src,
                        "avg_launch_ms": round(bt["ms"] / nl, 4), "launches": nl}

    res = {
        "metric": w["metric"], "value": round(value, 2), "unit": "utterances/s", "n_gpus": world, "steps": steps,
        "warmup": warmup, "ms_per_step": round(ms / steps, 3), "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": w["workload"], "batch_per_gpu": B, "global_batch": B * world, "samples": L, "parallelism": f"dp{world}",
                   "l2": "per-step working set (GBs of activations) >> 126 MB L2, no flush needed",
                   "precision": "fp32 storage, TF32 tensor-core operands (round-to-nearest at the producers), fp32 accumulation",
                   "final_loss": round(final_loss, 4)},
        "clocks": clocks,
        "e2e": {"value": round(e2e_value, 2), "unit": "utterances/s", "h2d_bytes_per_step": 2 * B * L * 4,
                "d2h_bytes_per_step": 4,
                "api": ("tools.stft x2 + mag_phase + build_complex_ideal_ratio_mask + models.FullSubNet + model.loss + backward + "
                        "sefd.train.FlatAdam") if fsn else "models.DCCRN + model.loss + backward + sefd.train.FlatAdam",
                "h2d": "pinned -> device every step, double-buffered on a copy stream (overlaps the previous step)"},
        "gpu_launches": int(launches),
        "launch_mode": ("one CUDA graph replay per step (gpu_launches = kernel nodes per graph x steps)" if graphed
                        else "every kernel launched from the host"),
        "roofline": roofline,
        "kernel_breakdown_ms": breakdown,
    }
    del ts, opt, model, noisy, clean, slots
    gc.collect()
    torch.cuda.empty_cache()
    return res


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--model", default="dccrn", choices=["dccrn", "fullsubnet"],
                    help="headline workload of the line: dccrn = BASELINE configs[1] (default), fullsubnet = configs[2]")
    ap.add_argument("--batch", type=int, default=0, help="utterances per GPU (default: the config's own, 32 / 64)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-extra", action="store_true", help="skip the other BASELINE configs (other_configs)")
    ap.add_argument("--no-graph", action="store_true", help="launch every kernel of the DCCRN step from the host instead of replaying one CUDA graph")
    ap.add_argument("--perceptual", default=None, choices=["PMSQE"],
                    help="headline = the perceptual train step (SI-SNR + PMSQE) / 2 of BASELINE configs[3] (per-GPU slice of 32)")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)
    args.warmup = max(args.warmup, 3)

    import torch.distributed as dist
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        import datetime
        dist.init_process_group("nccl", device_id=dev, timeout=datetime.timedelta(seconds=300))
    env = {"world": world, "rank": rank, "local": local, "dev": dev}

    mp = {}
    try:
        mp = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except OSError:
        pass
    probe = tf32_probe(dev) if rank == 0 else {"tf32_tflops": 1.0, "tf32_tflops_sustained": 1.0, "how": ""}
    peaks = {"hbm": mp.get("hbm_gbs", 6650.0), "tf32": probe["tf32_tflops_sustained"], "tf32_burst": probe["tf32_tflops"],
             "src": ("HBM: " + ("measured (MEASURED_PEAKS.json)" if mp else "fallback (B200_PROFILING.md)") +
                     "; TF32 dense: measured in this run, " + probe["how"] + f" -> {probe['tf32_tflops_sustained']} sustained / "
                     f"{probe['tf32_tflops']} burst TFLOP/s (bf16 in MEASURED_PEAKS.json: {mp.get('bf16_tflops_sustained')} / {mp.get('bf16_tflops')})")}
    if world > 1:
        dist.barrier()

    head = "fullsubnet" if args.model == "fullsubnet" else ("dccrn_pmsqe" if args.perceptual else "dccrn")
    out = run_workload(head, args, env, peaks, primary=True)
    extras = {}
    if not args.no_extra:
        for name in ("fullsubnet", "dccrn_pmsqe", "dccrn"):
            if name == head:
                continue
            if name == "dccrn" and head != "fullsubnet":
                continue
            r = run_workload(name, args, env, peaks, primary=False)
            extras[WORKLOADS[name]["workload"]] = {k: r[k] for k in ("metric", "value", "unit", "n_gpus", "steps", "ms_per_step", "e2e",
                                                                     "gpu_launches", "roofline", "kernel_breakdown_ms", "config")}
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    if rank != 0:
        return
    cpu = None
    if not args.no_cpu_baseline and world == 1:
        model = "fullsubnet" if head == "fullsubnet" else "dccrn"
        Bc = 2 if model == "fullsubnet" else 4
        rate, dt, cores, kind, desc = cpu_step_rate(model, Bc, 2, 1)
        cpu = {"value": round(rate, 3), "unit": "utterances/s", "cores": cores, "kind": kind,
               "sample": f"2 train steps of {Bc} utterances (3 s each) after 1 warm-up, {desc}"}
    out["cpu_baseline"] = cpu
    out["measured_peaks"] = {"hbm_gbs": peaks["hbm"], "tf32_tflops_sustained": peaks["tf32"], "tf32_tflops_burst": peaks["tf32_burst"]}
    if extras:
        out["other_configs"] = extras
    print(json.dumps(out), flush=True)


if __name__ == "__main__":
    main()
