#!/usr/bin/env python
"""Benchmark of the north-star path: DCCRN (mask C) SI-SNR train step on synthetic 3 s @ 16 kHz utterances.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--batch 32]

One JSON line on rank 0 (contract in the task statement).  `value` = utterances/s of K train steps
(forward + loss + backward + [NCCL all-reduce] + Adam) with inputs resident in HBM, timed with CUDA events,
max over ranks.  `e2e` = the reference's own loop body (trainer.py:27-37) driving the drop-in models.DCCRN
through pinned host buffers (H2D of both waveforms and a D2H read of the loss inside the timed region).
`roofline` = the dominant kernel category timed with CUDA events around its launches (a separate profiled
step).  `cpu_baseline` = the CPU oracle port of the reference step timed on this box's host cores.
`--impl reference` times that CPU port as the reference arm.
"""
import argparse
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
METRIC = "utterances/sec DCCRN train step (3s@16kHz)"
# kernel categories of csrc/prof.cuh: (name, roofline that bounds it)
CATS = [("tapgemm_tc (conv/convT/linear fwd+dgrad, tcgen05 tf32)", "tensor"), ("wgrad_tc (weight gradients, tcgen05 tf32)", "tensor"),
        ("bn_prelu (fwd + 2-pass bwd)", "hbm"), ("lstm recurrence (latency-bound)", "hbm"), ("stft / mask+istft / loss", "hbm"),
        ("pack / fold / reductions / adam", None), ("skinny 2-channel layers (CUDA cores)", "hbm")]
NCU_TRAFFIC = os.path.join(ROOT, "profiles", "ncu_traffic.json")   # per-kernel DRAM bytes per launch from the committed ncu capture


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


def cpu_step_rate(B, steps, warmup):
    """The CPU port of the reference train step (oracle/dccrn_oracle.py), all host threads."""
    from oracle import dccrn_oracle as O
    torch.set_num_threads(os.cpu_count())
    tr = O.OracleTrainer(O.init_state(0), masking_mode="C", loss="SI-SNR")
    noisy, clean = O.synthetic_batch(B)
    for _ in range(warmup):
        tr.step(noisy, clean)
    t0 = time.perf_counter()
    for _ in range(steps):
        tr.step(noisy, clean)
    dt = (time.perf_counter() - t0) / steps
    return B / dt, dt, torch.get_num_threads()


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    B = 4
    rate, dt, cores = cpu_step_rate(B, max(1, args.steps), min(args.warmup, 1))
    out = {
        "impl": "reference", "metric": METRIC, "value": rate, "unit": "utterances/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": min(args.warmup, 1), "ms_per_step": dt * 1e3, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": "DCCRN mask C, SI-SNR, 3s@16kHz (BASELINE configs[1]); CPU arm runs a bounded "
                               f"sample of {B} utterances per step"},
        "cpu_baseline": {"value": rate, "unit": "utterances/s", "cores": cores, "kind": "port",
                         "sample": f"{args.steps} train steps of {B} utterances, oracle/dccrn_oracle.py"},
        "e2e": {"value": rate, "unit": "utterances/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(out), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--batch", type=int, default=32, help="utterances per GPU (BASELINE configs[1]: 32)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--perceptual", default=None, choices=["PMSQE"],
                    help="time the perceptual train step (SI-SNR + PMSQE) / 2 of BASELINE configs[3] (per-GPU slice of 32) instead")
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
        dist.init_process_group("nccl", device_id=dev, timeout=datetime.timedelta(seconds=180))

    import models
    from sefd import _lib
    from sefd.train import FlatAdam, TrainStep
    lib = _lib.load()
    models.cfg.loss = "SI-SNR"
    torch.manual_seed(0)
    model = models.DCCRN(masking_mode="C").to(dev).train()
    B = args.batch
    noisy, clean = synthetic(B, 1234 + rank, dev)
    ts = TrainStep(model, lr=1e-3, loss="SI-SNR", perceptual=args.perceptual)
    if args.perceptual:
        models.cfg.perceptual = args.perceptual

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---------------- device-resident throughput ----------------
    # the clock sampler (nvidia-smi -lms 100) needs a few hundred ms to come up: it starts before the warm-up (the same
    # step, so the same load) and runs until the end of the timed region
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    for _ in range(args.warmup):
        ts.step(noisy, clean)
    barrier()
    for _ in range(25):                 # ~0.5 s of extra untimed steps on EVERY rank (same count: each holds an all-reduce)
        ts.step(noisy, clean)
    barrier()
    n0 = lib.sefd_launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        loss = ts.step(noisy, clean)
    e1.record()
    torch.cuda.synchronize()
    launches = lib.sefd_launch_count() - n0
    ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
    barrier()
    clocks = sampler.stop() if rank == 0 else None
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    ms = float(ms)
    value = world * B * args.steps / (ms / 1e3)
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
        real_spec, img_spec, outputs = model(inputs, targets)
        lo = model.loss(outputs, targets)
        if args.perceptual:                    # trainer.model_perceptual_train, trainer.py:59-67 (r1 = r2 = 1)
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
    for _ in range(args.steps):
        e2e_step()
    e1.record()
    torch.cuda.synchronize()
    ms2 = torch.tensor([e0.elapsed_time(e1)], device=dev)
    barrier()
    if world > 1:
        dist.all_reduce(ms2, op=dist.ReduceOp.MAX)
    e2e_value = world * B * args.steps / (float(ms2) / 1e3)

    # ---------------- per-kernel-category CUDA-event timing (one extra profiled step) ----------------
    roofline, breakdown = None, None
    # every rank runs the extra step (it contains the gradient all-reduce); only rank 0 records events
    if rank == 0:
        lib.sefd_prof_reset()
        lib.sefd_prof_enable(1)
    ts.step(noisy, clean)
    torch.cuda.synchronize()
    lib.sefd_prof_enable(0)
    if rank == 0:
        import ctypes as C
        if os.path.isdir(os.path.join(ROOT, "gpurun_out")):
            lib.sefd_prof_dump(os.path.join(ROOT, "gpurun_out", "launch_profile.csv").encode())
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except OSError:
            pass
        hbm_peak = peaks.get("hbm_gbs", 6650.0)
        bf16_peak = peaks.get("bf16_tflops_sustained", peaks.get("bf16_tflops", 1590.0))
        src = "measured (MEASURED_PEAKS.json)" if peaks else "fallback (B200_PROFILING.md)"
        tf32_peak = bf16_peak / 2.0                         # TF32 dense = 1/2 of the bf16 tensor rate
        try:
            traffic = json.load(open(NCU_TRAFFIC))
        except OSError:
            traffic = {}
        breakdown = {}
        for c, (name, bound) in enumerate(CATS):
            t, n, f, b = C.c_double(), C.c_longlong(), C.c_double(), C.c_double()
            lib.sefd_prof_get(c, C.byref(t), C.byref(n), C.byref(f), C.byref(b))
            e = {"ms": round(t.value, 3), "launches": n.value, "gflop": round(f.value / 1e9, 1),
                 "gbyte": round(b.value / 1e9, 3)}
            if t.value > 0 and bound == "tensor":
                e["tflops"] = round(f.value / 1e9 / t.value, 1)
                e["frac_of_tf32_peak"] = round(e["tflops"] / tf32_peak, 3)
            elif t.value > 0 and bound == "hbm":
                e["gbs"] = round(b.value / 1e6 / t.value, 1)
                e["frac_of_hbm_peak"] = round(e["gbs"] / hbm_peak, 3)
            breakdown[name] = e
        lib.sefd_prof_reset()
        top = max(breakdown, key=lambda k: breakdown[k]["ms"])
        bt = breakdown[top]
        nl = max(bt["launches"], 1)
        tkey = "tapgemm_tc_kernel" if top.startswith("tapgemm_tc") else ("wgrad_tc_kernel" if top.startswith("wgrad_tc") else None)
        tr = None
        if tkey:                      # launch-weighted mean over the template instances of the kernel
            inst = [v for k, v in traffic.items() if k.startswith(tkey)]
            if inst:
                nn_ = sum(v["launches"] for v in inst)
                tr = {"dram_bytes_per_launch": round(sum(v["dram_bytes_per_launch"] * v["launches"] for v in inst) / nn_),
                      "source": f"profiles/ncu_traffic.json ({nn_} launches of one step; {inst[0]['source']})"}
        if dict(CATS)[top] == "tensor":
            ach = bt["gflop"] / bt["ms"]                      # GFLOP / ms = TFLOP/s
            roofline = {"kernel": top, "bound": "tensor", "achieved": round(ach, 2), "peak": round(tf32_peak, 1),
                        "unit": "TFLOP/s", "frac": round(ach / tf32_peak, 4),
                        "traffic": tr["dram_bytes_per_launch"] if tr else None,
                        "algorithmic_bytes_per_launch": round(bt["gbyte"] * 1e9 / nl),
                        "algorithmic_flops_per_launch": round(bt["gflop"] * 1e9 / nl),
                        "peak_source": f"{src}: bf16 sustained / 2 (tcgen05 kind::tf32 dense rate)",
                        "traffic_source": tr["source"] if tr else None,
                        "avg_launch_ms": round(bt["ms"] / nl, 4), "launches": nl}
        else:
            ach = bt["gbyte"] / bt["ms"] * 1e3
            roofline = {"kernel": top, "bound": "hbm", "achieved": round(ach, 1), "peak": hbm_peak, "unit": "GB/s",
                        "frac": round(ach / hbm_peak, 4), "traffic": None, "peak_source": src,
                        "avg_launch_ms": round(bt["ms"] / nl, 4), "launches": nl}

    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    if rank != 0:
        return

    cpu = None
    if not args.no_cpu_baseline and world == 1 and not args.perceptual:
        rate, dt, cores = cpu_step_rate(4, 2, 1)
        cpu = {"value": round(rate, 3), "unit": "utterances/s", "cores": cores, "kind": "port",
               "sample": "2 train steps of 4 utterances (3 s each) after 1 warm-up, oracle/dccrn_oracle.py"}

    out = {
        "metric": METRIC, "value": round(value, 2), "unit": "utterances/s", "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": round(ms / args.steps, 3), "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": ("DCCRN mask C, (SI-SNR + PMSQE) / 2 perceptual step, Adam, 3s@16kHz, batch 32 per GPU (BASELINE "
                                "configs[3] slice; PMSQE parity unpinned)") if args.perceptual else
                               "DCCRN mask C, SI-SNR loss, Adam, 3s@16kHz, batch 32 per GPU (BASELINE configs[1])",
                   "batch_per_gpu": B, "global_batch": B * world, "samples": L, "parallelism": f"dp{world}",
                   "l2": "per-step working set (~6 GB of activations) >> 126 MB L2, no flush needed",
                   "final_loss": round(final_loss, 4)},
        "clocks": clocks,
        "e2e": {"value": round(e2e_value, 2), "unit": "utterances/s", "h2d_bytes_per_step": 2 * B * L * 4,
                "d2h_bytes_per_step": 4, "api": "models.DCCRN + model.loss + backward + sefd.train.FlatAdam",
                "h2d": "pinned -> device every step, double-buffered on a copy stream (overlaps the previous step)"},
        "gpu_launches": int(launches),
        "roofline": roofline,
        "kernel_breakdown_ms": breakdown,
        "cpu_baseline": cpu,
    }
    print(json.dumps(out), flush=True)


if __name__ == "__main__":
    main()
