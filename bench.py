#!/usr/bin/env python
"""bench.py — headline benchmark of the hot path: CAE shape-space training throughput in volumes/s.

Workload at every N: BASELINE.json configs[1] — CAE `1 16 24 32 100 200 1`, synthetic 1 x 28 x 128 x 128 core /
penumbra / follow-up masks, batch 8 PER GPU (weak scaling), fp32, one full training step = 3 encoder + 4 decoder
passes, CaeReconstructionLearner.loss_step (epoch 60 -> ramp factor 1), backward, gradient all-reduce (N > 1), fused
Adam.  A "volume" is one patient sample (one batch element).

  value  : whole-job volumes/s with the masks already resident in HBM (CUDA events, max over ranks)
  e2e    : the same step through the public API `CaeReconstructionLearner.train_batch(batch, epoch)` with a HOST
           (pinned) batch: H2D of labels + clinical and the D2H loss read are inside the timed region
  roofline / cpu_baseline / clocks / gpu_launches : see DESIGN.md §Measurement

`--impl reference` times the reference algorithm's CPU path (the oracle port, all host threads) on a bounded sample.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
for _p in (ROOT, os.path.join(ROOT, "oracle")):
    if _p not in sys.path:
        sys.path.insert(0, _p)

import torch  # noqa: E402

CHANNELS = {"cae200": [1, 16, 24, 32, 100, 200, 1], "cae800": [1, 16, 24, 32, 100, 800, 1],
            "unet": [2, 16, 32, 64, 32, 16, 32, 2]}
SIZE = (28, 128, 128)
EPOCH = 60            # ramp factor f = 1 (CaeReconstructionLearner.py:53)
UNIT = "volumes/s"
FFMA_PEAK_TFLOPS = 148 * 128 * 2 * 1.965e9 / 1e12     # 148 SMs x 128 FP32 lanes x 2 flop x max SM clock (no measured figure)


def metric_name(workload):
    return "unet_train_volumes_per_s" if workload == "unet" else "cae_train_volumes_per_s"


def default_batch(workload):
    return 4 if workload == "unet" else 8     # BASELINE.json configs[0] / configs[1]


def peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            p = json.load(f)
        return float(p["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except subprocess.TimeoutExpired:
            self.proc.kill()
        sm, mx, reasons = [], None, set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[0]))
                mx = float(r[1])
                for n, v in zip(names, r[3:7]):
                    if v.lower().startswith("active"):
                        reasons.add(n)
            except (ValueError, IndexError):
                continue
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": mx, "reasons": sorted(reasons),
                "samples": len(sm)}


# ----------------------------------------------------------------------------------------------------- CPU arm
def cpu_reference_step_time(channels, batch, steps, warmup):
    """Reference algorithm on the host cores: oracle port of the CAE train step (forward, loss_step, backward, Adam)."""
    import stroke_oracle as O
    from stroke_prediction_b200.common import data
    from stroke_prediction_b200.common.model.Cae3D import Cae3D, Dec3D, Enc3D
    torch.manual_seed(4)
    threads = os.cpu_count() or 1
    torch.set_num_threads(threads)
    cae = Cae3D(Enc3D(SIZE[1], SIZE[0], channels, 5, 1.0), Dec3D(SIZE[1], SIZE[0], channels, 5, 1.0))
    sd = O.clone_state(cae.state_dict(), requires_grad=True)
    names = [k for k, v in sd.items() if v.requires_grad]
    m = {k: torch.zeros_like(sd[k]) for k in names}
    v = {k: torch.zeros_like(sd[k]) for k in names}
    b = data.synthetic_cae_batch(batch, size=SIZE, seed=4)
    labels = b[data.KEY_LABELS]
    core, penu, lesion = labels[:, 0:1].contiguous(), labels[:, 1:2].contiguous(), labels[:, 2:3].contiguous()
    times = []
    for it in range(warmup + steps):
        t0 = time.perf_counter()
        step = O.time_to_treatment(b[data.KEY_GLOBAL])
        lat, rec = O.cae_forward(sd, channels, 1.0, True, core, penu, lesion, step)
        loss = O.cae_reconstruction_loss(lat, rec, core, penu, lesion, EPOCH)
        grads = O.grads_of(loss, sd)
        with torch.no_grad():
            for k in names:
                newp, m[k], v[k] = O.adam_step(sd[k], grads[k], m[k], v[k], it + 1)
                sd[k].copy_(newp)
        dt = time.perf_counter() - t0
        if it >= warmup:
            times.append(dt)
    return sum(times) / len(times), threads


def cpu_reference_unet_step_time(batch, steps, warmup):
    """Oracle port of the U-Net train step (UnetInference + UnetSegmentationLearner.loss_step + backward + Adam)."""
    import stroke_oracle as O
    from stroke_prediction_b200.common import data
    from stroke_prediction_b200.common.model.Unet3D import Unet3D
    torch.manual_seed(4)
    threads = os.cpu_count() or 1
    torch.set_num_threads(threads)
    sd = O.clone_state(Unet3D(CHANNELS["unet"]).state_dict(), requires_grad=True)
    names = [k for k, v in sd.items() if v.requires_grad]
    m = {k: torch.zeros_like(sd[k]) for k in names}
    v = {k: torch.zeros_like(sd[k]) for k in names}
    b = data.synthetic_unet_batch(batch, out_size=SIZE, seed=4)
    x, labels = b[data.KEY_IMAGES], b[data.KEY_LABELS]
    times = []
    for it in range(warmup + steps):
        t0 = time.perf_counter()
        c, p_ = O.unet_forward(sd, x, True)
        loss = O.unet_loss(c, p_, labels[:, 0:1], labels[:, 1:2])
        grads = O.grads_of(loss, sd)
        with torch.no_grad():
            for k in names:
                newp, m[k], v[k] = O.adam_step(sd[k], grads[k], m[k], v[k], it + 1, beta1=0.99)
                sd[k].copy_(newp)
        dt = time.perf_counter() - t0
        if it >= warmup:
            times.append(dt)
    return sum(times) / len(times), threads


def cpu_step_time(workload, batch, steps, warmup):
    if workload == "unet":
        return cpu_reference_unet_step_time(batch, steps, warmup)
    return cpu_reference_step_time(CHANNELS[workload], batch, steps, warmup)


def cpu_sample_text(workload, batch, steps, warmup):
    what = ("U-Net train step, batch %d x 2x68x168x168" % batch) if workload == "unet" else \
           ("CAE train step, batch %d x 1x28x128x128" % batch)
    return "oracle port of the reference %s (torch %s CPU), %d timed step(s) after %d warm-up" % (what, torch.__version__, steps, warmup)


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    sample_b = 2
    steps, warmup = max(1, min(args.steps, 5)), 1 if args.warmup > 0 else 0
    sec, threads = cpu_step_time(args.workload, sample_b, steps, warmup)
    val = sample_b / sec
    line = {
        "impl": "reference", "metric": metric_name(args.workload), "value": val, "unit": UNIT, "n_gpus": args.gpus, "steps": steps,
        "warmup": warmup, "ms_per_step": sec * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": workload_config(args, per_gpu_batch=sample_b),
        "cpu_baseline": {"value": val, "unit": UNIT, "cores": threads, "kind": "port",
                         "sample": cpu_sample_text(args.workload, sample_b, steps, warmup)},
        "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    emit(line)


def workload_config(args, per_gpu_batch):
    ch = " ".join(map(str, CHANNELS[args.workload]))
    if args.workload == "unet":
        wl = ("BASELINE configs[0]: Unet3D segmentation, channels %s, synthetic CBV/TTD 2x68x168x168 (28x128x128 padded by 20) "
              "-> 2 x 1x28x128x128, UnetSegmentationLearner forward + loss_step + backward + Adam" % ch)
        vol, act = "2x68x168x168 -> 28x128x128", "~1.1 GB"
    else:
        wl = ("BASELINE configs[1]: CAE channels %s, 1x28x128x128 core/penumbra/lesion masks, "
              "3 encoder + 4 decoder passes + loss_step + backward + Adam" % ch)
        vol, act = "1x28x128x128", "~0.75 GB"
    return {"workload": wl, "per_gpu_batch": per_gpu_batch, "global_batch": per_gpu_batch * args.gpus, "volume": vol,
            "parallelism": "dp%d (batch-sharded, gradient all-reduce, local BN/Dice statistics)" % args.gpus,
            "l2": "per-step working set (saved activations %s per volume) exceeds the 126 MB L2; no explicit flush" % act}


# ----------------------------------------------------------------------------------------------------- GPU arm
def algorithmic_bytes(name, key):
    """fwd 4(|X|+|Y|) + 4|W|; wgrad 4(|X|+|dY|) + 4|W| per launch, from the descriptor key."""
    try:
        parts = key.split()
        n = int(parts[0][1:])
        di, hi, wi, ci = (int(v) for v in parts[1][1:].split("x"))
        do, ho, wo, co = (int(v) for v in parts[2][1:].split("x"))
        k = int(parts[3][1:])
    except (ValueError, IndexError):
        return None
    return 4.0 * (n * di * hi * wi * ci + n * do * ho * wo * co) + 4.0 * (co * ci * k ** 3)


def algorithmic_flops(key):
    """2 x MACs of the layer (same for forward, dgrad and wgrad)."""
    try:
        parts = key.split()
        n = int(parts[0][1:])
        ci = int(parts[1][1:].split("x")[3])
        do, ho, wo, co = (int(v) for v in parts[2][1:].split("x"))
        k = int(parts[3][1:])
    except (ValueError, IndexError):
        return None
    return 2.0 * n * do * ho * wo * co * ci * k ** 3


def tensor_peak():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        with open(path) as f:
            return float(json.load(f)["bf16_tflops_sustained"])
    except (OSError, ValueError, KeyError):
        return 1405.0


def tensor_core_issue_factor(name, key):
    """bf16 MACs the tcgen05 tier executes per algorithmic fp32 MAC, or None when the layer runs on the FFMA tiers
    (mirrors sp_tc_corr_supported / sp_tc_wgrad_supported: 3x3x3 stride-1 layers with 9..16 channels on both sides)."""
    import re
    m = re.match(r"N(\d+) I(\d+)x(\d+)x(\d+)x(\d+) O(\d+)x(\d+)x(\d+)x(\d+) k(\d+) s(\d+)", key or "")
    if not m:
        return None
    ci, co, k, st = int(m.group(5)), int(m.group(9)), int(m.group(10)), int(m.group(11))
    if k != 3 or st != 1 or not (8 < ci <= 16 and 8 < co <= 16):
        return None
    if name == "sp_wgrad":
        return 12.0
    if name in ("sp_corr", "sp_corrT"):
        return 6.0
    return None


def measured_traffic(name, key):
    """dram bytes (read + write) per launch of this kernel from the committed `ncu --set full` capture, if one exists."""
    path = os.path.join(ROOT, "profiles", "traffic.json")
    try:
        with open(path) as f:
            return json.load(f).get("%s [%s]" % (name, key))
    except (OSError, ValueError):
        return None


def run_b200(args):
    import torch.distributed as dist
    from stroke_prediction_b200 import ops
    from stroke_prediction_b200.common import data
    from stroke_prediction_b200.common.dto import CaeDto as CaeDtoUtil
    from stroke_prediction_b200.common.metrics import BatchDiceLoss
    from stroke_prediction_b200.common.model.Cae3D import Cae3D, Dec3D, Enc3D
    from stroke_prediction_b200.learner.CaeReconstructionLearner import CaeReconstructionLearner
    from stroke_prediction_b200.optim import FusedAdam
    from stroke_prediction_b200.parallel import broadcast_parameters

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device — the hot path has no CPU fallback (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    if world != args.gpus and rank == 0:
        print("warning: --gpus %d but WORLD_SIZE %d" % (args.gpus, world), file=sys.stderr)

    channels = CHANNELS[args.workload]
    B = args.batch if args.batch > 0 else default_batch(args.workload)
    torch.manual_seed(4)
    if args.workload == "unet":
        from stroke_prediction_b200.common.dto import UnetDto as UnetDtoUtil
        from stroke_prediction_b200.common.model.Unet3D import Unet3D
        from stroke_prediction_b200.learner.UnetSegmentationLearner import UnetSegmentationLearner
        model = Unet3D(channels).to(dev).train()
        broadcast_parameters(model)
        opt = FusedAdam([p for p in model.parameters() if p.requires_grad], lr=1e-3, weight_decay=1e-5, betas=(0.99, 0.999))
        learner = UnetSegmentationLearner(None, None, model, opt, None, 1, BatchDiceLoss([1.0]), None, "/tmp/bench")
        host = data.synthetic_unet_batch(B, out_size=SIZE, seed=4 + rank)
        h2d = host[data.KEY_IMAGES].numel() * 4 + host[data.KEY_LABELS].numel() * 4
    else:
        model = Cae3D(Enc3D(SIZE[1], SIZE[0], channels, 5, 1.0), Dec3D(SIZE[1], SIZE[0], channels, 5, 1.0)).to(dev).train()
        broadcast_parameters(model)
        opt = FusedAdam([p for p in model.parameters() if p.requires_grad], lr=1e-3, weight_decay=1e-5, betas=(0.9, 0.999))
        learner = CaeReconstructionLearner(None, None, model, opt, None, 1, None, "/tmp/bench", BatchDiceLoss([1.0]))
        host = data.synthetic_cae_batch(B, size=SIZE, seed=4 + rank)
        h2d = host[data.KEY_LABELS].numel() * 4 + host[data.KEY_GLOBAL].numel() * 4 + 3 * B * 4
    if world > 1:
        learner.enable_data_parallel()
    host = {k: (v.pin_memory() if torch.is_tensor(v) else v) for k, v in host.items()}
    # the loss (4 bytes) + the evaluation counts of batch_metrics_step (4 doubles per compared pair) come back every step
    d2h = 4 + 32 * (2 if args.workload == "unet" else 3)

    # device-resident inputs for `value`
    if args.workload == "unet":
        x_dev = ops.as_vol(host[data.KEY_IMAGES].to(dev))
        lab = ops.as_vol(host[data.KEY_LABELS].to(dev))
        core_gt, penu_gt = ops.extract_channel(lab, 0), ops.extract_channel(lab, 1)

        def make_dto():
            return UnetDtoUtil.init_dto(x_dev, core_gt, penu_gt)
    else:
        with torch.no_grad():
            dto0 = learner.init_clinical_variables(host, None)
            dto0 = learner.init_gtruth_segm_variables(host, dto0)
        res = dto0.given_variables

        def make_dto():
            return CaeDtoUtil.init_dto(res.globals, res.time_to_treatment, res.scalar_types.core, res.scalar_types.penu,
                                       None, None, res.gtruth.core, res.gtruth.penu, res.gtruth.lesion)

    def step_resident():
        dto = model(make_dto())
        loss = learner.loss_step(dto, EPOCH)
        opt.zero_grad()
        loss.backward()
        if learner._grad_sync is not None:
            learner._grad_sync()
        opt.step()
        return loss

    def step_e2e():
        return learner.train_batch(host, EPOCH).loss

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            fn()
        e1.record()
        barrier()
        ms = e0.elapsed_time(e1)
        if world > 1:
            t = torch.tensor([ms], device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        return ms

    import contextlib
    with contextlib.redirect_stdout(open(os.devnull, "w")):   # loss_step-style prints must not pollute the JSON line
        for _ in range(args.warmup if args.quick else max(args.warmup, 3)):
            step_resident()
        sampler = ClockSampler(local)
        if rank == 0:
            sampler.start()
        ops.reset_launch_count()
        ms_total = timed(step_resident, args.steps)
        launches = ops.launch_count()
        clocks = sampler.stop() if rank == 0 else None
        if args.quick:
            if rank == 0:
                sys.stderr.write("quick: %.3f ms/step, %.2f volumes/s\n" % (ms_total / args.steps, world * B / (ms_total / args.steps * 1e-3)))
            if world > 1:
                dist.destroy_process_group()
            return
        for _ in range(2):
            step_e2e()
        ms_e2e = timed(step_e2e, args.steps)

        # per-kernel attribution with CUDA events on the launching stream (one extra step, not part of `value`)
        # (every rank runs the step — it contains the gradient all-reduce — but only rank 0 records events)
        prof = None
        if rank == 0:
            ops.start_profile()
        step_resident()
        if rank == 0:
            prof = ops.stop_profile()
        barrier()
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    ms_step = ms_total / args.steps
    value = world * B / (ms_step * 1e-3)
    e2e_val = world * B / (ms_e2e / args.steps * 1e-3)

    hbm_peak, peak_src = peaks()
    fam = {}
    for (name, key), ts in prof.items():
        f = fam.setdefault((name, key), [0.0, 0])
        f[0] += sum(ts)
        f[1] += len(ts)
    total_prof = sum(v[0] for v in fam.values())
    (top_name, top_key), (top_ms, top_n) = max(fam.items(), key=lambda kv: kv[1][0])
    abytes = algorithmic_bytes(top_name, top_key)
    aflops = algorithmic_flops(top_key)
    avg_ms = top_ms / top_n
    achieved = (abytes / (avg_ms * 1e-3) / 1e9) if abytes else None
    tflops = (aflops / (avg_ms * 1e-3) / 1e12) if aflops else None
    traffic = measured_traffic(top_name, top_key)
    roofline = {"bound": "hbm", "kernel": "%s [%s]" % (top_name, top_key), "achieved": achieved, "peak": hbm_peak,
                "peak_source": peak_src, "unit": "GB/s", "frac": (achieved / hbm_peak) if achieved else None,
                "traffic": traffic, "avg_launch_ms": avg_ms, "launches_per_step": top_n,
                "fp32_tflops": tflops, "ffma_peak_tflops": FFMA_PEAK_TFLOPS,
                "ffma_frac": (tflops / FFMA_PEAK_TFLOPS) if tflops else None,
                "share_of_step": top_ms / total_prof if total_prof else None,
                "algorithmic_bytes_per_launch": abytes}
    tc_factor = tensor_core_issue_factor(top_name, top_key)
    if tc_factor and tflops:
        # exact-fp32 emulation on tcgen05: every operand is three bf16 terms, so the tensor pipe executes `tc_factor` bf16
        # MACs per algorithmic fp32 MAC (9 products x 64/48 padded rows for wgrad, 6 products for forward / dgrad)
        tc_peak = tensor_peak()
        roofline.update({"tier": "tcgen05 split-bf16 (3 terms per fp32 operand)", "tensor_issue_factor": tc_factor,
                         "tensor_executed_tflops": tflops * tc_factor, "tensor_peak_tflops": tc_peak,
                         "tensor_frac": tflops * tc_factor / tc_peak,
                         "note": "algorithmic HBM fraction reported for the contract; the kernel is bound by the tensor pipe / "
                                 "shared-memory operand fetch of the split-bf16 MMAs it issues (M = 64 single-CTA MMAs cap "
                                 "at half the dense bf16 peak), see DESIGN.md 3.2"})
    else:
        roofline["note"] = ("fp32 FFMA direct convolution: arithmetic intensity of this layer is above the FFMA ridge, so the "
                            "HBM fraction is reported for the contract while the binding roof is FP32 FFMA (see DESIGN.md)")
    if args.dump_breakdown:
        with open(args.dump_breakdown, "w") as f:
            for k, v in sorted(fam.items(), key=lambda kv: -kv[1][0]):
                f.write("%9.3f ms  x%-3d %s [%s]\n" % (v[0], v[1], k[0], k[1]))
    top5 = sorted(fam.items(), key=lambda kv: -kv[1][0])[:12]
    breakdown = [{"kernel": "%s [%s]" % k, "ms_per_step": round(v[0], 3), "launches": v[1]} for k, v in top5]
    by_op = {}
    for (name, key), v in fam.items():
        o = by_op.setdefault(name, [0.0, 0])
        o[0] += v[0]
        o[1] += v[1]
    op_breakdown = {k: {"ms_per_step": round(v[0], 3), "launches": v[1]} for k, v in sorted(by_op.items(), key=lambda kv: -kv[1][0])}

    cpu = None
    if world == 1 and not args.no_cpu_baseline:
        sec, threads = cpu_step_time(args.workload, 2, 4, 1)
        cpu = {"value": 2 / sec, "unit": UNIT, "cores": threads, "kind": "port", "sample": cpu_sample_text(args.workload, 2, 4, 1)}

    line = {"metric": metric_name(args.workload), "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
            "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
            "data": "synthetic", "config": workload_config(args, B),
            "e2e": {"value": e2e_val, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "ms_per_step": ms_e2e / args.steps, "api": "%s.train_batch(host_batch, epoch)" % type(learner).__name__},
            "gpu_launches": launches, "clocks": clocks, "roofline": roofline, "kernel_breakdown": breakdown, "op_breakdown": op_breakdown,
            "cpu_baseline": cpu}
    emit(line)
    if world > 1:
        dist.destroy_process_group()


_REAL_STDOUT = None


def claim_stdout():
    """Keep stdout for the ONE JSON line: everything else that writes to fd 1 (NCCL's version banner, library prints,
    loss_step-style prints) is routed to stderr."""
    global _REAL_STDOUT
    if _REAL_STDOUT is None:
        sys.stdout.flush()
        _REAL_STDOUT = os.fdopen(os.dup(1), "w")
        os.dup2(2, 1)
        sys.stdout = sys.stderr


def emit(line):
    out = _REAL_STDOUT if _REAL_STDOUT is not None else sys.stdout
    out.write(json.dumps(line) + "\n")
    out.flush()


def main():
    claim_stdout()
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="cae200", choices=sorted(CHANNELS))
    ap.add_argument("--batch", type=int, default=0, help="per-GPU batch (default: 8 for the CAE, 4 for the U-Net)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--dump-breakdown", default=None, help="write the full per-kernel CUDA-event attribution to this file")
    ap.add_argument("--quick", action="store_true", help="profiling aid: resident-input steps only (no e2e / attribution / CPU legs)")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_b200(args)


if __name__ == "__main__":
    main()
