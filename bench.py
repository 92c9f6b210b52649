#!/usr/bin/env python
"""Benchmark of the phoneme_contrast training hot path on B200 (contract: see the task statement / DESIGN.md).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload train_cnn_deep|train_cnn_small|frontend]
    python bench.py --impl reference ...      # the reference's CPU path (oracle port) on the host cores

Primary metric: SupCon training samples/s (BASELINE.json `metric`): one step = forward + SupCon(T=0.15) + backward +
clip(1.0) + Adam on one synthetic batch. Default workload = BASELINE configs[2] (cnn_deep, 256 views per GPU); under
torchrun each rank processes its own 256 views (weak scaling) with the embedding all_gather / gradient all-reduce of
phoneme_contrast_b200.parallel. Rank 0 prints ONE JSON line. The other two BASELINE workloads (front-end clips/s on
65 536 clips, cnn_small at 64 views) are measured in the same run at N=1 and reported under "also".
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import tempfile
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

T_SUPCON = 0.15
WORKLOADS = {
    "train_cnn_deep": dict(arch="phoneme_cnn_deep", views=256, desc="cnn_deep PhonemeNetDeep (64->512 ch residual) + SupCon T=0.15, 256 views/GPU, 40x101 MFCC (BASELINE configs[2])"),
    "train_cnn_deep_4096": dict(arch="phoneme_cnn_deep", views=512, desc="cnn_deep PhonemeNetDeep + SupCon T=0.15 data-parallel, 512 views/GPU = global batch 4096 at 8 GPUs, 40x101 MFCC (BASELINE configs[4])"),
    "train_cnn_small": dict(arch="phoneme_cnn", views=64, desc="cnn_small PhonemeNet + SupCon T=0.15, emb 128, 64 views (8x4x2), 40x101 MFCC (BASELINE configs[0] on GPU)"),
    "frontend": dict(clips=65536, desc="MFCC front end + 2-view augmentation, 65 536 synthetic 1 s 16 kHz clips (BASELINE configs[1])"),
    "supcon_8192": dict(n=8192, d=128, desc="SupCon loss alone, forward+backward, 8192 views x 128-d, 38 classes, rows sharded over the ranks with embedding/label/row-stat all_gather (BASELINE configs[3])"),
}
FLOP_PER_SAMPLE = {"phoneme_cnn": 3 * 298.07e6, "phoneme_cnn_deep": 3 * 568.59e6}   # fwd MAC*2, x3 for training


# mean dram__bytes_read.sum + dram__bytes_write.sum per launch of the kernels behind each convolution entry point, from the round-2
# `ncu --set full` capture of one eager cnn_deep step (profiles/r2f_conv_full.md; round 1: profiles/r1d_tc_conv_full.md)
TRAFFIC_NCU = {}
TRAFFIC_SRC = None
try:
    import re as _re
    for _src in ("r2f_conv_full.md", "r1d_tc_conv_full.md"):
        _path = os.path.join(ROOT, "profiles", _src)
        if not os.path.exists(_path):
            continue
        _txt = open(_path).read()
        # forward and data gradient share kernels (conv_halo*, igemm_tc): one figure for both entry points
        for _entry, _pats in (("pc_conv_fwd", ("conv_halo", "igemm_tc_kernel")), ("pc_conv_dgrad", ("conv_halo", "igemm_tc_kernel")),
                              ("pc_conv_wgrad", ("wgrad_halo_kernel", "wgrad_tc_kernel"))):
            _vals = []
            for _sec in _txt.split("## ")[1:]:
                if any(_p in _sec.splitlines()[0] for _p in _pats):
                    _r = _re.search(r"dram__bytes_read.sum \| ([0-9.]+) \| (\w+)", _sec)
                    _w = _re.search(r"dram__bytes_write.sum \| ([0-9.]+) \| (\w+)", _sec)
                    _u = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
                    if _r and _w:
                        _vals.append(float(_r.group(1)) * _u.get(_r.group(2), 1.0) + float(_w.group(1)) * _u.get(_w.group(2), 1.0))
            if _vals:
                TRAFFIC_NCU[_entry] = sum(_vals) / len(_vals)
        if TRAFFIC_NCU:
            TRAFFIC_SRC = "profiles/" + _src
            break
except Exception:
    pass


def ncu_traffic(md_name, kernel_substrings, per="launch"):
    """dram__bytes_read.sum + dram__bytes_write.sum of the kernels whose section title contains one of the substrings, from an
    `ncu --set full` summary under profiles/: the mean over the launches of each kernel, summed over the kernels (per step of one launch
    each), or None when the file is absent."""
    try:
        import re
        txt = open(os.path.join(ROOT, "profiles", md_name)).read()
        unit = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
        total = 0.0
        for sub in kernel_substrings:
            vals = []
            for sec in txt.split("## ")[1:]:
                if sub in sec.splitlines()[0]:
                    r = re.search(r"dram__bytes_read.sum \| ([0-9.]+) \| (\w+)", sec)
                    w = re.search(r"dram__bytes_write.sum \| ([0-9.]+) \| (\w+)", sec)
                    if r and w:
                        vals.append(float(r.group(1)) * unit.get(r.group(2), 1.0) + float(w.group(1)) * unit.get(w.group(2), 1.0))
            if not vals:
                return None
            total += sum(vals) / len(vals)
        return total
    except Exception:
        return None


def peaks():
    try:
        return json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))), "measured"
    except Exception:
        return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0}, "fallback"


# ------------------------------------------------------------------------------------------------ clocks
class NvmlClockSampler:
    """SM clock + throttle reasons through NVML from a sampling thread (~2 ms cadence), so that even a 90 ms timed region holds
    dozens of samples. Falls back to the nvidia-smi poller (ClockSampler) when NVML is unavailable."""

    def __init__(self, index=0):
        import threading

        import pynvml
        self.nv = pynvml
        pynvml.nvmlInit()
        self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
        self.rows, self.stop_flag = [], False
        self.t = threading.Thread(target=self._run, daemon=True)
        self.t.start()

    def _run(self):
        nv = self.nv
        reasons_fn = getattr(nv, "nvmlDeviceGetCurrentClocksEventReasons", None) or nv.nvmlDeviceGetCurrentClocksThrottleReasons
        while not self.stop_flag:
            try:
                self.rows.append((nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM), int(reasons_fn(self.h))))
            except Exception:
                pass
            time.sleep(0.002)

    def stop(self):
        self.stop_flag = True
        self.t.join(timeout=2)
        nv = self.nv
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": len(self.rows), "source": "nvml thread, 2 ms cadence"}
        if not self.rows:
            return out
        try:
            out["sm_max_mhz"] = float(nv.nvmlDeviceGetMaxClockInfo(self.h, nv.NVML_CLOCK_SM))
        except Exception:
            pass
        sm = [r[0] for r in self.rows]
        hi = [v for v in sm if v >= 0.5 * max(sm)] or sm
        out["sm_mhz"] = float(np.median(hi))
        bits = 0
        for _, b in self.rows:
            bits |= b
        names = {"hw_slowdown": 0x8, "hw_thermal_slowdown": 0x40, "sw_thermal_slowdown": 0x20, "sw_power_cap": 0x4}
        out["reasons"] = sorted(n for n, m in names.items() if bits & m)
        return out


def clock_sampler(index=0):
    try:
        return NvmlClockSampler(index)
    except Exception:
        return ClockSampler(index)


class ClockSampler:
    Q = "clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index=0):
        self.f = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        self.p = None
        try:
            self.p = subprocess.Popen(["nvidia-smi", f"--id={index}", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "20"],
                                      stdout=self.f, stderr=subprocess.DEVNULL)
        except Exception:
            self.p = None

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        if self.p is None:
            return out
        self.p.terminate()
        try:
            self.p.wait(timeout=5)
        except Exception:
            self.p.kill()
        self.f.flush()
        rows = [r.strip().split(",") for r in open(self.f.name).read().strip().splitlines() if r.strip()]
        os.unlink(self.f.name)
        sm, reasons, mx = [], set(), None
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in rows:
            try:
                sm.append(float(r[0]))
                mx = float(r[1])
                for n, v in zip(names, r[3:7]):
                    if v.strip().lower().startswith("active"):
                        reasons.add(n)
            except Exception:
                continue
        if sm:
            hi = [v for v in sm if v >= 0.5 * max(sm)] or sm     # samples under load
            out.update(sm_mhz=float(np.median(hi)), sm_max_mhz=mx, reasons=sorted(reasons), samples=len(sm))
        return out


# ------------------------------------------------------------------------------------------------ synthetic data
def train_inputs(views, seed, device, n_buffers=4):
    """x = randn(views,1,40,101); y = repeat_interleave(arange(views/2)//4, 2) (K classes x 4 samples x 2 views)."""
    g = torch.Generator().manual_seed(seed)
    xs = [torch.randn(views, 1, 40, 101, generator=g) for _ in range(n_buffers)]
    y = torch.repeat_interleave(torch.arange(views // 2) // 4, 2).to(torch.int64)
    return xs, y


def build_trainer(arch, device, parallel):
    import logging
    if not logging.getLogger("bench").handlers:          # warnings (e.g. a failed graph capture) must be visible on stderr
        logging.getLogger("bench").addHandler(logging.StreamHandler(sys.stderr))

    from phoneme_contrast_b200.models import model_registry
    from phoneme_contrast_b200.training import ContrastiveTrainer, FusedClipAdam, get_loss_fn
    torch.manual_seed(42)
    model = model_registry.create(arch, {}).to(device)           # reference defaults: dropout 0.1 / 0.2, attention, emb 128
    if parallel is not None:
        parallel.broadcast_parameters(model)
    opt = FusedClipAdam(model.parameters(), lr=3e-4, weight_decay=1e-4)
    loss_fn = get_loss_fn("supervised_contrastive", temperature=T_SUPCON)
    tr = ContrastiveTrainer(model, [], None, loss_fn, opt, None, torch.device(device), {"gradient_clip_val": 1.0, "progress": False},
                            tempfile.mkdtemp(prefix="pc_bench_"), logging.getLogger("bench"), parallel=parallel)
    model.train()
    return tr


def barrier(parallel):
    if parallel is not None:
        torch.distributed.barrier()
    torch.cuda.synchronize()


def max_over_ranks(ms, parallel, device):
    if parallel is None:
        return ms
    t = torch.tensor([ms], device=device, dtype=torch.float64)
    torch.distributed.all_reduce(t, op=torch.distributed.ReduceOp.MAX)
    return float(t.item())


# ------------------------------------------------------------------------------------------------ GPU arms
def bench_train(workload, steps, warmup, parallel, device, want_profile=True, use_graph=True, views=None):
    from phoneme_contrast_b200 import _lib
    cfg = WORKLOADS[workload]
    arch, views = cfg["arch"], int(views or cfg["views"])
    rank = 0 if parallel is None else parallel.rank
    world = 1 if parallel is None else parallel.world_size
    tr = build_trainer(arch, device, parallel)
    # multi-rank steps replay ONE captured graph whose exchanges are this library's peer-memory kernels over NVLink (training/graph.py:
    # GraphedDPStepPeer; PC_DP_EXCHANGE=nccl selects the five captured segments with NCCL calls between them, GraphedDPStep)
    use_graph = bool(use_graph)
    tr.config["cuda_graph"] = use_graph
    xs_h, y_h = train_inputs(views, 1000 + rank, device)
    xs = [x.to(device) for x in xs_h]
    y = y_h.to(device)

    for i in range(warmup):
        tr.step(xs[i % len(xs)], y)
    barrier(parallel)
    sampler = clock_sampler(torch.cuda.current_device()) if rank == 0 else None
    l0 = _lib.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(steps):
        loss = tr.step(xs[i % len(xs)], y)
    e1.record()
    barrier(parallel)
    ms = max_over_ranks(e0.elapsed_time(e1), parallel, device)
    launches = _lib.launch_count() - l0
    if use_graph and tr._graphed is not None:
        # replays do not pass through the launch counter: count the kernels of one eagerly executed step instead
        c0 = _lib.launch_count()
        tr.train_step(xs[0], y)
        torch.cuda.synchronize()
        launches = (_lib.launch_count() - c0) * steps
    clocks = sampler.stop() if sampler else None
    assert torch.isfinite(loss).item(), "training diverged"
    if tr._graphed and hasattr(tr._graphed, "check"):
        tr._graphed.check()            # a peer barrier that timed out would have invalidated the timed steps
    exchange = type(tr._graphed).__name__ if tr._graphed else "eager"

    # end to end through the public API with HOST buffers: pinned H2D of this step's views + labels, loss read back every step
    xs_p = [x.pin_memory() for x in xs_h]
    y_p = y_h.pin_memory()
    # (the trainer's own loop: ContrastiveTrainer.prefetch copies batch i + 1 on a side stream while step i computes)
    def host_batches(n):
        for i in range(n):
            yield {"views": xs_p[i % len(xs_p)], "label": y_p}
    # ContrastiveTrainer.run_batches = the trainer's own epoch loop body: every step's loss is read back on the host (D2H into pinned
    # memory behind the step, handed over while the next step runs; the last one before the call returns)
    host_losses = []
    tr.run_batches(host_batches(2), on_loss=lambda i, v: host_losses.append(v))
    torch.cuda.synchronize()
    barrier(parallel)
    host_losses.clear()
    t0 = time.perf_counter()
    tr.run_batches(host_batches(steps), on_loss=lambda i, v: host_losses.append(v))
    torch.cuda.synchronize()
    barrier(parallel)
    assert len(host_losses) == steps and all(v == v for v in host_losses), "e2e leg: a step's loss did not reach the host"
    e2e_ms = max_over_ranks((time.perf_counter() - t0) * 1e3, parallel, device)

    prof = None
    if want_profile:
        # every rank runs the profiled steps (they contain collectives); only rank 0 records events
        n_prof = min(steps, 5)
        os.environ["PC_WGRAD_STREAM"] = "0"      # serialise the weight-gradient lane so per-entry-point durations do not overlap
        if rank == 0:
            _lib.profile_begin()
        for i in range(n_prof):
            tr.train_step(xs[i % len(xs)], y)
        os.environ["PC_WGRAD_STREAM"] = "1"
        if rank == 0:
            prof = _lib.profile_end()
            for v in prof.values():
                v["ms"] /= n_prof
                v["calls"] /= n_prof
                v["work"] /= n_prof
    barrier(parallel)
    return dict(arch=arch, views=views, ms_per_step=ms / steps, value=views * world * steps / (ms * 1e-3),
                e2e_value=views * world * steps / (e2e_ms * 1e-3), launches=launches / steps, clocks=clocks, prof=prof,
                h2d=views * 4040 * 4 + views * 8, d2h=4, loss=float(loss), exchange=exchange)


def bench_supcon(steps, warmup, parallel, device, n=8192, d=128):
    """fwd+bwd of the loss on N = 8192 global views; rank r owns rows [r*n/R, (r+1)*n/R). value = global views/s."""
    from phoneme_contrast_b200 import _lib
    from phoneme_contrast_b200.training import get_loss_fn
    rank = 0 if parallel is None else parallel.rank
    world = 1 if parallel is None else parallel.world_size
    g = torch.Generator().manual_seed(0)
    F_all = torch.nn.functional.normalize(torch.randn(n, d, generator=g), dim=1)
    y_all = torch.randint(0, 38, (n,), generator=g)
    nl = n // world
    f = F_all[rank * nl:(rank + 1) * nl].to(device).requires_grad_(True)
    y = y_all[rank * nl:(rank + 1) * nl].to(device)
    loss_fn = get_loss_fn("supervised_contrastive", temperature=T_SUPCON)

    graphed = parallel.graphed_loss(loss_fn, f.detach(), y) if parallel is not None else None

    def step():
        if graphed is not None:          # three graph segments + two all_gathers, static buffers (parallel.GraphedShardedLoss)
            loss, dF = graphed(f.detach(), None)
            f.grad = dF
            return loss
        f.grad = None
        loss = loss_fn(f, y)
        loss.backward()
        return loss

    for _ in range(warmup):
        step()
    barrier(parallel)
    l0 = _lib.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        loss = step()
    e1.record()
    barrier(parallel)
    ms = max_over_ranks(e0.elapsed_time(e1), parallel, device) / steps
    launches = (_lib.launch_count() - l0) / steps
    # e2e: embeddings start in pinned host memory, gradient rows are read back
    f_h = F_all[rank * nl:(rank + 1) * nl].clone().pin_memory()
    g_h = torch.empty(nl, d).pin_memory()
    barrier(parallel)
    t0 = time.perf_counter()
    for _ in range(steps):
        f.data.copy_(f_h, non_blocking=True)
        step()
        g_h.copy_(f.grad, non_blocking=True)
        torch.cuda.synchronize()
    barrier(parallel)
    e2e_ms = max_over_ranks((time.perf_counter() - t0) * 1e3, parallel, device) / steps
    if graphed is not None and getattr(graphed, "region", None) is not None:
        assert graphed.region.error() == 0, "a peer barrier timed out during the timed steps"
    flops = 8.0 * n * n * d / world          # 2N^2D forward + 2N^2D recompute + 4N^2D backward products, per rank 1/R
    return dict(ms_per_step=ms, value=n / (ms * 1e-3), e2e_value=n / (e2e_ms * 1e-3), launches=launches, tflops=flops / (ms * 1e-3) / 1e12,
                loss=float(loss), h2d=nl * d * 4, d2h=nl * d * 4,
                exchange=("single process" if graphed is None else ("peer memory, one graph" if getattr(graphed, "region", None) is not None else "NCCL all_gathers between three graph segments")))


def roofline_from_profile(prof, pk, pk_kind):
    """Dominant C-ABI entry point by device time; convolution entry points carry algorithmic FLOPs (2*M*N*K per launch set)."""
    if not prof:
        return None
    total = sum(v["ms"] for v in prof.values())
    name = max(prof, key=lambda k: prof[k]["ms"])
    conv = {k: v for k, v in prof.items() if k.startswith("pc_conv_") and v["work"] > 0}
    conv_ms = sum(v["ms"] for v in conv.values())
    conv_flop = sum(v["work"] for v in conv.values())
    top = prof[name]
    kernel_of = {"pc_conv_fwd": "halo::conv_halo_kernel / conv_halo_res_kernel + tcconv::igemm_tc_kernel for the strided layers (entry pc_conv_fwd)",
                 "pc_conv_dgrad": "halo::conv_halo_kernel / conv_halo_res_kernel + tcconv::igemm_tc_kernel for the strided layers (entry pc_conv_dgrad)",
                 "pc_conv_wgrad": "halowg::wgrad_halo_kernel (stride-1 3x3 layers) + tcwg::wgrad_tc_kernel (strided / 1x1 layers), incl. their split reduces (entry pc_conv_wgrad)"}
    ach = (top["work"] / (top["ms"] * 1e-3)) / 1e12 if top["work"] else None
    peak = pk.get("bf16_tflops_sustained", pk["bf16_tflops"])
    return {"bound": "tensor", "kernel": kernel_of.get(name, name), "achieved": ach, "peak": peak, "unit": "TFLOP/s", "frac": (ach / peak) if ach else None,
            "traffic": TRAFFIC_NCU.get(name), "traffic_note": f"mean dram read+write bytes per launch over this entry point's kernel launches captured with ncu --set full ({TRAFFIC_SRC})",
            "peak_source": f"{pk_kind} cuBLAS bf16 sustained (MEASURED_PEAKS.json); kernel timed inside a long step; achieved counts the convolution's algorithmic fp32 FLOPs once, while the kernels issue 3 fp16 tensor-core products per operand pair (FP16x2 split for fp32-level accuracy), i.e. their own ceiling is peak/3",
            "kernel_share_of_step": top["ms"] / total, "kernel_ms_per_step": top["ms"], "launches_per_step": top["calls"],
            "all_conv": {"ms_per_step": conv_ms, "tflops": conv_flop / (conv_ms * 1e-3) / 1e12 if conv_ms else None,
                         "share_of_step": conv_ms / total},
            "by_entry_point_ms": {k: round(v["ms"], 4) for k, v in sorted(prof.items(), key=lambda kv: -kv[1]["ms"])}}


def frontend_inputs(n_clips, device, seed=0):
    """w = 0.1*randn(n,16000); 1 clip in 8 has its last 25-75 % zeroed."""
    g = torch.Generator(device=device).manual_seed(seed)
    w = torch.randn(n_clips, 16000, generator=g, device=device) * 0.1
    idx = torch.arange(0, n_clips, 8, device=device)
    cut = (16000 * (0.25 + 0.5 * torch.rand(idx.numel(), generator=g, device=device))).long()
    mask = torch.arange(16000, device=device)[None, :] >= cut[:, None]
    w[idx] = w[idx].masked_fill(mask, 0.0)
    return w


AUG_CFG = {"time_mask": {"enabled": True, "max_width": 30, "prob": 0.5}, "freq_mask": {"enabled": True, "max_width": 10, "prob": 0.5},
           "noise": {"enabled": True, "min_snr": 0.001, "max_snr": 0.005, "prob": 0.5}}


def bench_frontend(steps, warmup, device, n_clips=65536, desc_clips=2048):
    """clips/s of MFCC + 2-view augmentation; one step = one pass over all n_clips clips (4.19 GB in, 2.1 GB out > L2)."""
    from phoneme_contrast_b200 import _lib
    from phoneme_contrast_b200.datasets import MFCCExtractor, build_augmentation_pipeline, build_view_descriptors, pack_view_descs
    ext = MFCCExtractor()
    pipe = build_augmentation_pipeline(AUG_CFG)
    # descriptor table from the reference's RNG calls for `desc_clips` items, tiled over the batch (it is a cached,
    # epoch-independent function of (idx, view); building all 131 072 rows with Python's RNG would only time the host)
    recs, _ = build_view_descriptors(range(desc_clips), 2, 40, 101, pipe)
    reps = n_clips // desc_clips
    recs = np.tile(recs, reps)
    recs["clip"] = np.repeat(np.arange(n_clips), 2)
    views = pack_view_descs(recs, device)
    wave = frontend_inputs(n_clips, device)
    for _ in range(warmup):
        out = ext.forward_views(wave, views, 2 * n_clips)
    torch.cuda.synchronize()
    l0 = _lib.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        out = ext.forward_views(wave, views, 2 * n_clips)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / steps
    launches = (_lib.launch_count() - l0) / steps
    assert torch.isfinite(out[:64]).all().item()
    # e2e: waveforms in pinned host memory, features read back to the host (bounded: 8192 clips per call)
    ne = 8192
    wave_h = wave[:ne].cpu().pin_memory()
    views_e = pack_view_descs(recs[:2 * ne], device)
    out_h = torch.empty(2 * ne, 1, 40, 101).pin_memory()
    for _ in range(2):
        out_h.copy_(ext.forward_views(wave_h.to(device, non_blocking=True), views_e, 2 * ne), non_blocking=True)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(3):
        out_h.copy_(ext.forward_views(wave_h.to(device, non_blocking=True), views_e, 2 * ne), non_blocking=True)
        torch.cuda.synchronize()
    e2e = 3 * ne / (time.perf_counter() - t0)
    bytes_per_clip = 64000 + 2 * 16160
    return dict(ms_per_step=ms, value=n_clips / (ms * 1e-3), e2e_value=e2e, launches=launches, gbs=n_clips * bytes_per_clip / (ms * 1e-3) / 1e9,
                h2d=ne * 64000, d2h=2 * ne * 16160, e2e_sample=f"{ne} clips per call, pinned host waveforms in, host features out")


def bench_pipeline(steps, warmup, device, n_items=2048, views_per_item=2, items_per_batch=128):
    """End-to-end 'clips -> optimiser step' through the device-resident dataset path: per step the host picks the batch's item
    indices (8 x ... class-balanced recipe scaled to 128 items x 2 views = 256 views), ONE fused front-end launch turns the cached
    GPU waveforms into augmented MFCC views, and ContrastiveTrainer.step (graph replay) trains cnn_deep on them."""
    from phoneme_contrast_b200.datasets import DeviceFrontendLoader, MFCCExtractor, PhonemeContrastiveDataset, build_augmentation_pipeline
    g = torch.Generator().manual_seed(7)
    waves = 0.1 * torch.randn(n_items, 16000, generator=g)
    labels = (torch.arange(n_items) % 38).tolist()
    ds = PhonemeContrastiveDataset(list(range(n_items)), labels, [{}] * n_items, MFCCExtractor(), build_augmentation_pipeline(AUG_CFG),
                                   {"target_sr": 16000, "max_length_ms": 1000, "contrastive": {"views_per_sample": views_per_item}},
                                   mode="train", device=device, device_frontend=True, waveforms=waves)
    ds.use_cache = True
    ds.waveform_cache = {i: waves[i:i + 1] for i in range(n_items)}
    t0 = time.perf_counter()
    ds.cache_on_device()
    ds.view_descriptors()
    setup_s = time.perf_counter() - t0
    tr = build_trainer("phoneme_cnn_deep", device, None)
    tr.config["cuda_graph"] = True
    rs = np.random.RandomState(3)
    by_label = {}
    for i, l in enumerate(labels):
        by_label.setdefault(l, []).append(i)

    def sample_batch():      # K classes x 4 items (ContrastiveBatchSampler's recipe, samplers.py:13-21)
        out = []
        for c in rs.choice(38, items_per_batch // 4, replace=False if items_per_batch // 4 <= 38 else True):
            out.extend(rs.choice(by_label[int(c)], 4, replace=False))
        return out
    batches = [sample_batch() for _ in range(warmup + steps)]
    loader = DeviceFrontendLoader(ds, batches)
    it = iter(loader)
    import itertools
    for v, y in tr.prefetch(itertools.islice(it, warmup)):
        tr.step(v, y)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    loss = None
    for v, y in tr.prefetch(itertools.islice(it, steps)):      # the trainer's own loop: batch i + 1 is produced on a side stream during step i
        loss = tr.step(v, y)
    loss = float(loss.item())
    dt = time.perf_counter() - t0
    n_views = items_per_batch * views_per_item
    return {"value": n_views * steps / dt, "unit": "samples/s", "ms_per_step": dt / steps * 1e3, "final_loss": loss,
            "workload": f"device-resident PhonemeContrastiveDataset ({n_items} cached 1 s clips) -> fused MFCC + 2-view augmentation -> cnn_deep SupCon step, {n_views} views/step; wall clock incl. host batch sampling and descriptor upload",
            "setup_s": setup_s, "h2d_bytes_per_step": n_views * 32 + items_per_batch * 8}


# ------------------------------------------------------------------------------------------------ CPU arms (reference / oracle port)
_REF = None


def have_reference():
    """True when the reference's own modules were vendored into oracle/_ref (oracle/build_ref.py) and import on this box."""
    global _REF
    if _REF is None:
        try:
            from oracle import build_ref
            _REF = bool(build_ref.import_reference())
        except Exception:
            _REF = False
    return _REF


def ref_kind():
    return "reference" if have_reference() else "port"


def _train_step_fn(arch, views, device="cpu"):
    """One training step of the reference path as the trainer runs it (trainer.py:126-164: forward, SupCon T=.15, backward,
    clip_grad_norm_(1.0), Adam(lr 3e-4, L2 wd 1e-4), loss.item()), on `device`: the reference's own nn.Modules when
    oracle/_ref is present (dropout drawn by nn.Dropout2d), else the oracle's restatement with the same ATen ops."""
    xs, y = train_inputs(views, 1000, device, n_buffers=2)
    xs, y = [x.to(device) for x in xs], y.to(device)
    if have_reference():
        from src.models import model_registry
        from src.training.losses import get_loss_fn
        torch.manual_seed(42)
        model = model_registry.create(arch, {}).to(device)
        model.train()
        loss_fn = get_loss_fn("supervised_contrastive", temperature=T_SUPCON)
        opt = torch.optim.Adam(model.parameters(), lr=3e-4, weight_decay=1e-4)

        def step(i):
            opt.zero_grad()
            loss = loss_fn(model(xs[i % 2]), y)
            loss.backward()
            torch.nn.utils.clip_grad_norm_(model.parameters(), 1.0)
            opt.step()
            return loss.item()
        return step
    from oracle import nets_oracle, supcon_oracle
    sd = nets_oracle.synthetic_state_dict(arch, {}, seed=0)
    sd = {k: v.to(device) for k, v in sd.items()}
    params = {k: v.clone().requires_grad_(True) for k, v in sd.items() if v.dtype.is_floating_point and "running" not in k}
    live = dict(sd)
    live.update(params)
    opt = torch.optim.Adam(list(params.values()), lr=3e-4, weight_decay=1e-4)
    p = 0.1 if arch == "phoneme_cnn" else 0.2
    chans = [32, 64, 128] if arch == "phoneme_cnn" else [64, 128, 256, 512]

    def step(i):
        drop = [torch.bernoulli(torch.full((views, c), 1 - p, device=device)) / (1 - p) for c in chans]
        emb = nets_oracle.forward(arch, live, xs[i % 2], training=True, drop=drop)
        loss = supcon_oracle.loss_torch_cpu(emb, y, temperature=T_SUPCON)
        opt.zero_grad()
        loss.backward()
        torch.nn.utils.clip_grad_norm_(list(params.values()), 1.0)
        opt.step()
        return loss.item()
    return step


def cpu_train_steps(arch, views, steps, warmup=1, threads=None, budget_s=150.0):
    """The reference's CPU training step on the host cores: -> (samples/s, threads, steps actually timed). Stops early once
    `budget_s` of timed work has accumulated, so a large --steps cannot run for an hour."""
    threads = threads or os.cpu_count()
    torch.set_num_threads(threads)
    step = _train_step_fn(arch, views, "cpu")
    times = []
    for i in range(warmup + steps):
        t0 = time.perf_counter()
        step(i)
        if i >= warmup:
            times.append(time.perf_counter() - t0)
            if sum(times) > budget_s:
                break
    return views / (sum(times) / len(times)), threads, len(times)


def _ref_mfcc():
    """(per-clip fn, batched fn) of the front end on the CPU: reference MFCCExtractor when vendored, else the oracle's torch-op port."""
    if have_reference():
        from src.datasets.features import MFCCExtractor
        ext = MFCCExtractor()
        return (lambda w: [ext(w[i:i + 1]) for i in range(w.shape[0])]), (lambda w: ext(w))
    from oracle import mfcc_oracle
    return (lambda w: mfcc_oracle.mfcc_torch_cpu(w, per_clip=True)), (lambda w: mfcc_oracle.mfcc_torch_cpu(w, per_clip=False))


def _ref_supcon():
    if have_reference():
        from src.training.losses import get_loss_fn
        fn = get_loss_fn("supervised_contrastive", temperature=T_SUPCON)
        return lambda f, y: fn(f, y)
    from oracle import supcon_oracle
    return lambda f, y: supcon_oracle.loss_torch_cpu(f, y, temperature=T_SUPCON)


def cpu_frontend(n_clips=1024, threads=None):
    """clips/s of the reference MFCCExtractor on the host: (a) one clip per call, the reference's real training behaviour
    (dataset.py:90), timed at 1 thread and at all threads (tiny per-clip ops often run faster single-threaded) and the better
    one reported; (b) one batched call with all threads (the reference's best case)."""
    per_clip_fn, batched_fn = _ref_mfcc()
    all_threads = threads or os.cpu_count()
    w = 0.1 * torch.randn(n_clips, 16000, generator=torch.Generator().manual_seed(0))
    best, best_t = 0.0, 1
    for th in sorted({1, all_threads}):
        torch.set_num_threads(th)
        per_clip_fn(w[:8])
        n = n_clips if th == 1 else max(64, n_clips // 8)
        t0 = time.perf_counter()
        per_clip_fn(w[:n])
        v = n / (time.perf_counter() - t0)
        if v > best:
            best, best_t = v, th
    torch.set_num_threads(all_threads)
    batched_fn(w[:64])
    t0 = time.perf_counter()
    batched_fn(w)
    batched = n_clips / (time.perf_counter() - t0)
    return best, batched, best_t


def cpu_supcon(n=8192, d=128, reps=2):
    fn = _ref_supcon()
    torch.set_num_threads(os.cpu_count())
    g = torch.Generator().manual_seed(0)
    fc = torch.nn.functional.normalize(torch.randn(n, d, generator=g), dim=1).requires_grad_(True)
    yc = torch.randint(0, 38, (n,), generator=g)
    fn(fc, yc).backward()
    t0 = time.perf_counter()
    for _ in range(reps):
        fc.grad = None
        fn(fc, yc).backward()
    return n / ((time.perf_counter() - t0) / reps)


def workload_config(workload, world, views=None):
    """The `config` object both arms print (identical dicts: the driver compares them)."""
    wl = WORKLOADS[workload]
    cfg = {"workload": wl["desc"]}
    if "arch" in wl:
        v = int(views or wl["views"])
        cfg.update(views_per_gpu=v, global_views=v * world)
    return cfg


def run_reference(args):
    """Reference arm: the reference's own CPU implementation of the path (oracle/_ref when vendored, else the oracle port) on
    all host cores. Rank 0 alone works; `steps` in the printed line is the number of steps ACTUALLY timed."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    wl = WORKLOADS[args.workload]
    kind = ref_kind()
    timed = args.steps
    if args.workload == "supcon_8192":
        reps = max(1, min(args.steps, 5))
        value, unit, metric, threads = cpu_supcon(reps=reps), "views/s", "supcon_fwd_bwd_views_per_sec", os.cpu_count()
        timed = reps
        sample = f"{reps} timed fwd+bwd of the same 8192 x 128 problem after 1 warm-up"
    elif args.workload == "frontend":
        per_clip, batched, threads = cpu_frontend(1024)
        value, unit = max(per_clip, batched), "clips/s"
        sample = "1024 of the 65 536 clips: one clip per call (dataset.py:90 behaviour, %d thread(s)) %.0f clips/s; one batched call (all %d threads) %.0f clips/s; value = the better" % (threads, per_clip, os.cpu_count(), batched)
        threads = os.cpu_count() if batched >= per_clip else threads
        metric = "mfcc_frontend_clips_per_sec"
        timed = 1
    else:
        views = int(args.views_per_gpu or wl["views"])
        value, threads, timed = cpu_train_steps(wl["arch"], views, max(1, args.steps), warmup=max(1, min(args.warmup, 2)))
        unit, metric = "samples/s", "supcon_train_samples_per_sec"
        sample = (f"{timed} timed step(s) of one {views}-view shard (the per-GPU batch; the global batch is {views * args.gpus} views) after "
                  f"{max(1, min(args.warmup, 2))} warm-up; a run stops early after 150 s of timed work")
    per = int(args.views_per_gpu or wl["views"]) if "arch" in wl else wl.get("n", 1)
    line = {"impl": "reference", "metric": metric, "value": value, "unit": unit, "n_gpus": args.gpus, "steps": timed, "steps_requested": args.steps,
            "warmup": args.warmup,
            "ms_per_step": (per / value * 1e3) if args.workload != "frontend" else None, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": workload_config(args.workload, args.gpus, args.views_per_gpu),
            "cpu_baseline": {"value": value, "unit": unit, "cores": threads, "kind": kind, "sample": sample},
            "e2e": {"value": value, "unit": unit, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    _emit(line)


def run_torch_cuda(args):
    """Yardstick arm (not the reference arm, not the product): the reference's modules -- or the oracle's restatement of them --
    run by stock PyTorch on cuda:0 (cuDNN / cuBLAS / ATen kernels, TF32 off so the arithmetic class matches), same step, same
    inputs. Answers "does the hand-written path beat what `device: cuda` in the reference's config would have given?"."""
    if int(os.environ.get("RANK", "0")) != 0:
        return
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cudnn.benchmark = True
    wl = WORKLOADS[args.workload]
    dev = "cuda:0"
    torch.cuda.set_device(0)
    if args.workload == "supcon_8192":
        fn = _ref_supcon()
        g = torch.Generator().manual_seed(0)
        fc = torch.nn.functional.normalize(torch.randn(8192, 128, generator=g), dim=1).to(dev).requires_grad_(True)
        yc = torch.randint(0, 38, (8192,), generator=g).to(dev)

        def step(i):
            fc.grad = None
            fn(fc, yc).backward()
        per, unit, metric = 8192, "views/s", "supcon_fwd_bwd_views_per_sec"
    elif args.workload == "frontend":
        _, batched_fn = _ref_mfcc()
        w = frontend_inputs(4096, dev)
        if have_reference():
            from src.datasets.features import MFCCExtractor
            ext = MFCCExtractor().to(dev)
            batched_fn = lambda t: ext(t)
        else:
            raise SystemExit("torch_cuda front end needs the vendored reference (oracle/_ref)")

        def step(i):
            batched_fn(w)
        per, unit, metric = 4096, "clips/s", "mfcc_frontend_clips_per_sec"
    else:
        views = int(args.views_per_gpu or wl["views"])
        fn = _train_step_fn(wl["arch"], views, dev)

        def step(i):
            fn(i)
        per, unit, metric = views, "samples/s", "supcon_train_samples_per_sec"
    for i in range(max(args.warmup, 3)):
        step(i)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(args.steps):
        step(i)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / args.steps
    _emit({"impl": "torch_cuda", "metric": metric, "value": per / (ms * 1e-3), "unit": unit, "n_gpus": 1, "steps": args.steps, "warmup": max(args.warmup, 3),
           "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
           "config": workload_config(args.workload, 1, args.views_per_gpu),
           "note": ("stock PyTorch %s on cuda:0, TF32 disabled, cudnn.benchmark on, eager; modules: %s. A yardstick, not the baseline: "
                    "library kernels (cuDNN/cuBLAS/ATen)" % (torch.__version__, "the reference's own (oracle/_ref)" if have_reference() else "oracle restatement"))})


# ------------------------------------------------------------------------------------------------ main
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference", "torch_cuda"])
    ap.add_argument("--views-per-gpu", type=int, default=None, help="override the workload's per-GPU batch (training workloads), e.g. 512 for BASELINE configs[4]")
    ap.add_argument("--workload", default="train_cnn_deep", choices=list(WORKLOADS))
    ap.add_argument("--no-also", action="store_true", help="skip the secondary workloads")
    ap.add_argument("--no-cpu", action="store_true", help="skip the CPU baseline leg")
    ap.add_argument("--no-graph", action="store_true", help="eager launches instead of whole-step CUDA-graph replay")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "b200" else args.warmup

    if args.impl == "reference":
        run_reference(args)
        return
    if args.impl == "torch_cuda":
        run_torch_cuda(args)
        return

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (no CPU fallback); use --impl reference for the CPU arm")
    from phoneme_contrast_b200.parallel import init_distributed
    parallel = init_distributed()
    rank = 0 if parallel is None else parallel.rank
    world = 1 if parallel is None else parallel.world_size
    device = f"cuda:{int(os.environ.get('LOCAL_RANK', '0'))}"
    torch.cuda.set_device(device)
    pk, pk_kind = peaks()
    wl = WORKLOADS[args.workload]

    if args.workload == "supcon_8192":
        r = bench_supcon(args.steps, args.warmup, parallel, device)
        line = {"metric": "supcon_fwd_bwd_views_per_sec", "value": r["value"], "unit": "views/s", "ms_per_step": r["ms_per_step"], "dtype": "f32",
                "roofline": {"bound": "tensor", "kernel": "sctc::supcon_bwd_tc_kernel (+ supcon_fwd_tc_kernel)", "achieved": r["tflops"], "peak": pk.get("bf16_tflops_sustained", pk["bf16_tflops"]),
                             "unit": "TFLOP/s", "frac": r["tflops"] / pk.get("bf16_tflops_sustained", pk["bf16_tflops"]),
                             "traffic": ncu_traffic("r3_supcon_full.md", ("supcon_fwd_tc_kernel", "supcon_bwd_tc_kernel")),
                             "traffic_note": "dram read+write bytes of one forward + one backward launch at N = 8192 on one GPU (profiles/r3_supcon_full.md): F is read once per kernel (4.2 MB), the N x N logits never reach HBM",
                             "note": "per-rank algorithmic FLOPs 8*N^2*D/R (forward S, backward S recompute, dF = W F counted twice as in SURVEY 8d's 6N^2D + recompute) / step time; tcgen05 kernels with the FP16x2 operand split (3 fp16 products per pair: own ceiling = peak/3)"},
                "e2e": {"value": r["e2e_value"], "unit": "views/s", "h2d_bytes_per_step": r["h2d"], "d2h_bytes_per_step": r["d2h"]},
                "gpu_launches": int(round(r["launches"] * args.steps)),
                "config": {"workload": wl["desc"], "final_loss": r["loss"], "exchange": r["exchange"], "l2": "F is 4 MB (L2-resident by design: every rank re-reads all N rows); no flush"}}
        clocks = None
    elif args.workload == "frontend":
        r = bench_frontend(args.steps, args.warmup, device)
        line = {"metric": "mfcc_frontend_clips_per_sec", "value": r["value"], "unit": "clips/s", "ms_per_step": r["ms_per_step"], "dtype": "f32",
                "roofline": {"bound": "hbm", "kernel": "frontend_kernel", "achieved": r["gbs"], "peak": pk["hbm_gbs"], "unit": "GB/s",
                             "frac": r["gbs"] / pk["hbm_gbs"], "traffic": ncu_traffic("r2f_frontend_full.md", ("frontend_kernel",)),
                             "traffic_note": "dram read+write bytes of one launch over 65 536 clips x 2 views (profiles/r2f_frontend_full.md): 6.29 GB against 6.31 GB algorithmic",
                             "peak_source": pk_kind},
                "e2e": {"value": r["e2e_value"], "unit": "clips/s", "h2d_bytes_per_step": r["h2d"], "d2h_bytes_per_step": r["d2h"], "sample": r["e2e_sample"]},
                "gpu_launches": r["launches"] * args.steps, "config": {"workload": wl["desc"], "l2": "4.19 GB in / 2.1 GB out per step, far larger than the 126 MB L2"}}
        clocks = None
    else:
        r = bench_train(args.workload, args.steps, args.warmup, parallel, device, use_graph=not args.no_graph, views=args.views_per_gpu)
        roof = roofline_from_profile(r["prof"], pk, pk_kind)
        line = {"metric": "supcon_train_samples_per_sec", "value": r["value"], "unit": "samples/s", "ms_per_step": r["ms_per_step"], "dtype": "f32",
                "roofline": roof,
                "e2e": {"value": r["e2e_value"], "unit": "samples/s", "h2d_bytes_per_step": r["h2d"], "d2h_bytes_per_step": r["d2h"],
                        "note": "ContrastiveTrainer.run_batches over pinned HOST batches: H2D of step i+1's views + labels on a side stream under step i, every step's loss copied D2H to pinned memory behind the step and read on the host while the next step runs (all K losses read inside the timed region, final synchronize included)"},
                "gpu_launches": int(round(r["launches"] * args.steps)),
                "config": workload_config(args.workload, world, args.views_per_gpu),
                "detail": {"parallelism": f"dp{world}: all_gather(embeddings, labels, row stats) + flat-bucket gradient all-reduce; per-rank BatchNorm statistics",
                           "precision": "fp32 storage and accumulate; convolutions on tcgen05 tensor cores with fp32 operands split into fp16 hi + lo (3 products per pair, ~22-bit operands: fp32-level accuracy, parity-tested at 1e-4), exact-fp32 SIMT for the Cin=1 stem",
                           "launch": "eager launches" if args.no_graph else ("whole step captured in one CUDA graph (inputs copied into static buffers each step)" if world == 1 else
                                                                            {"GraphedDPStepPeer": "whole data-parallel step captured in ONE CUDA graph; all_gathers and the split gradient all-reduce are this library's kernels over NVLink peer memory (csrc/peer.cu), no NCCL call on the step",
                                                                             "GraphedDPStep": "five captured graph segments per step with 4 NCCL calls issued eagerly between them"}.get(r["exchange"], r["exchange"])),
                           "l2": "no explicit flush: each step streams ~2 GB of activations (>> 126 MB L2); inputs rotate over 4 device buffers",
                           "step_tflops": FLOP_PER_SAMPLE[r["arch"]] * r["views"] / (r["ms_per_step"] * 1e-3) / 1e12,
                           "final_loss": r["loss"]}}
        clocks = r["clocks"]
    if rank != 0:
        if parallel is not None:
            torch.distributed.destroy_process_group()
        return

    line.update({"n_gpus": world, "steps": args.steps, "warmup": args.warmup, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
                 "data": "synthetic", "impl": "b200"})
    if clocks is not None:
        line["clocks"] = clocks

    kind = ref_kind() if (world == 1 and not args.no_cpu) else None
    ref_note = "the reference's own modules vendored in oracle/_ref" if kind == "reference" else "oracle port (same ATen ops as the reference)"
    if world == 1 and not args.no_also:
        also = {}
        try:
            if args.workload != "frontend":
                f = bench_frontend(3, 3, device)
                fe = {"value": f["value"], "unit": "clips/s", "ms_per_step": f["ms_per_step"], "workload": WORKLOADS["frontend"]["desc"],
                      "roofline": {"bound": "hbm", "kernel": "frontend_kernel", "achieved": f["gbs"], "peak": pk["hbm_gbs"], "unit": "GB/s", "frac": f["gbs"] / pk["hbm_gbs"],
                                   "traffic": ncu_traffic("r2f_frontend_full.md", ("frontend_kernel",))},
                      "e2e": {"value": f["e2e_value"], "unit": "clips/s", "sample": f["e2e_sample"]}}
                if kind:
                    pc, bt, th = cpu_frontend(512)
                    fe["cpu_baseline"] = {"value": max(pc, bt), "unit": "clips/s", "cores": os.cpu_count() if bt >= pc else th, "kind": kind,
                                          "sample": "512 clips, %s: one clip per call (%d thread(s)) %.0f clips/s; one batched call %.0f clips/s; value = the better" % (ref_note, th, pc, bt)}
                also["mfcc_frontend_clips_per_sec"] = fe
            other = "train_cnn_small" if args.workload != "train_cnn_small" else "train_cnn_deep"
            o = bench_train(other, args.steps, args.warmup, None, device, want_profile=True, use_graph=not args.no_graph)
            oe = {"value": o["value"], "unit": "samples/s", "ms_per_step": o["ms_per_step"], "e2e": o["e2e_value"], "workload": WORKLOADS[other]["desc"],
                  "roofline": roofline_from_profile(o["prof"], pk, pk_kind),
                  "step_tflops": FLOP_PER_SAMPLE[o["arch"]] * o["views"] / (o["ms_per_step"] * 1e-3) / 1e12}
            if kind:
                v, th, n = cpu_train_steps(o["arch"], o["views"], 5, warmup=1, budget_s=30.0)
                oe["cpu_baseline"] = {"value": v, "unit": "samples/s", "cores": th, "kind": kind, "sample": f"{ref_note}: 1 warm-up + {n} timed steps of the same {o['views']}-view step"}
            also[f"supcon_train_samples_per_sec[{other}]"] = oe
            # clips -> optimiser step with the device-resident dataset path (SURVEY 8f.1 / 8d's end-to-end row)
            p = bench_pipeline(max(args.steps, 20), args.warmup, device)
            also["clips_to_optimizer_step_samples_per_sec"] = p
        except Exception as e:  # secondary numbers must never sink the primary line
            also["error"] = repr(e)
        line["also"] = also

    if kind and args.workload == "supcon_8192":
        line["cpu_baseline"] = {"value": cpu_supcon(reps=2), "unit": "views/s", "cores": os.cpu_count(), "kind": kind,
                                "sample": f"2 timed fwd+bwd of the same 8192 x 128 problem, {ref_note}"}
    elif kind and args.workload == "frontend":
        per_clip, batched, threads = cpu_frontend(1024)
        line["cpu_baseline"] = {"value": max(per_clip, batched), "unit": "clips/s", "cores": os.cpu_count() if batched >= per_clip else threads, "kind": kind,
                                "sample": "1024 of the 65 536 clips, %s: one clip per call (%d thread(s)) %.0f clips/s; one batched call (all threads) %.0f clips/s; value = the better" % (ref_note, threads, per_clip, batched)}
    elif kind:
        v, threads, n = cpu_train_steps(wl["arch"], r["views"], 3, warmup=1)
        line["cpu_baseline"] = {"value": v, "unit": "samples/s", "cores": threads, "kind": kind,
                                "sample": f"{ref_note}: 1 warm-up + {n} timed steps of the same {r['views']}-view step"}
    _emit(line)
    if parallel is not None:
        torch.distributed.destroy_process_group()


def _emit(line: dict) -> None:
    """The contract is ONE JSON line on stdout. Libraries write there too (NCCL prints its version banner with plain printf),
    so main() runs with file descriptor 1 pointing at stderr and the line goes to the saved real stdout."""
    data = (json.dumps(line) + "\n").encode()
    fd = _REAL_STDOUT if _REAL_STDOUT is not None else 1
    while data:
        data = data[os.write(fd, data):]


_REAL_STDOUT = None

if __name__ == "__main__":
    sys.stdout.flush()
    _REAL_STDOUT = os.dup(1)
    os.dup2(2, 1)                 # everything else that prints to stdout (python or C) now lands on stderr
    main()
