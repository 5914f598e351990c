#!/usr/bin/env python
"""bench.py — images/sec through the SeTok tokenizer (224^2, ViT-L/14), BASELINE.json's metric.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]

Workload (config.workload): BASELINE config 2 — batch 256 synthetic 224^2 "Mondrian" images per GPU,
ViT-L/14 tower (24 layers, select_layer -2 -> 23 run), dynamic-K DPC-kNN clustering head
(C = C_tok = 1024, F = 4096, 2+2 attention layers, k = 16, threshold 0.5, min_cluster_num 64),
tokenizer only.  Random-init weights of that architecture (seeded), synthetic data.

One "step" = one pass of the whole tokenizer over one batch.  `value` is whole-job images/s with the
batch resident in HBM; `e2e` is the same through the plugin call with pinned HOST buffers, H2D of the images
and D2H of the ragged result inside the timed region.  Multi-GPU (torchrun, one rank per GPU): weak scaling,
each rank tokenises its own 256 images and the ranks repack their ragged token outputs with one all-gather
(the path's only exchange step); time is the max over ranks.

`--impl reference` times the reference's own CPU implementation of the path (the oracle port: reference
source cannot travel to the GPU box) on the host cores, on a bounded sample of the same workload.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

METRIC = "images/sec through SeTok tokenizer (224^2, ViT-L/14)"
UNIT = "images/s"
VIT = dict(hidden_size=1024, intermediate_size=4096, num_hidden_layers=24, num_attention_heads=16, image_size=224, patch_size=14)
HEAD = dict(hidden_dim=1024, token_feat_dim=1024, min_cluster_num=64, threshold=0.5, nheads=2, dim_feedforward=4096,
            inner_cluster_layers=2, intra_cluster_layers=2, mm_vision_select_layer=-2)
KNN_K = 16
BATCH = 256
# dram__bytes_read.sum + dram__bytes_write.sum per GEMM launch, mean of fc2/qkv/out_proj/fc1 (profiles/r01_ncu_summary.md)
NCU_GEMM_DRAM_BYTES_PER_LAUNCH = (804.5e6 + 494.0e6 + 363.0e6 + 631.8e6) / 4
NCU_CLUSTER_DRAM_BYTES_PER_LAUNCH = 269.0e6 + 7.5e6   # dpc_fused_kernel, dram read + write (profiles/r01_ncu_summary.md)


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return dict(hbm=float(d["hbm_gbs"]), tf_burst=float(d["bf16_tflops"]), tf_sustained=float(d.get("bf16_tflops_sustained", d["bf16_tflops"])), src="measured")
    return dict(hbm=6650.0, tf_burst=1590.0, tf_sustained=1400.0, src="fallback")


def vit_flops_per_image(layers_run: int) -> float:
    T, C, F = 257, VIT["hidden_size"], VIT["intermediate_size"]
    per_layer = 2 * T * C * 3 * C + 2 * T * C * C + 4 * T * C * F + 4 * T * T * C     # = 24TC^2 + 4T^2C for F = 4C
    return layers_run * per_layer + 2 * 256 * 588 * C


def gemm_flops_per_image(layers_run: int) -> float:
    T, C, F = 257, VIT["hidden_size"], VIT["intermediate_size"]
    return layers_run * (2 * T * C * 3 * C + 2 * T * C * C + 4 * T * C * F)


class ClockSampler:
    """nvidia-smi clocks + throttle reasons during the timed region (B200_PROFILING.md recipe)."""
    Q = "clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index: int):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "200"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.25)
        self.proc.terminate()
        sm, mx, reasons = [], None, set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[0])); mx = float(r[1])
                for n, v in zip(names, r[3:7]):
                    if v.lower().startswith("active"):
                        reasons.add(n)
            except Exception:
                pass
        busy = [s for s in sm if mx and s > 0.3 * mx] or sm
        return {"sm_mhz": statistics.median(busy) if busy else None, "sm_max_mhz": mx, "reasons": sorted(reasons), "samples": len(sm)}


# ------------------------------------------------------------------------------------------------
# reference arm / CPU baseline: the oracle port on the host cores
# ------------------------------------------------------------------------------------------------
class CpuOracle:
    """The oracle port of the reference (fp32, torch CPU, per-image head loop exactly as the reference does) with the
    bench model's seeded weights.  Built once; `run(n)` times one forward over `n` Mondrian images."""

    def __init__(self, max_images: int, seed: int = 1234):
        from oracle import setok_oracle as O
        from setok_b200.synth import mondrian_images
        self.O = O
        self.cores = os.cpu_count() or 1
        torch.set_num_threads(self.cores)
        C = VIT["hidden_size"]
        self.tp = O.make_tower_params(C, VIT["num_hidden_layers"], VIT["num_attention_heads"], VIT["patch_size"], VIT["image_size"], seed=0)
        self.hp = O.make_head_params(C, HEAD["token_feat_dim"], HEAD["dim_feedforward"], seed=0)
        self.imgs = mondrian_images(max_images, 224, seed, "cpu")
        self.noise = torch.rand(max_images, 256, generator=torch.Generator().manual_seed(seed))
        self.kw = dict(patch=14, heads=16, layers=24, select_layer=-2)

    def run(self, n: int):
        O = self.O
        with torch.no_grad():
            t0 = time.perf_counter()
            feats = O.tower_features(self.imgs[:n], self.tp, **self.kw)
            t1 = time.perf_counter()
            O.setok_forward(self.imgs[:n], self.noise[:n], self.tp, self.hp, min_cluster_num=64, threshold=0.5, k=KNN_K, feats=feats, **self.kw)
            t2 = time.perf_counter()
        return t2 - t0, {"vit_s": t1 - t0, "head_s": t2 - t1}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    orc = CpuOracle(8)
    t1, _ = orc.run(1)                                              # warms up and sizes the bounded sample
    budget = 150.0
    n = int(max(1, min(8, budget / max((args.steps + args.warmup) * t1, 1e-3))))
    for _ in range(args.warmup):
        orc.run(n)
    dt = 0.0
    for _ in range(args.steps):
        dt += orc.run(n)[0]
    v = args.steps * n / dt
    cores = orc.cores
    line = {"impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": 1e3 * dt / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
            "data": "synthetic", "config": workload_config(args.gpus, sample=n),
            "cpu_baseline": {"value": v, "unit": UNIT, "cores": cores, "kind": "port",
                             "sample": f"{n} images per step of the 256-image workload, oracle port of the reference (torch CPU fp32, {cores} threads)"},
            "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}, "gpu_launches": 0}
    print(json.dumps(line))


def workload_config(n_gpus, sample=None):
    c = {"workload": "BASELINE config 2: batch 256 synthetic 224^2 Mondrian images per GPU, ViT-L/14 (23 of 24 layers, select_layer -2), "
                     "dynamic-K DPC-kNN head C=C_tok=1024 F=4096 k=16 thr=0.5, tokenizer only",
         "batch_per_gpu": BATCH, "global_batch": BATCH * n_gpus, "parallelism": f"dp{n_gpus}",
         "l2": "inputs larger than L2: 154 MB fp32 images + >4 GB of activations streamed per step (L2 is 126 MB)"}
    if sample is not None:
        c["sample_images_per_step"] = sample
    return c


# ------------------------------------------------------------------------------------------------
# our arm
# ------------------------------------------------------------------------------------------------
def build_model(dev):
    import setok_b200
    torch.manual_seed(0)
    tok = setok_b200.SetokTokenizer("siglip-synthetic-vit-l-14", vision_config=VIT, **HEAD)
    return tok.to(dev)


def time_gemm_mix(dev, layers_run, reps=2):
    """Average launch duration of the dominant kernel (the tcgen05 GEMM) over the ViT's real launch mix, measured
    with CUDA events on the launching stream: per layer qkv / out_proj / fc1 / fc2 at M = 256*257 rows."""
    from setok_b200 import ops
    M, C, F = BATCH * 257, VIT["hidden_size"], VIT["intermediate_size"]
    g = torch.Generator(device=dev).manual_seed(0)
    a = torch.randn(M, C, device=dev, generator=g).to(torch.bfloat16)
    u = torch.randn(M, F, device=dev, generator=g).to(torch.bfloat16)
    ws = [(torch.randn(n, k, device=dev, generator=g) * k ** -0.5).to(torch.bfloat16) for n, k in ((3 * C, C), (C, C), (F, C), (C, F))]
    outs = [torch.empty(M, n, dtype=torch.bfloat16, device=dev) for n in (3 * C, C, F, C)]
    bias = [torch.zeros(n, device=dev) for n in (3 * C, C, F, C)]

    def layer():
        ops.gemm(a, ws[0], bias[0], out=outs[0])
        ops.gemm(a, ws[1], bias[1], out=outs[1], residual=outs[1])
        ops.gemm(a, ws[2], bias[2], out=outs[2], act=ops.ACT_QUICK_GELU)
        ops.gemm(u, ws[3], bias[3], out=outs[3], residual=outs[3])
    for _ in range(3):
        layer()
    torch.cuda.synchronize(dev)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    n_layers = layers_run * reps
    e0.record()
    for _ in range(n_layers):
        layer()
    e1.record()
    torch.cuda.synchronize(dev)
    ms = e0.elapsed_time(e1)
    launches = 4 * n_layers
    flops = n_layers * (2.0 * M * C * 3 * C + 2.0 * M * C * C + 4.0 * M * C * F)
    return ms / launches, flops / launches, flops / (ms * 1e-3) / 1e12


def time_tower(tok, images, reps=5):
    """The ViT tower alone (a1+a2) on the bench batch: CUDA events around `reps` forwards."""
    dev = images.device
    for _ in range(2):
        tok.image_feature_encoder(images)
    torch.cuda.synchronize(dev)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        tok.image_feature_encoder(images)
    e1.record()
    torch.cuda.synchronize(dev)
    return e0.elapsed_time(e1) / reps


def time_cluster(dev, reps=20):
    """Clustering (a3+a4) on feature-injected mixtures: achieved algorithmic HBM GB/s.  Timed the way the tokenizer runs it:
    the tower's last row pass has already added the position embedding (setok_vit_forward_pos), the fused kernel reads the
    embedded fp32 tensor once (setok_dpc_cluster_embedded).  `with_pos_ms` is the generic entry (pos add + x_pos output
    inside the kernel)."""
    from setok_b200 import ops
    from setok_b200.synth import mog_features
    N, C = 256, 1024
    feats = mog_features(BATCH, N, C, 7, dev)
    noise = torch.rand(BATCH, N, device=dev)

    def timed(fn):
        for _ in range(3):
            out = fn()
        torch.cuda.synchronize(dev)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            out = fn()
        e1.record()
        torch.cuda.synchronize(dev)
        return e0.elapsed_time(e1) / reps, out
    ms, out = timed(lambda: ops.dpc_cluster(feats, noise, (16, 16), KNN_K, 0.5, 64, embedded=True))
    ms_pos, _ = timed(lambda: ops.dpc_cluster(feats, noise, (16, 16), KNN_K, 0.5, 64))
    K = out[4].float()
    alg_bytes = BATCH * (N * C * 4 + N * 8 + N * 4 + N * 4) + float(K.sum()) * 8      # SURVEY §8d per-image figure x batch
    return ms, alg_bytes, alg_bytes / (ms * 1e-3) / 1e9, (float(K.min()), float(K.mean()), float(K.max())), ms_pos


def run_ours(args):
    import torch.distributed as dist
    from setok_b200 import _lib
    from setok_b200.dist import all_gather_ragged
    from setok_b200.synth import mondrian_images
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: setok_b200 has no CPU path (use --impl reference for the CPU arm)")
    dev = torch.device("cuda", local)
    torch.cuda.set_device(dev)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    if args.gpus != world and rank == 0:
        print(f"warning: --gpus {args.gpus} but WORLD_SIZE {world}; reporting n_gpus={world}", file=sys.stderr)

    tok = build_model(dev)
    layers_run = tok.image_feature_encoder.layers_to_run()
    images = mondrian_images(BATCH, 224, 1234 + rank, dev)
    noise = torch.rand(BATCH, 256, device=dev, generator=torch.Generator(device=dev).manual_seed(99 + rank))
    gather = world > 1 and not args.no_gather

    def step(imgs):
        rt, idx, score = tok(imgs, k=KNN_K, noise=noise)
        if gather:
            rt = all_gather_ragged(rt)
        return rt, idx, score

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    for _ in range(max(args.warmup, 3)):
        rt, idx, score = step(images)
    barrier()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    l0 = _lib.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record()
    for _ in range(args.steps):
        rt, idx, score = step(images)
    e1.record()
    barrier()
    ms = e0.elapsed_time(e1)
    launches = _lib.launch_count() - l0
    clocks = sampler.stop() if rank == 0 else None
    t = torch.tensor([ms], device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = float(t.item())
    value = world * BATCH * args.steps / (ms * 1e-3)
    counts = (rt.offsets[1:] - rt.offsets[:-1]).float()

    # ---- e2e: pinned host images in, ragged result out, per step -------------------------------------------
    host_images = images.cpu().pin_memory()
    host_noise = noise.cpu().pin_memory()
    h2d = host_images.numel() * 4 + host_noise.numel() * 4

    from setok_b200.pipeline import stream_tokenize

    def e2e_run(n_steps):
        """n_steps batches through the public streaming API: each step copies its images+noise in from pinned host memory
        and its ragged result (tokens, offsets, labels, scores) back out."""
        post = (lambda r, i, s_: (all_gather_ragged(r), i, s_)) if gather else None
        nbytes = 0
        for res in stream_tokenize(tok, ((host_images, host_noise) for _ in range(n_steps)), post=post, k=KNN_K):
            nbytes = res.nbytes
        return nbytes

    d2h = e2e_run(3)
    barrier()
    e0.record()
    d2h = e2e_run(args.steps)
    e1.record()
    barrier()
    ms_e = e0.elapsed_time(e1)
    t = torch.tensor([ms_e], device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    e2e_value = world * BATCH * args.steps / (float(t.item()) * 1e-3)

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    pk = peaks()
    vit_ms = time_tower(tok, images)
    vit_tf = BATCH * vit_flops_per_image(layers_run) / (vit_ms * 1e-3) / 1e12
    cl_ms, cl_bytes, cl_gbs, kstats, cl_ms_pos = time_cluster(dev, reps=20)
    gemm_ms, gemm_flops, gemm_tf = time_gemm_mix(dev, layers_run)
    step_ms = ms / args.steps
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
        "ms_per_step": step_ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "bf16",
        "data": "synthetic", "config": dict(workload_config(world), ragged_all_gather=bool(gather),
                                            k_per_image={"min": float(counts.min()), "mean": float(counts.mean()), "max": float(counts.max())}),
        "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h},
        "gpu_launches": int(launches),
        "clocks": clocks,
        "roofline": {"kernel": "gemm_bf16_tcgen05_kernel (ViT layer launch mix: qkv/out_proj/fc1/fc2 at M=65792)", "bound": "tensor",
                     "achieved": gemm_tf, "peak": pk["tf_sustained"], "unit": "TFLOP/s", "frac": gemm_tf / pk["tf_sustained"],
                     "traffic": NCU_GEMM_DRAM_BYTES_PER_LAUNCH, "traffic_source": "ncu --set full dram__bytes_read+write, mean over the 4 shapes (profiles/r01_ncu_summary.md); algorithmic operand bytes are 607e6",
                     "flops_per_launch": gemm_flops, "ms_per_launch": gemm_ms,
                     "peak_source": f"{pk['src']} sustained bf16 (kernel timed inside a long loop)",
                     "step_share": (4 * layers_run * gemm_ms) / step_ms,
                     "vit_tensor_frac_of_step": (BATCH * vit_flops_per_image(layers_run) / (step_ms * 1e-3) / 1e12) / pk["tf_sustained"]},
        "roofline_vit": {"kernel": "whole ViT-L/14 tower (im2col, patch GEMM, 23 x [LN, qkv, attention, out_proj, LN, fc1, fc2])", "bound": "tensor",
                         "achieved": vit_tf, "peak": pk["tf_sustained"], "unit": "TFLOP/s", "frac": vit_tf / pk["tf_sustained"],
                         "ms": vit_ms, "flops": BATCH * vit_flops_per_image(layers_run)},
        "roofline_cluster": {"kernel": "dpc_fused_kernel (a4 on the position-embedded tensor; a3 is fused into the tower's last row pass), B=256 N=256 C=1024 feature-injected", "bound": "hbm",
                             "achieved": cl_gbs, "peak": pk["hbm"], "unit": "GB/s", "frac": cl_gbs / pk["hbm"], "traffic": NCU_CLUSTER_DRAM_BYTES_PER_LAUNCH,
                             "bytes_per_launch": cl_bytes, "ms_per_launch": cl_ms, "with_pos_ms": cl_ms_pos, "k_min_mean_max": kstats,
                             "tensor_frac_if_compute": (BATCH * 2.0 * 256 * 256 * 1024 / (cl_ms * 1e-3) / 1e12) / pk["tf_burst"],
                             "tensor_frac_executed": (4 * BATCH * 2.0 * 256 * 256 * 1024 / (cl_ms * 1e-3) / 1e12) / pk["tf_burst"],
                             "note": "the Gram runs as an exact 4-term bf16 hi/lo split on tcgen05 (4x the algorithmic FLOPs); the 256x256 fp32 distance matrix fills TMEM, so the MMA phase (48 us/image) and the select phase (21 us/image) serialise; 256 images = 2 waves on 148 SMs"},
    }
    if world == 1 and not args.no_cpu:
        n = args.cpu_sample
        orc = CpuOracle(n)
        orc.run(1)
        secs, stages = orc.run(n)
        v = n / secs
        cores = orc.cores
        line["cpu_baseline"] = {"value": v, "unit": UNIT, "cores": cores, "kind": "port",
                                "sample": f"{n} of the 256 images (same generator), oracle port of the reference, torch CPU fp32, "
                                          f"{cores} threads: ViT {stages['vit_s']:.1f}s + per-image head loop {stages['head_s']:.1f}s"}
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-gather", action="store_true", help="skip the ragged all-gather at N>1")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--cpu-sample", type=int, default=8)
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
