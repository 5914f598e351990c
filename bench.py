#!/usr/bin/env python
"""bench.py — images/sec through the SeTok tokenizer (224^2, ViT-L/14), BASELINE.json's metric.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--config {2,3,4,5}] [--impl reference]

Default workload (config.workload): BASELINE config 2 — batch 256 synthetic 224^2 "Mondrian" images per GPU,
ViT-L/14 tower (24 layers, select_layer -2 -> 23 run), dynamic-K DPC-kNN clustering head
(C = C_tok = 1024, F = 4096, 2+2 attention layers, k = 16, threshold 0.5, min_cluster_num 64),
tokenizer only.  Random-init weights of that architecture (seeded), synthetic data.  `--config 3|4|5` time the other
BASELINE configurations (336^2 + reconstruction decoder; encode_images -> Vicuna projector; mixed-resolution batch).

One "step" = one pass of the whole path over one batch.  `value` is whole-job images/s with the batch resident in HBM;
`e2e` is the same through the plugin call with pinned HOST buffers, H2D of the images and D2H of the ragged result
inside the timed region.  Multi-GPU (torchrun, one rank per GPU): weak scaling, each rank processes its own batch and the
ranks repack their ragged token outputs with the two-phase all-gather (the path's only exchange step; its row transfer
overlaps the next step's tower); time is the max over ranks.

`--impl reference` times the reference's own CPU implementation of the path (the oracle port: reference
source cannot travel to the GPU box) on the host cores, on a bounded sample of the same workload.
"""
from __future__ import annotations

import argparse
import json
import math
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
DETOK = dict(token_feat_dim=1024, hidden_dim=768, patch_size=14, image_size=336, decoder_embed_dim=768, decoder_nheads=12, decoder_depth=16,
             num_hidden_layers=6, cross_attention_freq=2)
KNN_K = 16
BATCH = 256
SEED = 1234
# dram__bytes_read.sum + dram__bytes_write.sum per GEMM launch, mean of fc2/qkv/out_proj/fc1 with the f32 residual stream and the
# LayerNorms folded into them (profiles/r02_ncu_summary.md); the algorithmic operand bytes of the same mix are 815e6
NCU_GEMM_DRAM_BYTES_PER_LAUNCH = (1284.8e6 + 501.9e6 + 765.0e6 + 642.1e6) / 4   # fc2 / qkv / out_proj / fc1, LayerNorm-fold variants (profiles/r02_ncu_summary.md)
NCU_CLUSTER_DRAM_BYTES_PER_LAUNCH = 268.8e6 + 6.1e6   # dpc_fused_kernel, dram read + write (profiles/r02_ncu_summary.md)

WORKLOADS = {
    2: "BASELINE config 2: batch 256 synthetic 224^2 Mondrian images per GPU, ViT-L/14 (23 of 24 layers, select_layer -2), "
       "dynamic-K DPC-kNN head C=C_tok=1024 F=4096 k=16 thr=0.5, tokenizer only",
    3: "BASELINE config 3: batch 128 synthetic 336^2 Mondrian images per GPU (576 patches), bf16, ViT-L/14 tokenizer as config 2 + "
       "reconstruction decoder (Q-Former 6 layers / cross-attention every 2, hidden 768; 16 ViT blocks of 768, 576 queries)",
    4: "BASELINE config 4: Setokim encode_images -> mlp2x_gelu projector 1024 -> 4096 -> 4096 (Vicuna-7B width), 64 synthetic 224^2 "
       "Mondrian images per GPU, bf16, data parallel with the ragged all-gather of the projected rows",
    5: "BASELINE config 5: mixed-resolution ragged batch, 96 Mondrian images per GPU in equal thirds of 224^2 / 336^2 / 448^2, "
       "ViT-L/14 with bicubically resized position table, images dealt to ranks by N^2 cost, ragged all-gather in image order",
}
BATCHES = {2: 256, 3: 128, 4: 64, 5: 96}


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return dict(hbm=float(d["hbm_gbs"]), tf_burst=float(d["bf16_tflops"]), tf_sustained=float(d.get("bf16_tflops_sustained", d["bf16_tflops"])), src="measured")
    return dict(hbm=6650.0, tf_burst=1590.0, tf_sustained=1400.0, src="fallback")


def vit_flops_per_image(layers_run: int, size: int = 224) -> float:
    N = (size // 14) ** 2
    T, C, F = N + 1, VIT["hidden_size"], VIT["intermediate_size"]
    per_layer = 2 * T * C * 3 * C + 2 * T * C * C + 4 * T * C * F + 4 * T * T * C     # = 24TC^2 + 4T^2C for F = 4C
    return layers_run * per_layer + 2 * N * 588 * C


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
# synthetic inputs: generated on the HOST once, so that the CPU arm and the GPU arm consume the same tensors
# ------------------------------------------------------------------------------------------------
_HOST_BATCHES = {}


def host_batch(config: int, rank: int, n: int = None):
    """(uint8 pixels, float images, noise) of one rank's batch, generated once per process (the CPU arm takes a prefix of
    the very tensors the GPU arm uploads).  The float images are the processor's output for the uint8 pixels (rescale +
    normalize), so the uint8 end-to-end path and the float paths see identical pixels.  Config 5 returns lists (one entry
    per image, three resolutions)."""
    key = (config, rank, n)
    if key not in _HOST_BATCHES:
        _HOST_BATCHES[key] = _make_host_batch(config, rank, n)
    return _HOST_BATCHES[key]


def _make_host_batch(config: int, rank: int, n: int = None):
    from setok_b200.synth import mondrian_u8, normalize_u8
    B = n or BATCHES[config]
    if config == 5:
        sizes = ([224] * (B // 3) + [336] * (B // 3) + [448] * (B - 2 * (B // 3)))
        u8 = [mondrian_u8(1, s, SEED + 7 * i + 1000 * rank, g_min=16, g_max=96)[0] for i, s in enumerate(sizes)]
        imgs = [normalize_u8(u[None])[0] for u in u8]
        noise = [torch.rand((s // 14) ** 2, generator=torch.Generator().manual_seed(99 + i + 1000 * rank)) for i, s in enumerate(sizes)]
        return u8, imgs, noise
    size = 336 if config == 3 else 224
    u8 = mondrian_u8(B, size, SEED + rank)
    imgs = normalize_u8(u8)
    noise = torch.rand(B, (size // 14) ** 2, generator=torch.Generator().manual_seed(99 + rank))
    return u8, imgs, noise


# ------------------------------------------------------------------------------------------------
# reference arm / CPU baseline: the oracle port on the host cores
# ------------------------------------------------------------------------------------------------
class CpuOracle:
    """The oracle port of the reference (fp32, torch CPU, per-image head loop exactly as the reference does) with the
    bench model's architecture and seeded weights.  `run(n)` times one forward over the first `n` images of the batch."""

    def __init__(self, config: int, max_images: int):
        from oracle import setok_oracle as O
        self.O, self.config = O, config
        self.cores = os.cpu_count() or 1
        torch.set_num_threads(self.cores)
        C = VIT["hidden_size"]
        size = 336 if config == 3 else 224
        self.size = size
        self.tp = O.make_tower_params(C, VIT["num_hidden_layers"], VIT["num_attention_heads"], VIT["patch_size"], size if config == 3 else 224, seed=0)
        self.hp = O.make_head_params(C, HEAD["token_feat_dim"], HEAD["dim_feedforward"], seed=0)
        _, self.imgs, self.noise = host_batch(config, 0)      # rank 0's full batch; run(n) takes its first n images
        self.kw = dict(patch=14, heads=16, layers=24, select_layer=-2)
        if config == 3:
            from oracle import detok_oracle as D
            self.D = D
            self.dp = D.make_detok_params(token_dim=C, hidden=768, q_heads=12, q_inter=3072, q_layers=6, cross_freq=2, grid=24, dec_dim=768,
                                          dec_depth=16, dec_mlp=3072, seed=0)
        if config == 4:
            self.pp = O.make_projector_params(C, 4096, "mlp2x_gelu", seed=0)

    def run(self, n: int):
        O = self.O
        with torch.no_grad():
            t0 = time.perf_counter()
            if self.config == 5:
                idx = [0, len(self.imgs) // 2, len(self.imgs) - 1][:max(1, min(3, n))]      # one image of each resolution
                for i in idx:
                    f = O.tower_features(self.imgs[i][None], self.tp, interpolate_pos_encoding=True, **self.kw)
                    O.tokenizer_head(f[0], self.noise[i], self.hp, min_cluster_num=64, threshold=0.5, k=KNN_K)
                t1 = t2 = time.perf_counter()
                return t2 - t0, {"vit_s": t1 - t0, "head_s": 0.0, "images": len(idx), "feats": None}
            feats = O.tower_features(self.imgs[:n], self.tp, **self.kw)
            t1 = time.perf_counter()
            toks = []
            for b in range(n):
                tk = O.tokenizer_head(feats[b], self.noise[b], self.hp, min_cluster_num=64, threshold=0.5, k=KNN_K)[0]
                if self.config == 4:
                    tk = O.projector(tk, self.pp, "mlp2x_gelu")
                toks.append(tk)
            if self.config == 3:
                offs = [0]
                for tk in toks:
                    offs.append(offs[-1] + tk.shape[0])
                x, m = self.D.pad_ragged(torch.cat(toks, 0), offs)
                self.D.detok_forward(self.dp, x, m, q_heads=12, q_layers=6, cross_freq=2, grid=24, dec_heads=12, dec_depth=16, hidden=768)
            t2 = time.perf_counter()
        return t2 - t0, {"vit_s": t1 - t0, "head_s": t2 - t1, "images": n, "feats": feats}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cap = {2: 8, 3: 2, 4: 8, 5: 3}[args.config]
    orc = CpuOracle(args.config, cap)
    t1, st = orc.run(1)                                              # warms up and sizes the bounded sample
    per_img = t1 / st["images"]
    budget = 150.0
    n = int(max(1, min(cap, budget / max((args.steps + args.warmup) * per_img, 1e-3))))
    for _ in range(args.warmup):
        orc.run(n)
    dt, imgs = 0.0, 0
    for _ in range(args.steps):
        d, st = orc.run(n)
        dt += d
        imgs += st["images"]
    v = imgs / dt
    cores = orc.cores
    line = {"impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": 1e3 * dt / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
            "data": "synthetic", "config": workload_config(args.config, args.gpus, sample=imgs // args.steps),
            "cpu_baseline": {"value": v, "unit": UNIT, "cores": cores, "kind": "port",
                             "sample": f"{imgs // args.steps} images per step of the {BATCHES[args.config]}-image workload, oracle port of the reference (torch CPU fp32, {cores} threads)"},
            "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}, "gpu_launches": 0}
    print(json.dumps(line))


def workload_config(config, n_gpus, sample=None):
    B = BATCHES[config]
    c = {"workload": WORKLOADS[config], "batch_per_gpu": B, "global_batch": B * n_gpus, "parallelism": f"dp{n_gpus}",
         "l2": "inputs larger than L2: every step streams its whole image batch plus > 2 GB of activations (L2 is 126 MB)"}
    if sample is not None:
        c["sample_images_per_step"] = sample
    return c


# ------------------------------------------------------------------------------------------------
# our arm
# ------------------------------------------------------------------------------------------------
def build_model(dev, size: int = 224):
    import setok_b200
    torch.manual_seed(0)
    tok = setok_b200.SetokTokenizer("siglip-synthetic-vit-l-14", vision_config=dict(VIT, image_size=size), **HEAD)
    return tok.to(dev)


def cuda_time(fn, reps, warm=2):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


def cuda_time_graph(fn, reps, warm=3):
    """Device time per call of `fn` with the host out of the picture: `reps` back-to-back calls are captured into one CUDA
    graph and the replay is timed with CUDA events (the launches, their stream order and their arguments are exactly those of
    the eager calls).  Falls back to eager timing if the capture is refused."""
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    try:
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            for _ in range(reps):
                fn()
        g.replay()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        g.replay()
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / reps, "cuda-graph replay"
    except Exception as exc:      # pragma: no cover - depends on the driver
        torch.cuda.synchronize()
        return cuda_time(fn, reps, warm=1), f"eager loop (graph capture failed: {type(exc).__name__})"


def cuda_time_single(fn, n=9, idle_s=0.01):
    """Median device time of ONE call issued after the GPU has been idle (boost clocks, as under ncu): events bracket a single call."""
    ts = []
    for _ in range(n):
        torch.cuda.synchronize()
        time.sleep(idle_s)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    return statistics.median(ts)


def time_gemm_mix(dev, layers_run, rows, reps=2):
    """Average launch duration of the dominant kernel (the tcgen05 GEMM) over the ViT's real launch mix, measured
    with CUDA events on the launching stream: per layer qkv / out_proj / fc1 / fc2 at M = batch * T rows, f32 residual
    stream (the tower's default)."""
    from setok_b200 import ops
    M, C, F = rows, VIT["hidden_size"], VIT["intermediate_size"]
    g = torch.Generator(device=dev).manual_seed(0)
    a = torch.randn(M, C, device=dev, generator=g).to(torch.bfloat16)
    u = torch.randn(M, F, device=dev, generator=g).to(torch.bfloat16)
    x = torch.randn(M, C, device=dev, generator=g)
    ws = [(torch.randn(n, k, device=dev, generator=g) * k ** -0.5).to(torch.bfloat16) for n, k in ((3 * C, C), (C, C), (F, C), (C, F))]
    o_qkv = torch.empty(M, 3 * C, dtype=torch.bfloat16, device=dev)
    bias = [torch.zeros(n, device=dev) for n in (3 * C, C, F, C)]
    # the tower's own variants (SETOK_VIT_LN_FOLD): qkv / fc1 finish the LayerNorm in their epilogue, out_proj / fc2 emit xhat + records
    xhat, rec_a = ops.ln_fold_init(x)
    rec_b = ops.ln_records(M, C, dev)
    s_qkv, s_fc1 = ws[0].float().sum(1).contiguous(), ws[2].float().sum(1).contiguous()
    w_small = [w * 0.05 for w in ws]                          # keeps the in-place stream bounded over the timed launches

    def layer():
        ops.gemm_ln(xhat, w_small[0], bias[0], rec_a, ln_C=C, ln_s=s_qkv, out=o_qkv)
        ops.gemm_ln(a, w_small[1], bias[1], rec_a, ln_C=C, residual=x, rec_out=rec_b, xhat=xhat, out=x)
        ops.gemm_ln(xhat, w_small[2], bias[2], rec_b, ln_C=C, ln_s=s_fc1, act=ops.ACT_QUICK_GELU, out=u)
        ops.gemm_ln(u, w_small[3], bias[3], rec_b, ln_C=C, residual=x, rec_out=rec_a, xhat=xhat, out=x)
    n_layers = layers_run * reps
    ms = cuda_time(lambda: [layer() for _ in range(layers_run)], reps, warm=1) * reps
    launches = 4 * n_layers
    flops = n_layers * (2.0 * M * C * 3 * C + 2.0 * M * C * C + 4.0 * M * C * F)
    return ms / launches, flops / launches, flops / (ms * 1e-3) / 1e12


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
    ms, how = cuda_time_graph(lambda: ops.dpc_cluster(feats, noise, (16, 16), KNN_K, 0.5, 64, embedded=True), reps)
    ms_eager = cuda_time(lambda: ops.dpc_cluster(feats, noise, (16, 16), KNN_K, 0.5, 64, embedded=True), reps, warm=1)
    out = ops.dpc_cluster(feats, noise, (16, 16), KNN_K, 0.5, 64, embedded=True)
    ms_single = cuda_time_single(lambda: ops.dpc_cluster(feats, noise, (16, 16), KNN_K, 0.5, 64, embedded=True))
    ms_pos, _ = cuda_time_graph(lambda: ops.dpc_cluster(feats, noise, (16, 16), KNN_K, 0.5, 64), reps)
    K = out[4].float()
    alg_bytes = BATCH * (N * C * 4 + N * 8 + N * 4 + N * 4) + float(K.sum()) * 8      # SURVEY §8d per-image figure x batch
    return ms, alg_bytes, alg_bytes / (ms * 1e-3) / 1e9, (float(K.min()), float(K.mean()), float(K.max())), ms_pos, ms_eager, how, ms_single


def gpu_eager_baseline(dev, tok, images, noise, head_sample=16):
    """The only pre-existing GPU implementation of the path: torch eager on the same B200 -- transformers'
    CLIPVisionModel in bf16 with SDPA attention holding the SAME weights, then the reference's per-image head loop (the
    oracle port executed on the device, fp32).  Tower on the full batch; the head loop on `head_sample` images (it is
    launch- and sync-bound, so it scales linearly)."""
    from transformers import CLIPVisionConfig, CLIPVisionModel
    from oracle import setok_oracle as O
    cfg = CLIPVisionConfig(**dict(VIT, image_size=images.shape[-1]), attn_implementation="sdpa")
    hf = CLIPVisionModel(cfg)
    hf.load_state_dict(tok.image_feature_encoder.vision_tower.state_dict())
    hf = hf.to(dev, torch.bfloat16).eval()
    xb = images.to(torch.bfloat16)

    def tower():
        with torch.no_grad():
            return hf(xb, output_hidden_states=True).hidden_states[-2][:, 1:]
    ms_tower = cuda_time(tower, 3, warm=2)
    feats = tower().float()
    hp = {k: v.detach().float() for k, v in tok.state_dict().items() if not k.startswith("image_feature_encoder")}
    n = min(head_sample, images.shape[0])
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    with torch.no_grad(), torch.device(dev):                      # the oracle's factory calls (arange, zeros ...) land on the device
        for b in range(n):
            O.tokenizer_head(feats[b], noise[b], hp, min_cluster_num=64, threshold=0.5, k=KNN_K)
    torch.cuda.synchronize()
    ms_head_per_image = (time.perf_counter() - t0) * 1e3 / n
    B = images.shape[0]
    del hf
    torch.cuda.empty_cache()
    return {"what": "torch eager on this GPU: transformers CLIPVisionModel bf16 (sdpa) with the same weights + the reference's per-image head "
                    "loop in torch fp32 on the device",
            "tower_images_per_s": B / (ms_tower * 1e-3), "tower_ms": ms_tower, "head_ms_per_image": ms_head_per_image, "head_sample_images": n,
            "whole_path_images_per_s": B / ((ms_tower + B * ms_head_per_image) * 1e-3)}


def k32_variant(tok, images, noise, steps=5):
    """Config 2 with a threshold tuned so that the Mondrian images give K ~ 32 tokens per image (the ragged side at the size
    SURVEY.md sized the head / projector / all-gather for): the threshold is searched on the batch itself."""
    best = None
    for thr in (0.4, 0.3, 0.25, 0.2, 0.15, 0.1, 0.07, 0.05):
        rt, _, _ = tok(images, k=KNN_K, noise=noise, threshold=thr)
        kb = float((rt.offsets[1:] - rt.offsets[:-1]).float().mean())
        if best is None or abs(kb - 32) < abs(best[1] - 32):
            best = (thr, kb)
    thr, kb = best
    ms = cuda_time(lambda: tok(images, k=KNN_K, noise=noise, threshold=thr), steps, warm=2)
    return {"threshold": thr, "k_mean": kb, "ms_per_step": ms, "images_per_s": images.shape[0] / (ms * 1e-3)}


class Harness:
    def __init__(self, args):
        import torch.distributed as dist
        self.dist = dist
        self.world = int(os.environ.get("WORLD_SIZE", "1"))
        self.rank = int(os.environ.get("RANK", "0"))
        self.local = int(os.environ.get("LOCAL_RANK", "0"))
        if not torch.cuda.is_available():
            raise SystemExit("bench.py needs a CUDA device: setok_b200 has no CPU path (use --impl reference for the CPU arm)")
        self.dev = torch.device("cuda", self.local)
        torch.cuda.set_device(self.dev)
        if self.world > 1:
            dist.init_process_group("nccl", device_id=self.dev)
        if args.gpus != self.world and self.rank == 0:
            print(f"warning: --gpus {args.gpus} but WORLD_SIZE {self.world}; reporting n_gpus={self.world}", file=sys.stderr)
        self.args = args

    def barrier(self):
        if self.world > 1:
            self.dist.barrier()
        torch.cuda.synchronize(self.dev)

    def max_over_ranks(self, ms):
        t = torch.tensor([ms], device=self.dev)
        if self.world > 1:
            self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
        return float(t.item())

    def timed(self, run_steps):
        """run_steps(n) launches n steps (and everything they entail) on the current stream / its side streams and returns when
        all of it has been ENQUEUED or completed; the timed region is bracketed by barrier + synchronize on both sides."""
        from setok_b200 import _lib
        args = self.args
        run_steps(max(args.warmup, 3))
        self.barrier()
        sampler = ClockSampler(self.local)
        if self.rank == 0:
            sampler.start()
        l0 = _lib.launch_count()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        self.barrier()
        e0.record()
        run_steps(args.steps)
        e1.record()
        self.barrier()
        ms = self.max_over_ranks(e0.elapsed_time(e1))
        launches = _lib.launch_count() - l0
        clocks = sampler.stop() if self.rank == 0 else None
        return ms, launches, clocks

    def timed_e2e(self, run_steps):
        args = self.args
        # warm-up with as many steps as the timed region: the pipeline keeps several batches in flight, and the pinned-host /
        # device caching allocators must have seen that depth before the clock starts (a first cudaHostAlloc blocks the device)
        nbytes = run_steps(max(3, args.warmup, args.steps))
        self.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        nbytes = run_steps(args.steps)
        e1.record()
        self.barrier()
        return self.max_over_ranks(e0.elapsed_time(e1)), nbytes

    def finish(self):
        if self.world > 1:
            self.dist.destroy_process_group()


def pin(t):
    return [x.pin_memory() for x in t] if isinstance(t, list) else t.pin_memory()


def run_config2(args):
    from setok_b200.dist import RaggedAllGather
    from setok_b200.pipeline import stream_tokenize
    H = Harness(args)
    dev, world, rank = H.dev, H.world, H.rank
    tok = build_model(dev)
    layers_run = tok.image_feature_encoder.layers_to_run()
    u8, imgs_h, noise_h = host_batch(2, rank)
    images, noise = imgs_h.to(dev), noise_h.to(dev)
    gather = RaggedAllGather(BATCH, device=dev) if (world > 1 and not args.no_gather) else None
    last = {}

    def run_steps(n):
        pending = None
        for _ in range(n):
            rt, idx, score = tok(images, k=KNN_K, noise=noise)
            if gather is not None:
                h = gather.start(rt)                 # header exchange behind this step's kernels
                if pending is not None:
                    last["rt"] = gather.finish(pending)   # previous step's row exchange overlaps this step's tower
                pending = h
            else:
                last["rt"] = rt
        if pending is not None:
            last["rt"] = gather.finish(pending)
        last["local"] = rt

    ms, launches, clocks = H.timed(run_steps)
    value = world * BATCH * args.steps / (ms * 1e-3)
    counts = (last["local"].offsets[1:] - last["local"].offsets[:-1]).float()

    # ---- e2e: pinned host uint8 pixels in, ragged result out, per step ---------------------------------------
    h_u8, h_f32, h_noise = pin(u8), pin(imgs_h), pin(noise_h)

    def e2e_runner(src):
        def run(n_steps):
            nbytes = 0
            for res in stream_tokenize(tok, ((src, h_noise) for _ in range(n_steps)), gather=gather, k=KNN_K):
                nbytes = res.nbytes
            return nbytes
        return run
    ms_e, d2h = H.timed_e2e(e2e_runner(h_u8))
    e2e_value = world * BATCH * args.steps / (ms_e * 1e-3)
    ms_f, _ = H.timed_e2e(e2e_runner(h_f32))
    e2e_f32 = world * BATCH * args.steps / (ms_f * 1e-3)
    # the resident step once more, now in the power / thermal state the e2e legs ran in (they come ~1 s later than `value`'s
    # timed region): separates what the host<->device streaming costs from what the later, warmer window costs
    ms_again, _, _ = H.timed(run_steps)
    value_again = world * BATCH * args.steps / (ms_again * 1e-3)
    if rank != 0:
        H.finish()
        return
    pk = peaks()
    vit_ms = cuda_time(lambda: tok.image_feature_encoder(images), 5)
    vit_tf = BATCH * vit_flops_per_image(layers_run) / (vit_ms * 1e-3) / 1e12
    cl_ms, cl_bytes, cl_gbs, kstats, cl_ms_pos, cl_ms_eager, cl_how, cl_ms_single = time_cluster(dev, reps=20)
    gemm_ms, gemm_flops, gemm_tf = time_gemm_mix(dev, layers_run, BATCH * 257)
    step_ms = ms / args.steps
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
        "ms_per_step": step_ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "bf16",
        "data": "synthetic", "config": dict(workload_config(2, world), ragged_all_gather=gather is not None,
                                            tower_residual_stream="f32", tower_layernorm="folded into the qkv / out_proj / fc1 / fc2 GEMM epilogues (SETOK_VIT_LN_FOLD)", k_per_image={"min": float(counts.min()), "mean": float(counts.mean()), "max": float(counts.max())}),
        "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h_u8.numel() + h_noise.numel() * 4, "d2h_bytes_per_step": d2h,
                "input": "uint8 pixels (B,3,224,224) from pinned host memory; rescale + normalize inside the patch-embedding pass",
                "float32_input_value": e2e_f32, "float32_input_h2d_bytes_per_step": h_f32.numel() * 4 + h_noise.numel() * 4,
                "resident_value_remeasured_after_e2e": value_again,
                "note": "e2e is timed ~1 s into continuous load, `value` 0.1-0.6 s into it; resident_value_remeasured_after_e2e is the HBM-resident step "
                        "timed again right after the e2e legs, i.e. in their power state: the streaming itself costs e2e vs that number"},
        "gpu_launches": int(launches),
        "clocks": clocks,
        "roofline": {"kernel": "gemm_bf16_tcgen05_kernel (ViT layer launch mix: qkv/out_proj/fc1/fc2 at M=65792, f32 residual stream, LayerNorms folded into the epilogues)", "bound": "tensor",
                     "achieved": gemm_tf, "peak": pk["tf_sustained"], "unit": "TFLOP/s", "frac": gemm_tf / pk["tf_sustained"],
                     "traffic": NCU_GEMM_DRAM_BYTES_PER_LAUNCH, "traffic_source": "ncu --set full dram__bytes_read+write, mean over the 4 shapes (profiles/r02_ncu_summary.md); algorithmic operand bytes are 815e6 with the f32 residual stream and the xhat / row-record outputs of the folded LayerNorms",
                     "note": "the launches carry the layers' LayerNorm work (SETOK_VIT_LN_FOLD: xhat + row sums written by out_proj / fc2, normalisation finished in the qkv / fc1 epilogues): "
                             "per launch they are ~6 % slower than the plain GEMM variants, and the 46 LayerNorm passes per step are gone (tower -3.3 %, tools/bench_ln_fold.py)",
                     "flops_per_launch": gemm_flops, "ms_per_launch": gemm_ms,
                     "peak_source": f"{pk['src']} sustained bf16 (kernel timed inside a long loop)",
                     "step_share": (4 * layers_run * gemm_ms) / step_ms,
                     "vit_tensor_frac_of_step": (BATCH * vit_flops_per_image(layers_run) / (step_ms * 1e-3) / 1e12) / pk["tf_sustained"]},
        "roofline_vit": {"kernel": "whole ViT-L/14 tower (im2col, patch GEMM, pre-LN, 23 x [qkv, attention, out_proj, fc1, fc2] with the LayerNorms folded into the GEMMs)", "bound": "tensor",
                         "achieved": vit_tf, "peak": pk["tf_sustained"], "unit": "TFLOP/s", "frac": vit_tf / pk["tf_sustained"],
                         "ms": vit_ms, "flops": BATCH * vit_flops_per_image(layers_run)},
        "roofline_cluster": {"kernel": "dpc_fused_kernel (a4 on the position-embedded tensor; a3 is fused into the tower's last row pass), B=256 N=256 C=1024 feature-injected", "bound": "hbm",
                             "achieved": cl_gbs, "peak": pk["hbm"], "unit": "GB/s", "frac": cl_gbs / pk["hbm"], "traffic": NCU_CLUSTER_DRAM_BYTES_PER_LAUNCH,
                             "bytes_per_launch": cl_bytes, "ms_per_launch": cl_ms, "timing": cl_how + " of the clustering call (dpc_fused_kernel + the offsets scan)",
                             "ms_per_call_eager_loop": cl_ms_eager, "ms_single_call_after_idle": cl_ms_single,
                             "frac_single_call_after_idle": (cl_bytes / (cl_ms_single * 1e-3) / 1e9) / pk["hbm"],
                             "note": "achieved / frac: back-to-back calls (the SM clock settles at the power-capped level of the rest of the step); "
                                     "the single call after idle runs at boost clocks like the ncu capture (profiles/r02_ncu_summary.md: 130 us, 0.32)", "with_pos_ms": cl_ms_pos, "k_min_mean_max": kstats,
                             "tensor_frac_if_compute": (BATCH * 2.0 * 256 * 256 * 1024 / (cl_ms * 1e-3) / 1e12) / pk["tf_burst"],
                             "tensor_frac_executed": (3 * BATCH * 2.0 * 256 * 256 * 1024 / (cl_ms * 1e-3) / 1e12) / pk["tf_burst"]},
    }
    if world == 1 and not args.no_extras:
        line["k32_variant"] = k32_variant(tok, images, noise)
        line["gpu_eager_baseline"] = gpu_eager_baseline(dev, tok, images, noise)
        line["gpu_eager_baseline"]["speedup_tower"] = (BATCH / (vit_ms * 1e-3)) / line["gpu_eager_baseline"]["tower_images_per_s"]
        line["gpu_eager_baseline"]["speedup_whole_path"] = value / line["gpu_eager_baseline"]["whole_path_images_per_s"]
    if world == 1 and not args.no_cpu:
        n = args.cpu_sample
        orc = CpuOracle(2, n)
        orc.tp = {k: v.detach().float().cpu() for k, v in tok.image_feature_encoder.vision_tower.state_dict().items()}   # the bench model's own weights
        orc.run(1)
        secs, st = orc.run(n)
        cores = orc.cores
        line["cpu_baseline"] = {"value": n / secs, "unit": UNIT, "cores": cores, "kind": "port",
                                "sample": f"the first {n} of the 256 images (the same host tensors the GPU arm uploads), oracle port of the reference, torch CPU fp32, "
                                          f"{cores} threads: ViT {st['vit_s']:.1f}s + per-image head loop {st['head_s']:.1f}s"}
        # float parity of the bench run itself: this GPU's tower features against the fp32 oracle's on those images
        got = tok.image_feature_encoder(images[:n]).float().cpu()
        ref = st["feats"]
        line["parity"] = {"tower_rel_frobenius_vs_fp32_oracle": float((got - ref).norm() / ref.norm()),
                          "tower_max_abs_over_max_vs_fp32_oracle": float((got - ref).abs().max() / ref.abs().max()),
                          "north_star_tolerance": 1e-3, "images": n,
                          "note": "23 layers of bf16-operand tensor-core GEMMs with fp32 accumulation, f32 residual stream, split patch embedding; "
                                  "torch's own bf16 evaluation of the same tower is 1.2e-2 (profiles/r02_parity_tower.json)"}
    print(json.dumps(line))
    H.finish()


def run_config3(args):
    """336^2 tokenizer + reconstruction decoder, bf16."""
    import setok_b200
    H = Harness(args)
    dev, world, rank = H.dev, H.world, H.rank
    B = BATCHES[3]
    tok = build_model(dev, 336)
    torch.manual_seed(1)
    det = setok_b200.SetokDeTokenizer(**DETOK).to(dev)
    layers_run = tok.image_feature_encoder.layers_to_run()
    _, imgs_h, noise_h = host_batch(3, rank)
    imgs_h = imgs_h.to(torch.bfloat16)
    images, noise = imgs_h.to(dev), noise_h.to(dev)
    last = {}

    def run_steps(n):
        for _ in range(n):
            rt, idx, score = tok(images, k=KNN_K, noise=noise)
            last["rt"], last["recon"] = rt, det(rt)
    ms, launches, clocks = H.timed(run_steps)
    value = world * B * args.steps / (ms * 1e-3)
    h_img, h_noise = pin(imgs_h), pin(noise_h)
    copy_stream, d2h_stream = torch.cuda.Stream(dev), torch.cuda.Stream(dev)
    h_out = torch.empty(B, 576, DETOK["decoder_embed_dim"], dtype=torch.bfloat16).pin_memory()

    def e2e_steps(n):
        main = torch.cuda.current_stream(dev)
        nxt = None
        for i in range(n + 1):
            cur = nxt
            if i < n:
                with torch.cuda.stream(copy_stream):
                    d_img, d_noise = h_img.to(dev, non_blocking=True), h_noise.to(dev, non_blocking=True)
                    ev = torch.cuda.Event(); ev.record(copy_stream)
                nxt = (d_img, d_noise, ev)
            if cur is not None:
                main.wait_event(cur[2])
                cur[0].record_stream(main); cur[1].record_stream(main)
                rt, _, _ = tok(cur[0], k=KNN_K, noise=cur[1])
                recon = det(rt)
                done = torch.cuda.Event(); done.record(main)
                d2h_stream.wait_event(done)
                with torch.cuda.stream(d2h_stream):
                    recon.record_stream(d2h_stream)
                    h_out.copy_(recon, non_blocking=True)
        d2h_stream.synchronize()
        return h_out.numel() * 2
    ms_e, d2h = H.timed_e2e(e2e_steps)
    e2e_value = world * B * args.steps / (ms_e * 1e-3)
    if rank != 0:
        H.finish()
        return
    pk = peaks()
    counts = (last["rt"].offsets[1:] - last["rt"].offsets[:-1]).float()
    tok_ms = cuda_time(lambda: tok(images, k=KNN_K, noise=noise), 3)
    vit_ms = cuda_time(lambda: tok.image_feature_encoder(images), 3)
    rt = last["rt"]
    det_ms = cuda_time(lambda: det(rt), 3)
    gemm_ms, gemm_flops, gemm_tf = time_gemm_mix(dev, layers_run, B * 577, reps=1)
    Q, Hd, Dd = 576, 768, 768
    det_flops = B * (6 * (8 * Q * Hd * Hd + 4 * Q * Q * Hd + 4 * Q * Hd * 3072) + 16 * (24 * Q * Dd * Dd + 4 * Q * Q * Dd))
    flops = B * vit_flops_per_image(layers_run, 336) + det_flops
    step_ms = ms / args.steps
    line = {"metric": "images/sec through SeTok tokenizer + reconstruction decoder (336^2, ViT-L/14)", "value": value, "unit": UNIT, "n_gpus": world,
            "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": step_ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "bf16", "data": "synthetic",
            "config": dict(workload_config(3, world), k_per_image={"min": float(counts.min()), "mean": float(counts.mean()), "max": float(counts.max())},
                           decoder="decoder_embed_dim 768 (reference default 4096 cannot run: exceeds the position-embedding channels, DESIGN.md D2)"),
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h_img.numel() * 2 + h_noise.numel() * 4, "d2h_bytes_per_step": d2h},
            "gpu_launches": int(launches), "clocks": clocks,
            "stages": {"tokenizer_ms": tok_ms, "detokenizer_ms": det_ms},
            "roofline": {"kernel": "gemm_bf16_tcgen05_kernel (ViT layer launch mix at M = 128 x 577 rows)", "bound": "tensor", "achieved": gemm_tf,
                         "peak": pk["tf_sustained"], "unit": "TFLOP/s", "frac": gemm_tf / pk["tf_sustained"], "traffic": None,
                         "flops_per_launch": gemm_flops, "ms_per_launch": gemm_ms, "peak_source": f"{pk['src']} sustained bf16",
                         "whole_step_tensor_frac": (flops / (step_ms * 1e-3) / 1e12) / pk["tf_sustained"], "whole_step_flops": flops},
            "roofline_vit": {"kernel": "whole ViT-L/14 tower at 336^2 (T = 577; chunked tcgen05 attention)", "bound": "tensor", "peak": pk["tf_sustained"], "unit": "TFLOP/s",
                             "achieved": B * vit_flops_per_image(layers_run, 336) / (vit_ms * 1e-3) / 1e12,
                             "frac": (B * vit_flops_per_image(layers_run, 336) / (vit_ms * 1e-3) / 1e12) / pk["tf_sustained"], "ms": vit_ms,
                             "flops": B * vit_flops_per_image(layers_run, 336)}}
    if world == 1 and not args.no_cpu:
        orc = CpuOracle(3, 1)
        secs, st = orc.run(1)
        line["cpu_baseline"] = {"value": 1 / secs, "unit": UNIT, "cores": orc.cores, "kind": "port",
                                "sample": f"1 of the 128 images, oracle port (tokenizer + detokenizer), torch CPU fp32, {orc.cores} threads"}
    print(json.dumps(line))
    H.finish()


def run_config4(args):
    """encode_images -> mlp2x_gelu projector, bf16, DP with the ragged all-gather of the projected rows."""
    import setok_b200
    from setok_b200.dist import RaggedAllGather
    H = Harness(args)
    dev, world, rank = H.dev, H.world, H.rank
    B = args.batch or BATCHES[4]
    tok = build_model(dev)
    torch.manual_seed(2)
    proj = setok_b200.build_vision_projector("mlp2x_gelu", mm_hidden_size=1024, hidden_size=4096).to(dev)
    u8, imgs_h, noise_h = host_batch(4, rank, B)
    imgs_h = imgs_h.to(torch.bfloat16)
    images, noise = imgs_h.to(dev), noise_h.to(dev)
    gather = RaggedAllGather(B, device=dev) if world > 1 else None
    last = {}

    def run_steps(n):
        pending = None
        for _ in range(n):
            out = setok_b200.encode_images(tok, proj, images, k=KNN_K, noise=noise)
            if gather is not None:
                h = gather.start(out)
                if pending is not None:
                    last["out"] = gather.finish(pending)
                pending = h
            else:
                last["out"] = out
        if pending is not None:
            last["out"] = gather.finish(pending)
        last["local"] = out
    ms, launches, clocks = H.timed(run_steps)
    value = world * B * args.steps / (ms * 1e-3)
    from setok_b200.pipeline import stream_tokenize
    h_img, h_noise = pin(imgs_h), pin(noise_h)

    def e2e_steps(n):
        nbytes = 0
        post = lambda r, i, s_: (proj(r), i, s_)
        for res in stream_tokenize(tok, ((h_img, h_noise) for _ in range(n)), post=post, gather=gather, k=KNN_K):
            nbytes = res.nbytes
        return nbytes
    ms_e, d2h = H.timed_e2e(e2e_steps)
    e2e_value = world * B * args.steps / (ms_e * 1e-3)
    if rank != 0:
        H.finish()
        return
    pk = peaks()
    counts = (last["local"].offsets[1:] - last["local"].offsets[:-1]).float()
    rt, _, _ = tok(images, k=KNN_K, noise=noise)
    proj_ms = cuda_time(lambda: proj(rt), 10)
    rows = float(counts.sum())
    proj_flops = 2 * rows * (1024 * 4096 + 4096 * 4096)
    layers_run = tok.image_feature_encoder.layers_to_run()
    gemm_ms, gemm_flops, gemm_tf = time_gemm_mix(dev, layers_run, B * 257, reps=2)
    step_ms = ms / args.steps
    line = {"metric": "images/sec through encode_images -> mm_projector (224^2, ViT-L/14, mlp2x_gelu to 4096)", "value": value, "unit": UNIT, "n_gpus": world,
            "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": step_ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "bf16", "data": "synthetic",
            "config": dict(workload_config(4, world), batch_per_gpu=B, global_batch=B * world, ragged_all_gather=gather is not None,
                           k_per_image={"min": float(counts.min()), "mean": float(counts.mean()), "max": float(counts.max())}),
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h_img.numel() * 2 + h_noise.numel() * 4, "d2h_bytes_per_step": d2h},
            "gpu_launches": int(launches), "clocks": clocks,
            "stages": {"projector_ms": proj_ms, "projector_rows": rows, "projector_tflops": proj_flops / (proj_ms * 1e-3) / 1e12},
            "roofline": {"kernel": f"gemm_bf16_tcgen05_kernel (ViT layer launch mix at M = {B} x 257 rows)", "bound": "tensor", "achieved": gemm_tf,
                         "peak": pk["tf_sustained"], "unit": "TFLOP/s", "frac": gemm_tf / pk["tf_sustained"], "traffic": None,
                         "flops_per_launch": gemm_flops, "ms_per_launch": gemm_ms, "peak_source": f"{pk['src']} sustained bf16",
                         "vit_tensor_frac_of_step": (B * vit_flops_per_image(layers_run) / (step_ms * 1e-3) / 1e12) / pk["tf_sustained"]}}
    if world == 1 and not args.no_cpu:
        orc = CpuOracle(4, 4)
        orc.run(1)
        secs, st = orc.run(4)
        line["cpu_baseline"] = {"value": 4 / secs, "unit": UNIT, "cores": orc.cores, "kind": "port",
                                "sample": f"4 of the {B} images, oracle port (tokenizer + projector), torch CPU fp32, {orc.cores} threads"}
    print(json.dumps(line))
    H.finish()


def run_config5(args):
    """Mixed-resolution ragged batch: the global list of images is dealt to the ranks by N^2 cost; outputs come back in order."""
    from setok_b200.dist import RaggedAllGather, deal_by_cost
    H = Harness(args)
    dev, world, rank = H.dev, H.world, H.rank
    B = BATCHES[5]
    tok = build_model(dev)
    # the global batch is the concatenation of every rank's host_batch(5, r); each rank takes the images dealt to it
    sizes_global, imgs_global, noise_global = [], [], []
    for r in range(world):
        _, im, nz = host_batch(5, r)
        imgs_global += im
        noise_global += nz
        sizes_global += [int(t.shape[-1]) for t in im]
    mine = deal_by_cost(sizes_global, world)[rank]
    imgs_h = [imgs_global[i] for i in mine]
    images = [t.to(dev) for t in imgs_h]
    noise = [noise_global[i].to(dev) for i in mine]
    gather = RaggedAllGather(max(len(p_) for p_ in deal_by_cost(sizes_global, world)), device=dev) if world > 1 else None
    last = {}

    def run_steps(n):
        pending = None
        for _ in range(n):
            rt, idxs, scores = tok(images, k=KNN_K, noise=noise, interpolate_pos_encoding=True)
            if gather is not None:
                h = gather.start(rt, order=mine)
                if pending is not None:
                    last["rt"] = gather.finish(pending)
                pending = h
            else:
                last["rt"] = rt
        if pending is not None:
            last["rt"] = gather.finish(pending)
        last["local"] = rt
    ms, launches, clocks = H.timed(run_steps)
    n_img = len(sizes_global)
    value = n_img * args.steps / (ms * 1e-3)
    h_imgs = pin(imgs_h)

    def e2e_steps(n):
        nbytes = 0
        for _ in range(n):
            d = [t.to(dev, non_blocking=True) for t in h_imgs]
            rt, idxs, scores = tok(d, k=KNN_K, noise=noise, interpolate_pos_encoding=True)
            if gather is not None:
                rt = gather.finish(gather.start(rt, order=mine))
            host = rt.packed().cpu()
            nbytes = host.numel() * host.element_size()
        return nbytes
    ms_e, d2h = H.timed_e2e(e2e_steps)
    e2e_value = n_img * args.steps / (ms_e * 1e-3)
    if rank != 0:
        H.finish()
        return
    pk = peaks()
    counts = (last["local"].offsets[1:] - last["local"].offsets[:-1]).float()
    layers_run = tok.image_feature_encoder.layers_to_run()
    flops = sum(vit_flops_per_image(layers_run, s) for s in sizes_global)
    step_ms = ms / args.steps
    gemm_ms, gemm_flops, gemm_tf = time_gemm_mix(dev, layers_run, 32 * 1025, reps=1)
    cost = [sum(sizes_global[i] ** 2 for i in p_) for p_ in deal_by_cost(sizes_global, world)]
    line = {"metric": "images/sec through SeTok tokenizer (mixed 224^2/336^2/448^2, ViT-L/14)", "value": value, "unit": UNIT, "n_gpus": world,
            "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": step_ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "bf16", "data": "synthetic",
            "config": dict(workload_config(5, world), ragged_all_gather=gather is not None, deal_cost_max_over_mean=max(cost) / (sum(cost) / len(cost)),
                           k_per_image={"min": float(counts.min()), "mean": float(counts.mean()), "max": float(counts.max())}),
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": sum(t.numel() * 4 for t in h_imgs), "d2h_bytes_per_step": d2h},
            "gpu_launches": int(launches), "clocks": clocks,
            "roofline": {"kernel": "gemm_bf16_tcgen05_kernel (ViT layer launch mix at M = 32 x 1025 rows, the 448^2 third of the batch)", "bound": "tensor",
                         "achieved": gemm_tf, "peak": pk["tf_sustained"], "unit": "TFLOP/s", "frac": gemm_tf / pk["tf_sustained"], "traffic": None,
                         "flops_per_launch": gemm_flops, "ms_per_launch": gemm_ms, "peak_source": f"{pk['src']} sustained bf16",
                         "vit_tensor_frac_of_step": (flops / world / (step_ms * 1e-3) / 1e12) / pk["tf_sustained"]}}
    if world == 1 and not args.no_cpu:
        orc = CpuOracle(5, 3)
        secs, st = orc.run(3)
        line["cpu_baseline"] = {"value": st["images"] / secs, "unit": UNIT, "cores": orc.cores, "kind": "port",
                                "sample": f"one image of each resolution (224/336/448) of the 96-image batch, oracle port, torch CPU fp32, {orc.cores} threads"}
    print(json.dumps(line))
    H.finish()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--config", type=int, default=2, choices=[2, 3, 4, 5], help="BASELINE.json configuration (default 2: the metric's own)")
    ap.add_argument("--batch", type=int, default=0, help="config 4 only: images per GPU (default 64; 8 = the latency variant)")
    ap.add_argument("--no-gather", action="store_true", help="skip the ragged all-gather at N>1")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--no-extras", action="store_true", help="skip the gpu_eager_baseline and K~32 legs")
    ap.add_argument("--cpu-sample", type=int, default=8)
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        {2: run_config2, 3: run_config3, 4: run_config4, 5: run_config5}[args.config](args)


if __name__ == "__main__":
    main()
