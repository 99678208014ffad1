#!/usr/bin/env python
"""bench.py -- denoised motion-frames/s of the SynTalker sampling hot path on B200 (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--engine simt|tc] [--batch 32]

One "step" = one pass of the hot path over one batch of synthetic windows: conditioning encode ->
DDIM-50 sampling with classifier-free guidance 2.0 (2 denoiser evaluations per diffusion step) ->
x5 -> RVQ latent2origin x3 -> 330-d pose features  (BASELINE config 2: diffusion_rvqvae_128, batch 32).
`value` times that with inputs resident in HBM (st_generate_330); `e2e` times the same through the
host-buffer C-ABI call (st_generate_330_host: pinned host inputs, H2D, compute, D2H of the poses).
N > 1 (torchrun): every rank runs its own 32 clips (weak scaling, no data-path collective) and the ranks
all-gather the poses at the end of each step over NCCL.
`--impl reference` times the reference's own algorithm as written (audio encoder re-run in every
evaluation, 2 evaluations per step) through the CPU oracle port on the host cores, on a bounded sample.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "denoised motion-frames/sec (128-frame seq, 1000->50 DDIM)"
UNIT = "frames/s"
S_STEPS = 50
CFG_SCALE = 2.0
# algorithmic FLOPs per clip: cond encoder once, trunk per evaluation (+512-d style branch), decode once
E_COND, E_TRUNK, E_DEC = 4.677e9, 1.2342e9 + 0.0336e9, 3.899e9


def ncu_traffic():
    """(dram__bytes_read.sum + dram__bytes_write.sum per launch of the dominant kernel, file it was read from): the mean over the
    launches of the newest committed `ncu --set full` capture of the trunk kernel under profiles/ (cold-cache replays of the kernel inside
    the sampling loop; tests/gpu_evidence.sh writes it).  (None, None) if absent."""
    import csv
    for name in ("r2c_ncu_step_fast.csv", "r2b_ncu_step_fast.csv", "r2_ncu_step_fast.csv", "r1c_ncu_gemm_tc_fast_full.csv", "r1b_ncu_gemm_tc_fast_full.csv"):
        p = os.path.join(ROOT, "profiles", name)
        if os.path.exists(p):
            break
    else:
        return None, None
    rows = list(csv.reader(open(p)))
    hdr, units, data = rows[0], rows[1], rows[2:]
    scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
    tot = 0.0
    for r in data:
        for col in ("dram__bytes_read.sum", "dram__bytes_write.sum"):
            i = hdr.index(col)
            tot += float(r[i].replace(",", "")) * scale.get(units[i], 1.0)
    return tot / max(len(data), 1), "profiles/" + name


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return {"tflops": d.get("bf16_tflops_sustained", 1385.6), "hbm": d.get("hbm_gbs", 6534.1), "src": "measured (MEASURED_PEAKS.json, sustained bf16)"}
    return {"tflops": 1400.0, "hbm": 6650.0, "src": "fallback (B200_PROFILING.md)"}


class ClockSampler(threading.Thread):
    """SM clock, power and throttle reasons DURING the timed region: one `nvidia-smi -lms 20` process streams samples
    (spawning nvidia-smi per sample costs 100+ ms, longer than a timed step)."""
    Q = "index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, gpu):
        super().__init__(daemon=True)
        self.gpu, self.rows, self.proc = gpu, [], None

    def run(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.gpu}", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "20"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            import atexit
            atexit.register(lambda p=self.proc: p.poll() is None and p.kill())      # never leave the sampler behind
            for line in self.proc.stdout:
                c = [x.strip() for x in line.split(",")]
                if len(c) >= 8:
                    self.rows.append(c)
        except Exception:
            pass

    def summary(self):
        if self.proc is not None:
            try:
                self.proc.terminate()
            except Exception:
                pass
        rows = list(self.rows)
        isnum = lambda v: v.replace(".", "", 1).isdigit()
        sm = sorted(float(r[1]) for r in rows if isnum(r[1]))
        mx = max([float(r[2]) for r in rows if isnum(r[2])] or [0])
        pw = [float(r[3]) for r in rows if isnum(r[3])]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for i, n in enumerate(names) if any(r[4 + i].lower().startswith("active") for r in rows)]
        # under load = upper half of the power-ranked samples (the sampler also sees the idle gaps between steps)
        by_power = sorted((float(r[3]), float(r[1])) for r in rows if isnum(r[1]) and isnum(r[3]))
        load = sorted(c for _, c in by_power[len(by_power) // 2:]) if by_power else sm
        return {"sm_mhz": load[len(load) // 2] if load else None, "sm_max_mhz": mx or None, "reasons": reasons, "samples": len(sm),
                "power_w_max": max(pw) if pw else None}


def cpu_reference_sample(B_s, S_s, threads, reps=1):
    """The reference's algorithm as written, on the CPU oracle port: S_s of the 50 DDIM steps with the
    ClassifierFreeSampleModel wrapper (2 full MDM.forward per step, WavEncoder inside each), then x5 ->
    latent2origin x3 -> 330-d. Per-step cost is constant in t, so the 50-step time is extrapolated."""
    import torch
    from oracle import diffusion as odiff, mdm as omdm, pose as opose, rvq as orvq
    from syntalker_b200 import synth
    from syntalker_b200.pipeline import load_mean_std
    torch.set_num_threads(threads)
    torch.set_grad_enabled(False)
    W = synth.mdm_state_dict("beatx_motionclip", seed=0)
    vqw = [synth.rvq_state_dict(d, seed=0) for d in synth.PART_DIMS_BEATX]
    inp = synth.make_inputs(B_s, seed=1, variant="beatx_motionclip")
    y = {k: inp[k] for k in ("audio", "word", "seed", "style_feature")}
    y["scale"] = torch.ones(1) * CFG_SCALE
    sched = odiff.make_schedule(use_ddim=True)
    fn = lambda x, t, yy: omdm.cfg_text(lambda a, b, c: omdm.mdm_forward(W, a, b, c, "beatx_motionclip"), x, t, yy)
    ms = load_mean_std()
    best = None
    for _ in range(reps):
        x = inp["noise"]
        t0 = time.perf_counter()
        for k in range(S_STEPS - 1, S_STEPS - 1 - S_s, -1):
            t = torch.full((B_s,), sched.timestep_map[k], dtype=torch.int64)
            x0 = fn(x, t, y)
            eps = (float(sched.sqrt_recip_alphas_cumprod[k]) * x - x0) / float(sched.sqrt_recipm1_alphas_cumprod[k])
            x = x0 * float(sched.alphas_cumprod_prev[k]) ** 0.5 + (1 - float(sched.alphas_cumprod_prev[k])) ** 0.5 * eps
        t_loop = time.perf_counter() - t0
        t0 = time.perf_counter()
        lats = opose.sample_to_parts(x, 5.0)
        recs = [orvq.latent2origin(w, l)[0] for w, l in zip(vqw, lats)]
        opose.assemble_330(recs[0], recs[1], recs[2], ms, None)
        t_dec = time.perf_counter() - t0
        full = t_loop * S_STEPS / S_s + t_dec
        best = full if best is None else min(best, full)
    return B_s * 128 / best, best


def run_reference(args, rank, world):
    if rank != 0:
        return
    import torch
    threads = os.cpu_count() or 1
    B_s, S_s = 16, 2
    for _ in range(args.warmup):
        cpu_reference_sample(2, 1, threads)
    vals = []
    t0 = time.perf_counter()
    for _ in range(args.steps):
        v, _full = cpu_reference_sample(B_s, S_s, threads)
        vals.append(v)
    wall = time.perf_counter() - t0
    v = sum(vals) / len(vals)
    sample = (f"B={B_s} clips x {S_s} of {S_STEPS} DDIM steps with CFG (2 full MDM.forward per step incl. WavEncoder, as the reference "
              f"is written) + decode + 330-d; 50-step time extrapolated (per-step cost constant in t)")
    line = {"impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": 1000 * wall / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
            "data": "synthetic", "config": {"workload": "diffusion_rvqvae_128 + use_motionclip, batch=32, 50 DDIM steps, CFG scale 2.0 (BASELINE config 2)",
                                            "global_batch": 32 * args.gpus, "frames": 128},
            "cpu_baseline": {"value": v, "unit": UNIT, "cores": threads, "kind": "port", "sample": sample, "extrapolated": True},
            "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


def other_configs(dev, rank, world, pk):
    """BASELINE configs 3, 4, 5 (the headline above is config 2), device-timed with CUDA events, inputs resident in HBM, after the
    warm-up each needs to reach CUDA-graph replay.  Not the headline: one object on the same JSON line so that the driver's run
    observes them.  Under torchrun config 3 runs in its real shape (32 clips per rank, poses all-gathered); 4 and 5 are N = 1 only."""
    import torch
    import torch.distributed as dist
    from syntalker_b200 import _lib, synth
    from syntalker_b200.cfg_sampler import TwoClassifierFreeSampleModel_Bodypart
    from syntalker_b200.denoiser import MDM
    from syntalker_b200.denoiser_h3d import MDM as MDM_H3D
    from syntalker_b200.diffusion import create_gaussian_diffusion
    from syntalker_b200.pipeline import Window623, load_mean_std, pose_assemble_330
    from syntalker_b200.vq import RVQVAE
    out = {}

    def timed(fn, n, warm):
        for _ in range(warm):
            fn()
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(n):
            fn()
        e1.record()
        torch.cuda.synchronize()
        t = torch.tensor([e0.elapsed_time(e1) / n], device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    # ---- config 3: diffusion_rvqvae_128, 1000-step p_sample_loop (noise drawn on the device, one randn per step like the
    # reference, gaussian_diffusion.py:541), 32 clips per GPU, decode + 330-d, poses gathered over NCCL when sharded ----
    B = 32
    model = MDM(None).load_state_dict(synth.mdm_state_dict("beatx", seed=0))
    vqs = [RVQVAE(None, d).load_state_dict(synth.rvq_state_dict(d, seed=0)) for d in synth.PART_DIMS_BEATX]
    inp = synth.make_inputs(B, seed=1 + 1000 * rank, variant="beatx")
    y = {k: inp[k].to(dev).contiguous() for k in ("audio", "word", "seed")}
    x0 = inp["noise"].to(dev).contiguous()
    diff = create_gaussian_diffusion()
    ms_ = {k: v.to(dev) for k, v in load_mean_std().items()}
    gathered = torch.empty((world * B, 128, 330), device=dev) if world > 1 else None
    n0 = _lib.launch_count()

    def cfg3():
        sample = diff.p_sample_loop(model, (B, 1536, 1, 32), noise=x0, clip_denoised=False, model_kwargs={"y": y})
        lat = sample.squeeze(2).permute(0, 2, 1) * 5.0                   # trainer:457-476 (batched like h3d trainer:735)
        recs = [v.latent2origin(lat[..., 512 * k:512 * (k + 1)].contiguous())[0] for k, v in enumerate(vqs)]
        pose, _ = pose_assemble_330(recs[0], recs[1], recs[2], ms_, None)
        if world > 1:
            dist.all_gather_into_tensor(gathered, pose)
        return pose

    ms3 = timed(cfg3, 2, 2)
    alg3 = B * (E_COND + 1000 * 1.2342e9 + E_DEC)
    out["config3"] = {"what": "1000-step DDPM p_sample_loop (z recursion, W_x eps per 50-step chunk, device noise draws) + decode + 330-d, "
                              f"32 clips per GPU{' x ' + str(world) + ' GPUs, poses all-gathered (NCCL)' if world > 1 else ' (one shard of B=256)'}",
                      "ms": ms3, "frames_per_s": world * B * 128 / (ms3 / 1e3), "us_per_diffusion_step": ms3, "n_gpus": world,
                      "algorithmic_tflops_per_gpu": alg3 / (ms3 / 1e3) / 1e12, "frac_of_peak": alg3 / (ms3 / 1e3) / 1e12 / pk["tflops"]}
    del model
    if world > 1:
        return out
    # ---- config 4: diffusion_h3d, upper + lower prompts + audio, B = 64, DDIM-50, body-part CFG (9 evaluations -> 4) ----
    B = 64
    m_h = MDM_H3D(None).load_state_dict(synth.mdm_state_dict("h3d", seed=0))
    vqs_h = [RVQVAE(None, d).load_state_dict(synth.rvq_state_dict(d, seed=0)) for d in synth.PART_DIMS_H3D]
    inp = synth.make_inputs(B, seed=41, variant="h3d")
    sf = {"upper_mask": inp["style_upper"].to(dev), "hands_mask": None, "lower_mask": inp["style_lower"].to(dev)}
    w623 = Window623(TwoClassifierFreeSampleModel_Bodypart(m_h), create_gaussian_diffusion(use_ddim=True), *vqs_h)
    d_in = {k: inp[k].to(dev).contiguous() for k in ("audio", "word", "seed", "noise")}
    ms4 = timed(lambda: w623.run(d_in["audio"], d_in["word"], d_in["seed"], d_in["noise"], sf), 3, 3)
    alg4 = B * (E_COND + 50 * 4 * (1.2342e9 + 0.0252e9) + 4.1e9)
    out["config4"] = {"what": "denoiser_h3d in TwoClassifierFreeSampleModel_Bodypart (9 evaluations per step as written, 4 de-duplicated), "
                              "DDIM-50, B=64, decode 156/360/107 + 623-d", "ms": ms4, "frames_per_s": B * 128 / (ms4 / 1e3),
                      "algorithmic_tflops": alg4 / (ms4 / 1e3) / 1e12, "frac_of_peak": alg4 / (ms4 / 1e3) / 1e12 / pk["tflops"]}
    del m_h, w623
    # ---- config 5: RVQ decode-only, 3 body parts, B = 1024 x 128 frames ----
    B = 1024
    g = torch.Generator().manual_seed(5)
    lats = [(5.0 * torch.randn(B, 32, 512, generator=g)).to(dev) for _ in range(3)]
    work = [l.clone() for l in lats]

    def dec():
        for w, l in zip(work, lats):
            w.copy_(l)                                                   # latent2origin leaves the residual in its input
        return [v.latent2origin(w)[0] for v, w in zip(vqs, work)]

    ms5 = timed(dec, 5, 3)
    out["config5"] = {"what": "RVQ decode-only (latent2origin x 3 body parts), B=1024 x 128 frames", "ms": ms5,
                      "frames_per_s": B * 128 / (ms5 / 1e3), "algorithmic_tflops": B * E_DEC / (ms5 / 1e3) / 1e12,
                      "frac_of_peak": B * E_DEC / (ms5 / 1e3) / 1e12 / pk["tflops"]}
    out["gpu_launches"] = int(_lib.launch_count() - n0)
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--engine", default=os.environ.get("ST_ENGINE", "tc"), choices=["simt", "tc"])
    ap.add_argument("--batch", type=int, default=32)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    rank, world = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
    local = int(os.environ.get("LOCAL_RANK", 0))
    if args.impl == "reference":
        return run_reference(args, rank, world)

    import torch
    import torch.distributed as dist
    from syntalker_b200 import _lib, synth
    from syntalker_b200.cfg_sampler import ClassifierFreeSampleModel
    from syntalker_b200.denoiser import MDM
    from syntalker_b200.diffusion import create_gaussian_diffusion
    from syntalker_b200.pipeline import Window330
    from syntalker_b200.vq import RVQVAE

    assert torch.cuda.is_available(), "bench.py needs a GPU (use --impl reference for the CPU arm)"
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    torch.set_grad_enabled(False)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
    _lib.set_engine(args.engine)
    _lib.check(_lib.lib().st_set_pdl(int(os.environ.get("ST_PDL", "1"))))
    _lib.check(_lib.lib().st_set_graphs(int(os.environ.get("ST_GRAPHS", "1"))))
    if os.environ.get("ST_PROBE"):          # A/B experiments only (st_debug_probe bits); the default run sets nothing
        _lib.check(_lib.lib().st_debug_probe(int(os.environ["ST_PROBE"])))
    B = args.batch
    model = MDM(None).load_state_dict(synth.mdm_state_dict("beatx_motionclip", seed=0))
    wrapped = ClassifierFreeSampleModel(model)
    vqs = [RVQVAE(None, d).load_state_dict(synth.rvq_state_dict(d, seed=0)) for d in synth.PART_DIMS_BEATX]
    diff = create_gaussian_diffusion(use_ddim=True)
    inp = synth.make_inputs(B, seed=1 + 1000 * rank, variant="beatx_motionclip")
    y_host = {"scale": torch.ones(1) * CFG_SCALE, "style_feature": inp["style_feature"]}
    win = Window330(wrapped, diff, *vqs, B=B, use_ddim=True)
    d_in = {k: inp[k].to(dev).contiguous() for k in ("audio", "word", "seed", "noise", "style_feature")}
    y_dev = {"scale": torch.ones(1) * CFG_SCALE, "style_feature": d_in["style_feature"]}
    outs = (torch.empty((B, 128, 330), device=dev), torch.empty((B, 128, 3), device=dev), torch.empty((B, 1536, 1, 32), device=dev))
    gathered = torch.empty((world * B, 128, 330), device=dev) if world > 1 else None
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)          # > 126 MB L2

    def step_device():
        pose, _, _ = win.run_device(d_in["audio"], d_in["word"], d_in["seed"], d_in["noise"], y=y_dev, out=outs)
        if world > 1:
            dist.all_gather_into_tensor(gathered, pose)

    pinned = {k: inp[k].contiguous().pin_memory() for k in ("audio", "word", "seed", "noise")}   # the step's inputs live in pinned host memory

    def step_host():
        win.run(pinned["audio"], pinned["word"], pinned["seed"], pinned["noise"], y=y_host)

    def timed(fn, n):
        evs = []
        for _ in range(n):
            flush.zero_()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); fn(); e1.record()
            evs.append((e0, e1))
        torch.cuda.synchronize()
        return [a.elapsed_time(b) for a, b in evs]

    for _ in range(args.warmup):
        step_device()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    sampler = ClockSampler(local)
    sampler.start()
    n0 = _lib.launch_count()
    torch.cuda.synchronize()
    w0 = time.perf_counter()
    ms = timed(step_device, args.steps)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    wall = time.perf_counter() - w0
    launches = _lib.launch_count() - n0
    clocks = sampler.summary()
    tot = torch.tensor([sum(ms)], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(tot, op=dist.ReduceOp.MAX)
    tot_ms = float(tot.item())
    value = world * B * 128 * args.steps / (tot_ms / 1000)

    # ---- end to end through the host-buffer call ----
    # every step: its inputs go H2D from pinned host memory and its poses come back D2H, inside the timed region.  The
    # blocking call (st_generate_330_host = begin + wait per step) is timed first; the headline keeps TWO window batches in
    # flight (Window330.begin / wait on alternating staging sets): the copies of step i + 1 / i - 1 overlap the computation
    # of step i.  One event pair around all K steps (L2 flush between steps inside the region).
    for _ in range(2):
        step_host()
    ms_h = timed(step_host, args.steps)
    serial_ms = sum(ms_h) / args.steps

    def pipelined(n):
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(n):
            flush.zero_()
            win.begin(i % 2, pinned["audio"], pinned["word"], pinned["seed"], pinned["noise"], y=y_host)
            if i >= 1:
                win.wait((i - 1) % 2)
        win.wait((n - 1) % 2)
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1)

    pipelined(2)
    if world > 1:
        dist.barrier()
    toth = torch.tensor([pipelined(args.steps)], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(toth, op=dist.ReduceOp.MAX)
    e2e = world * B * 128 * args.steps / (float(toth.item()) / 1000)

    # ---- roofline of the dominant kernel class (the GEMM engine), CUDA events around every launch ----
    pk = peaks()
    _lib.profile_begin()
    step_device()
    g_ms, g_flops, g_n = _lib.profile_end()
    torch.cuda.synchronize()
    achieved = g_flops / (g_ms / 1000) / 1e12 if g_ms > 0 else 0.0
    alg = B * (E_COND + S_STEPS * 2 * E_TRUNK + E_DEC)
    step_ms = tot_ms / args.steps
    roof = {"bound": "tensor", "achieved": achieved, "peak": pk["tflops"], "unit": "TFLOP/s", "frac": achieved / pk["tflops"], "traffic": ncu_traffic()[0], "traffic_source": ncu_traffic()[1],
            "kernel": "gemm_tc_fast_kernel / gemm_tc_kernel (tcgen05 split-fp16; the qkv launches carry the fused attention)" if args.engine == "tc" else "gemm_simt_kernel (exact fp32 FMA)",
            "launches_profiled": g_n, "kernel_ms_per_step": g_ms, "kernel_share_of_step": g_ms / step_ms if step_ms else None,
            "executed_flops_per_step": g_flops, "algorithmic_flops_per_step": alg,
            "whole_step_tflops": alg / (step_ms / 1000) / 1e12, "peak_source": pk["src"]}

    line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": step_ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32" if args.engine == "simt" else "f32 via split-fp16 (3 tcgen05 MMAs per product, fp32 accumulate)",
            "data": "synthetic",
            "config": {"workload": "diffusion_rvqvae_128 + use_motionclip, batch=32 per GPU, 50 DDIM steps, CFG scale 2.0 (BASELINE config 2)",
                       "global_batch": world * B, "frames": 128, "parallelism": f"dp{world} (clips sharded, weights replicated)",
                       "engine": args.engine, "l2": "256 MiB buffer written between timed steps (flush)", "wall_s": wall},
            "clocks": clocks, "gpu_launches": int(launches),
            "e2e": {"value": e2e, "unit": UNIT, "h2d_bytes_per_step": int(win.h2d_bytes), "d2h_bytes_per_step": int(win.d2h_bytes),
                    "ms_per_step": float(toth.item()) / args.steps, "ms_per_step_blocking_call": serial_ms,
                    "pipeline": "two window batches in flight (begin / wait on two staging sets): H2D of step i+1 and D2H of step i-1 overlap the computation of step i"},
            "roofline": roof}
    # ---- extra (not the headline): TWO window batches of the same configuration in flight on two handle sets / two streams.
    # Every layer of the loop is one wave of 128 CTAs on 148 SMs and a quarter of a layer's time is hand-over between
    # dependent kernels; a second, independent window batch fills those gaps (tests/concurrency_probe.py).  `value` and
    # `e2e` above stay one batch at a time -- this object reports what a caller with >= 64 clips queued gets.
    if world == 1 and not os.environ.get("ST_NO_TWO_IN_FLIGHT"):
        try:
            model_b = MDM(None).load_state_dict(synth.mdm_state_dict("beatx_motionclip", seed=0))
            vqs_b = [RVQVAE(None, d).load_state_dict(synth.rvq_state_dict(d, seed=0)) for d in synth.PART_DIMS_BEATX]
            win_b = Window330(ClassifierFreeSampleModel(model_b), diff, *vqs_b, B=B, use_ddim=True)
            outs_b = tuple(torch.empty_like(t) for t in outs)
            sets = [(win, outs), (win_b, outs_b)]
            streams = [torch.cuda.Stream(), torch.cuda.Stream()]
            main_s = torch.cuda.current_stream()

            def two(n, host):
                torch.cuda.synchronize()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                for st in streams:
                    st.wait_stream(main_s)
                for i in range(n):
                    w_i, o_i = sets[i % 2]
                    with torch.cuda.stream(streams[i % 2]):
                        flush.zero_()
                        if host:
                            slot = (i // 2) % 2                        # both staging sets of both handle sets: four batches queued
                            if i >= 4:
                                w_i.wait(slot)
                            w_i.begin(slot, pinned["audio"], pinned["word"], pinned["seed"], pinned["noise"], y=y_host)
                        else:
                            w_i.run_device(d_in["audio"], d_in["word"], d_in["seed"], d_in["noise"], y=y_dev, out=o_i)
                if host:
                    for i in range(max(0, n - 4), n):
                        sets[i % 2][0].wait((i // 2) % 2)
                for st in streams:
                    main_s.wait_stream(st)
                e1.record()
                torch.cuda.synchronize()
                return e0.elapsed_time(e1)

            n2 = max(2, args.steps + args.steps % 2)
            two(4, False)
            t_dev = two(n2, False)
            same = bool(torch.equal(outs[0], outs_b[0]))               # same inputs, same weights: the two handle sets must agree bit for bit
            two(4, True)
            t_host = two(n2, True)
            line["two_in_flight"] = {"value": B * 128 * n2 / (t_dev / 1000), "e2e": B * 128 * n2 / (t_host / 1000), "unit": UNIT, "steps": n2,
                                     "ms_per_step": t_dev / n2, "e2e_ms_per_step": t_host / n2, "results_bitwise_equal": same,
                                     "what": "two independent 32-clip window batches in flight (two handle sets, two streams); not the headline"}
        except Exception as e:                                         # the extra must never cost the headline line
            line["two_in_flight"] = {"error": f"{type(e).__name__}: {e}"[:300]}
    if not os.environ.get("ST_NO_OTHER_CONFIGS"):
        try:
            oc = other_configs(dev, rank, world, pk)
            line["other_configs"] = oc
        except Exception as e:                                             # the extra must never cost the headline line
            line["other_configs"] = {"error": f"{type(e).__name__}: {e}"[:300]}
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        threads = os.cpu_count() or 1
        cpu_reference_sample(2, 1, threads)
        v, full = cpu_reference_sample(8, 2, threads)
        line["cpu_baseline"] = {"value": v, "unit": UNIT, "cores": threads, "kind": "port", "extrapolated": True,
                                "sample": f"oracle port, B=8 clips x 2 of 50 DDIM steps with CFG as the reference is written (WavEncoder inside every "
                                          f"evaluation) + decode + 330-d, 50-step time extrapolated: {full:.1f} s per 8 clips"}
    if rank == 0:
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
