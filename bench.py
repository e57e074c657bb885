#!/usr/bin/env python
"""Benchmark of the attribution hot path (BASELINE.json metric: latent-opt steps/sec at 1024 px).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

A step = one Adam step of every trajectory in flight on a rank (default: the 20 Latin-hypercube
guesses of one 1024 px image, BASELINE.json configs[2]): fingerprint embed -> StyleGAN2 synthesis
forward -> MSE-to-target -> synthesis backward to the latent -> Adam.  ``value`` is
trajectory-steps/s summed over all ranks (each rank owns its own image: weak scaling, no
data-path collective; the only exchange is the final gather of keys and losses).

`--impl reference` times the CPU oracle port of the same step (synthesis forward + backward to the
latent with the MSE loss, B=1, 1024 px) on the host cores; rank 0 only.
"""
import argparse
import json
import math
import os
import statistics
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
PKG = os.path.join(ROOT, "attributing-image-generative-models-using-latent-fingerprints-sg2_b200")
for p in (ROOT, PKG, os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)

METRIC = "attribution latent-opt steps/sec (images x guesses) at 1024px"
# average DRAM bytes per conv launch (read + write) of the default workload, from the ncu capture summarised in
# profiles/r01_conv_dram_traffic.md; None until that capture exists
CONV_DRAM_BYTES_PER_LAUNCH = 9.327e+08
UNIT = "trajectory-steps/s"


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--size", type=int, default=1024)
    ap.add_argument("--guesses", type=int, default=20, help="trajectories in flight per rank (n of src/params.py:17)")
    ap.add_argument("--precision", default="tf32", choices=["fp32", "tf32"],
                    help="tf32 = tcgen05 tensor-core convs (the numerics the reference runs on GPU); fp32 = CUDA-core convs")
    ap.add_argument("--cpu-baseline-steps", type=int, default=2)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    return ap.parse_args()


def peaks():
    try:
        return json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))), "measured"
    except Exception:
        return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0}, "fallback"


class ClockSampler:
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.proc, self.path = index, None, f"/tmp/lfp_clocks_{os.getpid()}.csv"

    def start(self):
        try:
            self.f = open(self.path, "w")
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "100"], stdout=self.f,
                                         stderr=subprocess.DEVNULL)
        except Exception:
            self.proc = None

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        self.proc.wait()
        self.f.close()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for line in open(self.path):
            parts = [p.strip() for p in line.split(",")]
            if len(parts) < 6:
                continue
            try:
                sm.append(float(parts[0]))
                mx.append(float(parts[1]))
            except ValueError:
                continue
            for n, v in zip(names, parts[2:6]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        try:
            os.remove(self.path)
        except OSError:
            pass
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def synthetic_setup(size, torch):
    """Random-init generator of the named architecture with perturbed zero-init parameters, a
    synthetic orthonormal PCA basis, N(0,1) noise maps (SURVEY.md 8d 'Synthetic inputs')."""
    import fixtures as fx
    params = fx.make_params(size, seed=1346)
    noise = fx.make_noise(size, seed=2002)
    pc, sigma, mean = fx.make_pca_basis(2)
    return params, noise, pc, sigma, mean


def oracle_step_fn(size, torch):
    """One reference-algorithm step on CPU: synthesis forward + MSE + backward to the latent."""
    import fixtures as fx
    import oracle
    params, noise, pc, sigma, mean = synthetic_setup(size, torch)
    sp = fx.split_basis(pc, sigma, 64, 448, 1.0)
    alpha = sp["sigma_main"] * fx.seeded((448, 1), 5)
    key = (fx.seeded((64, 1), 6) > 0).float()
    with torch.no_grad():
        target, _, _ = oracle.generate_with_alpha(params, size, alpha, sp["u_cap"], sp["v_cap"], sp["sigma_key"],
                                                  mean, key, noise)
    a = (sp["sigma_main"] * fx.seeded((448, 1), 7)).requires_grad_(True)
    k = torch.zeros(64, 1, requires_grad=True)

    def step():
        w0 = oracle.latent_from_alpha(sp["u_cap"], a, mean)
        wx = oracle.embed_fingerprint(sp["v_cap"], sp["sigma_key"], torch.sigmoid(k), w0, 1.0)
        est = oracle.generator_forward(params, [wx.reshape(1, -1)], size, input_is_latent=True, noise=noise)
        loss = oracle.mse_loss(target, est) + 0.1 * oracle.alpha_bound(a, sp["max_alpha"], sp["min_alpha"])
        torch.autograd.grad(loss, [a, k])
        return float(loss.detach())

    return step


def time_cpu(step, n, warm=1):
    for _ in range(warm):
        step()
    t0 = time.perf_counter()
    for _ in range(n):
        step()
    return (time.perf_counter() - t0) / n


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import torch
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    step = oracle_step_fn(args.size, torch)
    t_first = time_cpu(step, 1, warm=0)
    budget = 240.0
    k = max(1, min(args.steps, int(budget / max(t_first, 1e-3)) - args.warmup))
    w = min(args.warmup, max(0, int(0.25 * budget / max(t_first, 1e-3))))
    per = time_cpu(step, k, warm=w)
    v = 1.0 / per
    sample = f"{k} steps x 1 trajectory, synthesis fwd+bwd+MSE at {args.size}px, oracle port on torch-CPU"
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus, "steps": k,
        "warmup": w, "ms_per_step": per * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": {"workload": f"attribution_{args.size}px_n{args.guesses}_mse", "size": args.size,
                   "trajectories_per_rank": args.guesses, "loss": "mse", "key_len": 64, "shift": 448,
                   "precision": "fp32 (torch-CPU)", "sample": "one trajectory per step (the CPU runs them one at a time)"},
        "cpu_baseline": {"value": v, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }))


def run_ours(args):
    import torch
    import torch.distributed as dist
    from lfp_native import capi
    from lfp_native.synthesis import SynthesisPlan
    from attribution import AttributionEngine
    import fixtures as fx

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py (impl ours) needs a CUDA device: there is no CPU fallback")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    size, B = args.size, args.guesses
    prec = capi.PREC_TF32 if args.precision == "tf32" else capi.PREC_FP32

    params, noise, pc, sigma, mean = synthetic_setup(size, torch)
    plan = SynthesisPlan(size, device=dev)
    plan.load(params)
    eng = AttributionEngine(plan, noise, pc, sigma, mean, key_len=64, shift=448, sigma=1.0, sd=1.0, lr=0.2,
                            precision=prec)
    # per-rank image (independent units): target from a seeded alpha and key, as generate_with_alpha does
    alpha_t = (eng.sigma_main * fx.seeded((1, eng.n_main), 100 + rank).to(dev))
    key_t = (fx.seeded((1, 64), 200 + rank) > 0).to(dev)
    _, wx_t = eng.embed_with_key(alpha_t, key_t)
    target = eng.render(wx_t).clone()
    rs = __import__("numpy").random.RandomState(300 + rank)
    lhs = torch.from_numpy(__import__("numpy").stack([(rs.permutation(B) + 0.5) / B for _ in range(eng.n_main)], 1)).float()
    st = eng.init_state(eng.alpha0_from_lhs(lhs))

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(max(args.warmup, 3)):
        eng.step(st, target)
    barrier()
    # ---- timed region: device-resident (value) ----
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    L = capi.lib()
    launches0 = capi.launch_count()
    capi.check(L.lfp_synth_profile_begin(plan._h, 0b11))  # conv fwd + dgrad
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record()
    for _ in range(args.steps):
        eng.step(st, target)
    e1.record()
    barrier()
    import ctypes as C
    n = len(capi.KINDS)
    ms, cnt, fl, by = (C.c_double * n)(), (C.c_int64 * n)(), (C.c_double * n)(), (C.c_double * n)()
    capi.check(L.lfp_synth_profile_end(plan._h, ms, cnt, fl, by))
    launches = capi.launch_count() - launches0
    clocks = sampler.stop() if rank == 0 else None
    elapsed = torch.tensor([e0.elapsed_time(e1)], device=dev)
    if world > 1:
        dist.all_reduce(elapsed, op=dist.ReduceOp.MAX)
    ms_total = float(elapsed.item())

    # ---- end-to-end through the host-buffer call: H2D latents, fwd+bwd, D2H loss + d_wx ----
    wx_host = torch.empty(B, eng.dim).pin_memory()
    _, wx_dev = eng.embed(st["alpha"], st["key"])
    wx_host.copy_(wx_dev)
    loss_host = torch.empty(B).pin_memory()
    dwx_host = torch.empty(B, eng.dim).pin_memory()
    e2e_steps = max(3, min(args.steps, 10))
    eng.loss_and_grad_host(wx_host, target, loss_host, dwx_host)
    barrier()
    g0, g1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    g0.record()
    for _ in range(e2e_steps):
        eng.loss_and_grad_host(wx_host, target, loss_host, dwx_host)   # H2D + fwd + bwd + D2H, synchronous
    g1.record()
    barrier()
    e2e_s = torch.tensor([g0.elapsed_time(g1) * 1e-3], device=dev)     # device clock, max over ranks below
    if world > 1:
        dist.all_reduce(e2e_s, op=dist.ReduceOp.MAX)

    # ---- the path's one exchange: gather final losses / keys (NCCL), rank 0 picks per-image minima ----
    final = torch.cat([st["loss"][:, None], st["key"]], 1)
    if world > 1:
        gathered = [torch.empty_like(final) for _ in range(world)]
        dist.all_gather(gathered, final)
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    pk, pk_src = peaks()
    conv_ms = ms[0] + ms[1]
    conv_fl = fl[0] + fl[1]
    conv_by = by[0] + by[1]
    conv_launch = cnt[0] + cnt[1]
    achieved = conv_fl / (conv_ms * 1e-3) / 1e12 if conv_ms > 0 else 0.0
    peak = pk.get("bf16_tflops_sustained", pk.get("bf16_tflops"))
    peak_note = f"{pk_src} bf16 sustained"
    if prec == capi.PREC_TF32:
        peak, peak_note = peak / 2.0, f"{pk_src} bf16 sustained / 2 (kind::tf32 runs at half the bf16 rate)"
    else:
        peak_note += " (fp32 CUDA-core path: the tensor peak is not its bound)"
    value = world * B * args.steps / (ms_total * 1e-3)
    out = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
        "warmup": max(args.warmup, 3), "ms_per_step": ms_total / args.steps, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32" if prec == capi.PREC_FP32 else "tf32",
        "data": "synthetic",
        "config": {"workload": f"attribution_{size}px_n{B}_mse", "size": size, "trajectories_per_rank": B,
                   "loss": "mse", "key_len": 64, "shift": 448, "precision": args.precision,
                   "l2": "activations per step (>= 10 GB at 1024px) exceed the 126 MB L2; no flush needed"},
        "clocks": clocks,
        "e2e": {"value": world * B * e2e_steps / float(e2e_s.item()), "unit": UNIT,
                "h2d_bytes_per_step": B * eng.dim * 4, "d2h_bytes_per_step": B * eng.dim * 4 + B * 4,
                "note": "host-buffer call: latents in, loss + d(loss)/d(wx) out, per step"},
        "gpu_launches": int(launches),
        "roofline": {"kernel": "modulated-conv gather kernels (forward + data-gradient)", "bound": "tensor",
                     "achieved": achieved, "peak": peak, "unit": "TFLOP/s", "frac": achieved / peak if peak else None,
                     "traffic": CONV_DRAM_BYTES_PER_LAUNCH if (prec == capi.PREC_TF32 and size == 1024 and B == 20) else None,
                     "traffic_note": "dram__bytes_read.sum + dram__bytes_write.sum averaged over the conv launches of one "
                                     "step, ncu capture profiles/r01_conv_dram_traffic.md (bytes)",
                     "algorithmic_bytes_per_launch": conv_by / conv_launch if conv_launch else None,
                     "peak_source": peak_note, "launches": int(conv_launch),
                     "share_of_step": conv_ms / ms_total if ms_total else None,
                     "avg_launch_ms": conv_ms / conv_launch if conv_launch else None,
                     "algorithmic_gflop_per_launch": conv_fl / conv_launch / 1e9 if conv_launch else None},
    }
    if not args.no_cpu_baseline:
        cores = os.cpu_count() or 1
        torch.set_num_threads(cores)
        per = time_cpu(oracle_step_fn(size, torch), args.cpu_baseline_steps, warm=1)
        out["cpu_baseline"] = {"value": 1.0 / per, "unit": UNIT, "cores": cores, "kind": "port",
                               "sample": f"{args.cpu_baseline_steps} steps x 1 trajectory at {size}px "
                                         f"(synthesis fwd+bwd+MSE, oracle port on torch-CPU), after 1 warm-up"}
    print(json.dumps(out))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    a = parse()
    if a.impl == "reference":
        run_reference(a)
    else:
        run_ours(a)
