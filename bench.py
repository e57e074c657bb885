#!/usr/bin/env python
"""Benchmark of the attribution hot path (BASELINE.json metric: latent-opt steps/sec at 1024 px).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload NAME]

Workloads (``config.workload`` in the JSON line):
  attribution_1024px_n20_mse       BASELINE.json configs[2] per-image slice (default; the metric's configuration)
  attribution_512px_k128_n20_mse   configs[3]: 512 px, key_len 128, sigma 1.5, shift 384
  generation_1024px_b64            configs[1]: fingerprinted generation, batch 64, forward only (images/s)

A step = one Adam step of every trajectory in flight on a rank (the n Latin-hypercube guesses of one image):
fingerprint embed -> StyleGAN2 synthesis forward -> MSE-to-target -> synthesis backward to the latent -> Adam.
``value`` is trajectory-steps/s summed over all ranks with everything resident in HBM (each rank owns its own
image: weak scaling, no data-path collective; the only exchange is the final gather of keys and losses).
``e2e`` is the same step through ``AttributionEngine.step_host``: the trajectory state (alpha, key logits, Adam
moments) lives in pinned HOST buffers, is copied in before and out after every step together with the loss.
The roofline pass (CUDA events around every launch, recorded by the library on its launching stream) runs in
separate steps AFTER the timed regions.

``--impl reference`` times the UNMODIFIED reference (baseline/_ref/src, shipped by tools/ship_reference.py) on the
host cores: its own ``main.optimization`` (src/main.py:45-89) on one trajectory, CPU tensors -> ``upfirdn2d_native``
/ ``F.leaky_relu`` / MKL-DNN convs, MSE loss.  Falls back to the oracle port only when the copy is missing.
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

METRIC = "attribution latent-opt steps/sec (images x guesses) at 1024px"
UNIT = "trajectory-steps/s"

WORKLOADS = {
    "attribution_1024px_n20_mse": dict(kind="attribution", size=1024, key_len=64, shift=448, sigma=1.0, guesses=20),
    "attribution_512px_k128_n20_mse": dict(kind="attribution", size=512, key_len=128, shift=384, sigma=1.5, guesses=20),
    "generation_1024px_b64": dict(kind="generation", size=1024, key_len=64, shift=448, sigma=1.0, guesses=64),
    # the reference's default loss (LPIPS-VGG16, src/utils.py:44-50) on the native kernels; random VGG16 weights (no ImageNet
    # weights offline), target features cached per image
    "attribution_1024px_n20_lpips": dict(kind="attribution", size=1024, key_len=64, shift=448, sigma=1.0, guesses=20, loss="lpips"),
    "attribution_256px_n20_lpips": dict(kind="attribution", size=256, key_len=64, shift=448, sigma=1.0, guesses=20, loss="lpips"),
}
# whole-step bound of SURVEY.md 8d: (fwd + dgrad GFLOP, ideal fused fp32 GB) per trajectory-step
WHOLE_STEP = {256: (180.5, 0.83), 512: (238.7, 1.74), 1024: (297.0, 3.60)}


def add_paths():
    for p in (ROOT, PKG, os.path.join(ROOT, "tests")):
        if p not in sys.path:
            sys.path.insert(0, p)


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="attribution_1024px_n20_mse", choices=sorted(WORKLOADS))
    ap.add_argument("--guesses", type=int, default=0, help="override trajectories in flight per rank")
    ap.add_argument("--precision", default="tf32", choices=["fp32", "tf32"],
                    help="tf32 = tcgen05 tensor-core convs (the numerics the reference runs on GPU); fp32 = CUDA-core convs")
    ap.add_argument("--sustained-steps", type=int, default=200, help="steps of the sustained figure (0 = skip)")
    ap.add_argument("--fp32-steps", type=int, default=3, help="steps of the fp32-path record (0 = skip)")
    ap.add_argument("--cpu-budget", type=float, default=25.0, help="seconds of CPU work for the cpu_baseline leg")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--ref-budget", type=float, default=240.0, help="wall-clock budget of --impl reference, seconds")
    ap.add_argument("--python-steps", action="store_true", help="drive every launch of a step from Python (no CUDA graph)")
    ap.add_argument("--ref-port", action="store_true", help="--impl reference: time the oracle port instead of baseline/_ref")
    return ap.parse_args()


def peaks():
    try:
        return json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))), "measured"
    except Exception:
        return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0}, "fallback (B200_PROFILING.md)"


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


def synthetic_setup(size):
    """Random-init generator of the named architecture with perturbed zero-init parameters, a
    synthetic orthonormal PCA basis, N(0,1) noise maps (SURVEY.md 8d 'Synthetic inputs')."""
    import fixtures as fx
    params = fx.make_params(size, seed=1346)
    noise = fx.make_noise(size, seed=2002)
    pc, sigma, mean = fx.make_pca_basis(2)
    return params, noise, pc, sigma, mean


def config_of(args, wl):
    """Identical for both arms (the driver compares it); arm-specific facts go to top-level keys."""
    return {"workload": args.workload, "size": wl["size"], "trajectories_per_rank": wl["guesses"], "loss": wl.get("loss", "mse"),
            "key_len": wl["key_len"], "shift": wl["shift"], "sigma": wl["sigma"]}


# ------------------------------------------------------------------------------------------------
# reference arm (host cores)
# ------------------------------------------------------------------------------------------------
def reference_step_runner(wl, use_port):
    """Returns (run(k) -> seconds for k consecutive steps of ONE trajectory, kind, description)."""
    import torch
    add_paths()
    import fixtures as fx
    size = wl["size"]
    params, noise, pc, sigma, mean = synthetic_setup(size)
    if not use_port:
        from oracle import reference_harness as rh
        if rh.available():
            loop = rh.ReferenceLoop(size, params, noise, pc, sigma, mean, wl["key_len"], wl["shift"], wl["sigma"], device="cpu")
            lhs = rh.lhs_sample(1, loop.n_main, 300)

            def run(k):
                t0 = time.perf_counter()
                loop.run(lhs, k)
                return time.perf_counter() - t0

            return run, "reference", ("unmodified reference main.optimization (baseline/_ref/src), CPU tensors: upfirdn2d_native + "
                                      "F.leaky_relu + MKL-DNN convs, MSE loss, parameter gradients computed as the reference does")
    import oracle
    sp = fx.split_basis(pc, sigma, wl["key_len"], wl["shift"], wl["sigma"])
    alpha = sp["sigma_main"] * fx.seeded((sp["sigma_main"].shape[0], 1), 5)
    key = (fx.seeded((wl["key_len"], 1), 6) > 0).float()
    with torch.no_grad():
        target, _, _ = oracle.generate_with_alpha(params, size, alpha, sp["u_cap"], sp["v_cap"], sp["sigma_key"], mean, key, noise)
    a0 = sp["sigma_main"] * fx.seeded((sp["sigma_main"].shape[0], 1), 7)

    def render(wx):
        return oracle.generator_forward(params, [wx.reshape(1, -1)], size, input_is_latent=True, noise=noise)

    def run(k):
        t0 = time.perf_counter()
        oracle.attribute_one_guess(render, target, a0, sp["u_cap"], sp["v_cap"], sp["sigma_key"], mean, sp["max_alpha"],
                                   sp["min_alpha"], steps=k, key_len=wl["key_len"])
        return time.perf_counter() - t0

    return run, "port", "oracle port of main.optimization on torch-CPU (no parameter gradients), MSE loss"


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import torch
    wl = dict(WORKLOADS[args.workload])
    if wl["kind"] != "attribution":
        print(json.dumps({"impl": "reference", "unavailable": "the reference arm times the attribution step only"}))
        return
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    run, kind, how = reference_step_runner(wl, args.ref_port)
    t_first = run(1)                                   # also the first warm-up step
    per_guess = max(t_first, 1e-3)
    k = max(1, min(args.steps, int(args.ref_budget / per_guess) - args.warmup - 1))
    w = max(0, min(args.warmup - 1, int(0.2 * args.ref_budget / per_guess)))
    if w > 0:
        run(w)
    per = run(k) / k
    v = 1.0 / per
    sample = f"{k} steps x 1 trajectory at {wl['size']} px after {w + 1} warm-up steps; {how}"
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus, "steps": k,
        "warmup": w + 1, "ms_per_step": per * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic", "host_only": True,
        "config": config_of(args, wl), "precision": "fp32 (torch-CPU)",
        "sample_note": "one trajectory per step (the reference runs its trajectories one at a time)",
        "cpu_baseline": {"value": v, "unit": UNIT, "cores": cores, "kind": kind, "sample": sample},
        "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }))


def cpu_baseline_subprocess(args):
    """The cpu_baseline leg: the reference arm in a fresh process (the reference's module names `model`, `op`,
    `generator`, `main` collide with this package's drop-in modules inside one interpreter)."""
    cmd = [sys.executable, os.path.abspath(__file__), "--impl", "reference", "--workload", args.workload, "--steps", "3",
           "--warmup", "1", "--ref-budget", str(args.cpu_budget)]
    env = {k: v for k, v in os.environ.items() if k not in ("RANK", "WORLD_SIZE", "LOCAL_RANK")}
    try:
        out = subprocess.run(cmd, capture_output=True, text=True, timeout=max(240.0, 12 * args.cpu_budget), env=env)
        for line in reversed(out.stdout.strip().splitlines()):
            if line.startswith("{"):
                return json.loads(line).get("cpu_baseline")
        return {"error": (out.stderr or out.stdout)[-300:]}
    except Exception as e:  # noqa: BLE001
        return {"error": repr(e)[:300]}


# ------------------------------------------------------------------------------------------------
# our arm
# ------------------------------------------------------------------------------------------------
def roofline_from_launches(kinds, ms, flops, bytes_, peak_tf, hbm_gbs, conv_kinds=(0, 1)):
    """Per-launch bound = max(flops / tensor peak, bytes / HBM peak); frac = sum(bound) / sum(measured)."""
    tot_ms = tot_bound = tot_fl = tot_by = 0.0
    n = n_hbm = 0
    for k, t, f, b in zip(kinds, ms, flops, bytes_):
        if k not in conv_kinds:
            continue
        tf_ms = f / (peak_tf * 1e12) * 1e3
        hb_ms = b / (hbm_gbs * 1e9) * 1e3
        tot_bound += max(tf_ms, hb_ms)
        n_hbm += hb_ms >= tf_ms
        tot_ms += t
        tot_fl += f
        tot_by += b
        n += 1
    return dict(ms=tot_ms, bound_ms=tot_bound, flops=tot_fl, bytes=tot_by, launches=n, hbm_bound_launches=n_hbm)


def read_traffic(workload, precision):
    """Average DRAM bytes (read + write) per conv launch from the committed ncu capture summary, or None."""
    try:
        t = json.load(open(os.path.join(ROOT, "profiles", "roofline_traffic.json")))
        e = t.get(f"{workload}/{precision}")
        return (e["dram_bytes_per_conv_launch"], e["source"]) if e else (None, None)
    except Exception:
        return None, None


def run_ours(args):
    add_paths()
    import ctypes as C
    import numpy as np
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
    wl = dict(WORKLOADS[args.workload])
    if args.guesses:
        wl["guesses"] = args.guesses
    size, B = wl["size"], wl["guesses"]
    prec = capi.PREC_TF32 if args.precision == "tf32" else capi.PREC_FP32

    params, noise, pc, sigma, mean = synthetic_setup(size)
    plan = SynthesisPlan(size, device=dev)
    plan.load(params)
    loss_kind = wl.get("loss", "mse")
    lpips_params = None
    if loss_kind == "lpips":
        lpips_params = fx.make_vgg_params(seed=1346)   # seeded random-init VGG16 weights + non-negative heads
    eng = AttributionEngine(plan, noise, pc, sigma, mean, key_len=wl["key_len"], shift=wl["shift"], sigma=wl["sigma"],
                            sd=1.0, lr=0.2, precision=prec, loss=loss_kind, lpips_params=lpips_params)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, n):
        """n calls of fn between barriers, device time (CUDA events on the current stream), max over ranks, ms."""
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        barrier()
        e0.record()
        for _ in range(n):
            fn()
        e1.record()
        barrier()
        t = torch.tensor([e0.elapsed_time(e1)], device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    if wl["kind"] == "generation":
        return run_generation(args, wl, eng, plan, dev, world, rank, local, timed, barrier)

    # per-rank image (independent units): target from a seeded alpha and key, as generate_with_alpha does
    alpha_t = (eng.sigma_main * fx.seeded((1, eng.n_main), 100 + rank).to(dev))
    key_t = (fx.seeded((1, wl["key_len"]), 200 + rank) > 0).to(dev)
    _, wx_t = eng.embed_with_key(alpha_t, key_t)
    target = eng.render(wx_t).clone()
    rs = np.random.RandomState(300 + rank)
    lhs = torch.from_numpy(np.stack([(rs.permutation(B) + 0.5) / B for _ in range(eng.n_main)], 1)).float()
    st = eng.init_state(eng.alpha0_from_lhs(lhs))
    W = max(args.warmup, 3)
    # the whole Adam step is one native call replayed as a CUDA graph (lfp_attrib_run); --python-steps drives the same
    # kernels launch by launch from Python instead
    stepper = None if args.python_steps else eng.native_stepper(st, target, max_steps=W + args.steps + args.sustained_steps + 64)
    step_fn = (lambda: eng.step(st, target)) if stepper is None else (lambda: stepper.run(1))
    for _ in range(W):
        step_fn()

    # ---- timed region 1: device-resident (value) ----
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    launches0 = capi.launch_count()
    ms_total = timed(step_fn, args.steps)
    launches = capi.launch_count() - launches0
    clocks = sampler.stop() if rank == 0 else None
    value = world * B * args.steps / (ms_total * 1e-3)
    # a replayed graph launches its kernels without passing through the library's launch counter: count one step's
    # kernels with an eager step outside the timers
    if stepper is not None:
        l0 = capi.launch_count()
        stepper.run(1, graph=False)
        torch.cuda.synchronize()
        launches_per_step = capi.launch_count() - l0
    else:
        launches_per_step = launches / args.steps

    # ---- timed region 2: sustained figure (the real loop is 2000 steps; power cap bites after ~1 s) ----
    sustained = None
    if args.sustained_steps > 0:
        ms_s = timed(step_fn, args.sustained_steps)
        sustained = {"value": world * B * args.sustained_steps / (ms_s * 1e-3), "unit": UNIT, "steps": args.sustained_steps,
                     "ms_per_step": ms_s / args.sustained_steps}

    # ---- timed region 3: end to end, trajectory state in pinned host buffers, every step ----
    hs = eng.host_state(st)
    dst = eng.step_host(hs, target, st, stepper)
    e2e_steps = max(3, min(args.steps, 20))
    ms_e = timed(lambda: eng.step_host(hs, target, dst, stepper), e2e_steps)
    state_bytes = sum(hs[k].numel() * 4 for k in eng.STATE_KEYS)
    e2e = {"value": world * B * e2e_steps / (ms_e * 1e-3), "unit": UNIT, "h2d_bytes_per_step": state_bytes,
           "d2h_bytes_per_step": state_bytes + B * 4, "steps": e2e_steps,
           "note": "AttributionEngine.step_host: alpha, key logits and Adam moments H2D from pinned memory, full step "
                   f"(embed, synthesis fwd, {loss_kind.upper()} loss fwd + bwd, synthesis bwd, Adam), state + loss D2H, synchronous, "
                   "every step"}

    # ---- roofline pass: per-launch CUDA events, outside every timed region ----
    L = capi.lib()
    prof_steps = 3
    nk = len(capi.KINDS)
    capi.check(L.lfp_synth_profile_begin(plan._h, (1 << nk) - 1))
    barrier()
    for _ in range(prof_steps):
        eng.step(st, target)
    barrier()
    pms, pcnt, pfl, pby = (C.c_double * nk)(), (C.c_int64 * nk)(), (C.c_double * nk)(), (C.c_double * nk)()
    capi.check(L.lfp_synth_profile_end(plan._h, pms, pcnt, pfl, pby))
    nl = L.lfp_synth_profile_launches(plan._h, 0, None, None, None, None)
    kinds, lms, lfl, lby = (C.c_int * nl)(), (C.c_float * nl)(), (C.c_double * nl)(), (C.c_double * nl)()
    L.lfp_synth_profile_launches(plan._h, nl, kinds, lms, lfl, lby)
    classes = {capi.KINDS[k]: {"ms_per_step": pms[k] / prof_steps, "launches_per_step": pcnt[k] / prof_steps,
                               "algorithmic_gb_per_step": pby[k] / prof_steps / 1e9,
                               "gbs": (pby[k] / (pms[k] * 1e-3) / 1e9) if pms[k] > 0 else None}
               for k in range(nk) if pcnt[k]}

    # ---- fp32 (exact-parity) path, a few steps, reported inside the same line ----
    fp32 = None
    if args.fp32_steps > 0 and prec != capi.PREC_FP32 and loss_kind == "mse":
        eng.precision = capi.PREC_FP32
        st32 = eng.init_state(eng.alpha0_from_lhs(lhs))
        eng.step(st32, target)
        ms32 = timed(lambda: eng.step(st32, target), args.fp32_steps)
        fp32 = {"value": world * B * args.fp32_steps / (ms32 * 1e-3), "unit": UNIT, "steps": args.fp32_steps,
                "dtype": "f32", "note": "LFP_PREC_FP32: CUDA-core FFMA convolutions, the exact-parity path"}
        eng.precision = prec

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
    hbm = pk["hbm_gbs"]
    peak_tc = pk.get("bf16_tflops_sustained", pk.get("bf16_tflops"))
    if prec == capi.PREC_TF32:
        peak_tc, peak_note = peak_tc / 2.0, f"{pk_src} bf16 sustained / 2 (kind::tf32 runs at half the bf16 rate)"
    else:
        peak_note = f"{pk_src} bf16 sustained (fp32 CUDA-core path: the tensor peak is not its bound)"
    r = roofline_from_launches(list(kinds), list(lms), list(lfl), list(lby), peak_tc, hbm)
    hbm_bound = r["hbm_bound_launches"] * 2 >= r["launches"]
    traffic, traffic_src = read_traffic(args.workload, args.precision)
    gflop, gbytes = WHOLE_STEP.get(size, (None, None))
    step_ms = ms_total / args.steps
    whole = None
    if gflop:
        t_fl = B * gflop * 1e9 / (peak_tc * 1e12) * 1e3
        t_by = B * gbytes * 1e9 / (hbm * 1e9) * 1e3
        whole = {"bound": "hbm" if t_by >= t_fl else "tensor", "bound_ms_per_step": max(t_fl, t_by), "ms_per_step": step_ms,
                 "frac": max(t_fl, t_by) / step_ms,
                 "note": f"SURVEY.md 8d: {gflop} GFLOP and {gbytes} GB of ideal fused fp32 traffic per trajectory-step"}
    out = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": W,
        "ms_per_step": step_ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32" if prec == capi.PREC_FP32 else "tf32", "data": "synthetic",
        "config": config_of(args, wl), "precision": args.precision,
        "l2_note": "inputs larger than L2: the activations a step streams (>= 10 GB at 1024 px, B = 20) exceed the 126 MB L2; no flush",
        "clocks": clocks, "e2e": e2e, "sustained": sustained, "gpu_launches": int(launches_per_step * args.steps),
        "launches_per_step": launches_per_step,
        "step_driver": "python (one ctypes call per kernel group)" if stepper is None else
                       "lfp_attrib_run: the whole step captured once as a CUDA graph and replayed (one cudaGraphLaunch per step)",
        "roofline": {
            "kernel": "conv_tc_kernel: modulated-conv gather kernels, forward + data-gradient (all launches of a step)",
            "bound": "hbm" if hbm_bound else "tensor",
            "achieved": (r["bytes"] / (r["ms"] * 1e-3) / 1e9) if hbm_bound else (r["flops"] / (r["ms"] * 1e-3) / 1e12),
            "peak": hbm if hbm_bound else peak_tc, "unit": "GB/s" if hbm_bound else "TFLOP/s",
            "frac": r["bound_ms"] / r["ms"] if r["ms"] > 0 else None,
            "frac_definition": "sum over conv launches of max(flops / tensor peak, algorithmic bytes / HBM peak) divided by the "
                               "sum of their measured durations; achieved / peak are quoted for the binding term of the "
                               "majority of launches",
            "traffic": traffic, "traffic_source": traffic_src,
            "algorithmic_bytes_per_launch": r["bytes"] / r["launches"] if r["launches"] else None,
            "algorithmic_gflop_per_launch": r["flops"] / r["launches"] / 1e9 if r["launches"] else None,
            "launches_per_step": r["launches"] / prof_steps, "hbm_bound_launches_per_step": r["hbm_bound_launches"] / prof_steps,
            "avg_launch_ms": r["ms"] / r["launches"] if r["launches"] else None,
            "share_of_step": (r["ms"] / prof_steps) / step_ms,
            "tensor_tflops": r["flops"] / (r["ms"] * 1e-3) / 1e12 if r["ms"] > 0 else None,
            "peak_source": f"HBM {hbm} GB/s ({pk_src}); tensor {peak_tc:.1f} TFLOP/s = {peak_note}",
            "classes": classes, "whole_step": whole,
        },
    }
    if fp32:
        out["fp32"] = fp32
    if not args.no_cpu_baseline:
        out["cpu_baseline"] = cpu_baseline_subprocess(args)
    print(json.dumps(out))
    if world > 1:
        dist.destroy_process_group()


def run_generation(args, wl, eng, plan, dev, world, rank, local, timed, barrier):
    """BASELINE.json configs[1]: fingerprinted generation (src/generator.py:185-198, 69-107), batch 64 at 1024 px, forward
    only through the forward-only workspace.  value = images/s; e2e = alpha / key H2D from pinned memory, image D2H."""
    import torch
    import torch.distributed as dist
    import fixtures as fx
    from lfp_native import capi
    B = wl["guesses"]
    alpha_h = (eng.sigma_main.cpu() * fx.seeded((B, eng.n_main), 100 + rank)).pin_memory()
    key_h = (fx.seeded((B, wl["key_len"]), 200 + rank) > 0).float().pin_memory()
    alpha, key = alpha_h.to(dev), key_h.to(dev)

    def gen():
        _, wx = eng.embed_with_key(alpha, key)
        return eng.render(wx)

    W = max(args.warmup, 3)
    for _ in range(W):
        gen()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    launches0 = capi.launch_count()
    ms_total = timed(gen, args.steps)
    launches = capi.launch_count() - launches0
    clocks = sampler.stop() if rank == 0 else None
    img_h = torch.empty(B, 3, wl["size"], wl["size"]).pin_memory()

    def gen_host():
        a, k = alpha_h.to(dev, non_blocking=True), key_h.to(dev, non_blocking=True)
        _, wx = eng.embed_with_key(a, k)
        img_h.copy_(eng.render(wx), non_blocking=True)
        torch.cuda.current_stream().synchronize()

    gen_host()
    e2e_steps = max(3, min(args.steps, 10))
    ms_e_sync = timed(gen_host, e2e_steps)

    # The same call sequence with the image read-back of step i overlapping the render of step i + 1: two device images, two
    # pinned host images, a copy stream.  Every step's 805 MB image is in host memory before the timed region ends (the last
    # event waits for the last copy).
    copy_stream = torch.cuda.Stream()
    main_stream = torch.cuda.current_stream()
    img_d = [torch.empty(B, 3, wl["size"], wl["size"], device=dev) for _ in range(2)]
    img_hh = [img_h, torch.empty(B, 3, wl["size"], wl["size"]).pin_memory()]
    copied = [torch.cuda.Event(), torch.cuda.Event()]

    def gen_host_overlapped(i):
        a, k = alpha_h.to(dev, non_blocking=True), key_h.to(dev, non_blocking=True)
        _, wx = eng.embed_with_key(a, k)
        main_stream.wait_event(copied[i & 1])          # the device image of two steps ago has left
        img_d[i & 1].copy_(eng.render(wx))
        rendered = torch.cuda.Event()
        rendered.record(main_stream)
        copy_stream.wait_event(rendered)
        with torch.cuda.stream(copy_stream):
            img_hh[i & 1].copy_(img_d[i & 1], non_blocking=True)
            copied[i & 1].record(copy_stream)

    for i in range(2):
        gen_host_overlapped(i)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record()
    for i in range(e2e_steps):
        gen_host_overlapped(i)
    main_stream.wait_event(copied[0])
    main_stream.wait_event(copied[1])
    e1.record()
    barrier()
    t = torch.tensor([e0.elapsed_time(e1)], device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_e = float(t.item())
    assert torch.equal(img_hh[0], img_hh[1]) and torch.isfinite(img_hh[0]).all()   # same inputs every step: same images
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    print(json.dumps({
        "metric": "fingerprinted generation images/sec at 1024px, batch 64", "value": world * B * args.steps / (ms_total * 1e-3),
        "unit": "images/s", "n_gpus": world, "steps": args.steps, "warmup": W, "ms_per_step": ms_total / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": args.precision, "data": "synthetic",
        "config": config_of(args, wl), "precision": args.precision, "clocks": clocks, "gpu_launches": int(launches),
        "workspace_gb": plan.workspace_bytes(B, forward_only=True) / 1e9,
        "e2e": {"value": world * B * e2e_steps / (ms_e * 1e-3), "unit": "images/s",
                "h2d_bytes_per_step": int(alpha_h.numel() + key_h.numel()) * 4, "d2h_bytes_per_step": img_h.numel() * 4,
                "note": "alpha / key H2D from pinned memory, render, image D2H into pinned memory every step; the read-back of step i "
                        "runs on a copy stream under the render of step i + 1 (two device and two host images)",
                "synchronous": world * B * e2e_steps / (ms_e_sync * 1e-3)},
    }))
    if world > 1:
        dist.destroy_process_group()


def main():
    a = parse()
    world_env = os.environ.get("WORLD_SIZE")
    if a.impl == "ours" and a.gpus > 1 and world_env is None:
        # `python bench.py --gpus N` without torchrun: launch the ranks ourselves
        port = 29500 + (os.getpid() % 2000)
        cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={a.gpus}", "--master-addr",
               "127.0.0.1", "--master-port", str(port), os.path.abspath(__file__)] + sys.argv[1:]
        raise SystemExit(subprocess.call(cmd))
    if a.impl == "ours" and world_env is not None and int(world_env) != a.gpus:
        raise SystemExit(f"bench.py: --gpus {a.gpus} does not match WORLD_SIZE={world_env}")
    if a.impl == "reference":
        run_reference(a)
    else:
        run_ours(a)


if __name__ == "__main__":
    main()
