"""Sharded attribution driver - the experiment of src/main.py:93-154 on N GPUs of one box.

    python main.py --img_size 1024 --sample_size 100 --n 20 --steps 2000 --key_len 64 --shift 448
    torchrun --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 main.py ... --guesses 20   (not --n: torchrun's parser claims it)

Every (image, guess) pair is an independent trajectory (sharding.py); each rank runs its slice in
batches through ``attribution.AttributionEngine`` and the final rows are gathered once over NCCL.
Per-image targets, keys and Latin-hypercube guesses derive from ``--seed`` and the image index only,
so any world size produces the same rows bit for bit.  Loss: MSE (the LPIPS-VGG16 loss of the
reference is outside the native path).  Random-init weights unless ``--ckpt`` is given.
"""
import argparse
import json
import os
import time

import numpy as np
import torch
import torch.distributed as dist

import sharding
from attribution import AttributionEngine
from lfp_native import capi
from generator import GetGen, get_noise


def parse():
    ap = argparse.ArgumentParser(description="latent-fingerprint attribution (B200-native path)")
    ap.add_argument("--ckpt", type=str, default=None)
    ap.add_argument("--img_size", type=int, default=256)
    ap.add_argument("--sample_size", type=int, default=100)
    ap.add_argument("--sd", type=int, default=1)
    ap.add_argument("--steps", type=int, default=2000)
    ap.add_argument("--n", "--guesses", dest="n", type=int, default=20,
                    help="Latin-hypercube guesses per image (src/params.py:17); spell it --guesses under torchrun, whose own "
                         "argument parser claims --n as an abbreviation of --nnodes / --nproc-per-node")
    ap.add_argument("--key_len", type=int, default=64)
    ap.add_argument("--save_dir", type=str, default="../result/")
    ap.add_argument("--lr", type=float, default=0.2)
    ap.add_argument("--shift", type=int, default=448)
    ap.add_argument("--sigma", type=float, default=1.0)
    ap.add_argument("--augmentation", type=str, default="None", help="Augmentation method: Noise, Blur, Jpeg, Combination")
    ap.add_argument("--jpeg_quality", type=int, default=50)
    ap.add_argument("--noise_sigma", type=float, default=0.1)
    ap.add_argument("--blur_sigma", type=float, default=0.5)
    ap.add_argument("--batch", type=int, default=0, help="trajectories per launch sequence (default: n)")
    ap.add_argument("--seed", type=int, default=1346)
    ap.add_argument("--dump_rows", type=str, default=None, help="rank 0 saves the gathered [pairs, 1 + key_len + n_main] rows here (torch.save)")
    ap.add_argument("--precision", default="tf32", choices=["tf32", "fp32"],
                    help="convolution arithmetic: tf32 = tcgen05 tensor cores (what the reference's cuDNN convs use by default), "
                         "fp32 = CUDA-core FFMA")
    return ap.parse_args()


def image_fixture(opt, gen: GetGen, image: int):
    """Target latent, key and LHS guesses of one image from (seed, image) only."""
    rs = np.random.RandomState((opt.seed * 100003 + image) % (2 ** 31 - 1))
    n_main = gen.num_main_pc
    alpha = torch.from_numpy(rs.standard_normal((n_main, 1)).astype(np.float32)).to(gen.device) * gen.sigma_448
    key = torch.from_numpy(rs.randint(0, 2, (opt.key_len, 1))).to(gen.device)
    lhs = np.stack([(rs.permutation(opt.n) + 0.5) / opt.n for _ in range(n_main)], 1).astype(np.float32)
    return alpha, key, torch.from_numpy(lhs).to(gen.device)


def main():
    opt = parse()
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    torch.manual_seed(opt.seed)
    np.random.seed(2022)  # Generator.__init__ of the reference (src/model.py:404)
    gen = GetGen(opt.img_size, opt.key_len, opt.shift, opt.sigma, opt.sd, ckpt=opt.ckpt, device=dev, seed=opt.seed,
                 augmentation=opt.augmentation, noise_sigma=opt.noise_sigma, blur_sigma=opt.blur_sigma,
                 jpeg_quality=opt.jpeg_quality)
    noise = get_noise(opt.img_size, dev)
    plan = gen.g_ema._plan()
    eng = AttributionEngine(plan, noise, gen.pc, gen.sigma_512, gen.latent_mean, opt.key_len, opt.shift, opt.sigma,
                            opt.sd, opt.lr, precision=capi.PREC_TF32 if opt.precision == "tf32" else capi.PREC_FP32)
    pairs = sharding.trajectory_list(opt.sample_size, opt.n)
    mine = sharding.partition(len(pairs), rank, world)
    B = opt.batch or opt.n
    cache = {}

    def fixture(i):
        if i not in cache:
            alpha, key, lhs = image_fixture(opt, gen, i)
            _, wx = eng.embed_with_key(alpha.t(), key.t())
            cache.clear()  # consecutive pairs share an image; keep one
            cache[i] = (gen.augmentation(eng.render(wx).clone()), key, lhs)   # attack once per target image (src/main.py:124)
        return cache[i]

    rows, t0 = [], time.time()
    for b in sharding.batches(mine, B):
        targets, alpha0 = [], []
        for t in b:
            i, g = pairs[t]
            tgt, _, lhs = fixture(i)
            targets.append(tgt)
            alpha0.append(eng.alpha0_from_lhs(lhs[g:g + 1]))
        st = eng.run(torch.cat(alpha0), torch.cat(targets), opt.steps)
        rows.append(torch.cat([st["loss"][:, None], st["key"], st["alpha"]], 1))
    local_rows = torch.cat(rows) if rows else torch.empty(0, 1 + opt.key_len + gen.num_main_pc, device=dev)
    full = sharding.gather_rows(local_rows, len(pairs), rank, world)
    torch.cuda.synchronize()
    if rank == 0:
        if opt.dump_rows:
            torch.save(full.cpu(), opt.dump_rows)
        best, keys, _ = sharding.select_best(full, opt.n, opt.key_len)
        true = torch.cat([image_fixture(opt, gen, i)[1].t() for i in range(opt.sample_size)])
        acc = sharding.bit_accuracy(keys, true)
        dt = time.time() - t0
        os.makedirs(opt.save_dir, exist_ok=True)
        with open(os.path.join(opt.save_dir, "result.txt"), "w") as f:
            for i in range(opt.sample_size):
                f.write(f"\n sample index: {i}, bit acc: {acc[i].item()}, attribution acc: "
                        f"{(acc[: i + 1] == 1).float().mean().item()}")
        print(json.dumps({"images": opt.sample_size, "guesses": opt.n, "steps": opt.steps, "world": world,
                          "seconds": dt, "trajectory_steps_per_s": len(pairs) * opt.steps / dt,
                          "mean_bit_acc": acc.mean().item(), "attribution_acc": (acc == 1).float().mean().item()}))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
