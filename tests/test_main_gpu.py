"""The sharded attribution driver (main.py: the experiment of src/main.py:93-154) executed on the GPU: one process, and -
when the box has two GPUs - two ranks over NCCL; the gathered (loss, key logits, alpha) rows must be bit-identical for any
world size and any per-launch batch (trajectories are independent: SURVEY.md 8e)."""
import json
import os
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
MAIN = os.path.join(ROOT, "attributing-image-generative-models-using-latent-fingerprints-sg2_b200", "main.py")
ARGS = ["--img_size", "64", "--sample_size", "3", "--n", "4", "--steps", "6", "--key_len", "64", "--shift", "448"]


def run_main(tmp, tag, world, extra=()):
    rows = os.path.join(tmp, f"rows_{tag}.pt")
    cmd = [sys.executable, MAIN] + ARGS + ["--save_dir", os.path.join(tmp, tag), "--dump_rows", rows] + list(extra)
    # one process per GPU, rendezvous through the environment (what torchrun sets; torchrun's own argparse would claim the
    # reference's `--n` flag as an abbreviation of its `--nnodes` / `--nproc-per-node`)
    procs = []
    for r in range(world):
        env = dict(os.environ, RANK=str(r), LOCAL_RANK=str(r), WORLD_SIZE=str(world), MASTER_ADDR="127.0.0.1",
                   MASTER_PORT=str(29600 + world))
        if world == 1:
            for k in ("RANK", "LOCAL_RANK", "WORLD_SIZE"):
                env.pop(k)
        procs.append(subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True, env=env))
    outs = [p.communicate(timeout=600) for p in procs]
    for p, (so, se) in zip(procs, outs):
        assert p.returncode == 0, se[-2000:]
    summary = json.loads([l for l in outs[0][0].splitlines() if l.startswith("{")][-1])
    assert os.path.isfile(os.path.join(tmp, tag, "result.txt"))
    return torch.load(rows), summary


def test_main_runs_and_batching_does_not_change_rows(tmp_path):
    rows, summary = run_main(str(tmp_path), "w1", 1)
    assert rows.shape == (12, 1 + 64 + 448) and torch.isfinite(rows).all()
    assert summary["world"] == 1 and summary["images"] == 3
    rows_b3, _ = run_main(str(tmp_path), "w1b3", 1, ["--batch", "3"])     # 12 pairs in batches of 3 instead of 4
    assert torch.equal(rows, rows_b3)
    rows_py, _ = run_main(str(tmp_path), "w1fp32", 1, ["--precision", "fp32"])
    assert not torch.equal(rows, rows_py)                                  # a different arithmetic path really ran
    assert (rows[:, 0] - rows_py[:, 0]).abs().max() < 0.05 * rows_py[:, 0].abs().max()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs two GPUs (gpurun --gpus 2)")
def test_two_ranks_give_the_single_rank_rows_bit_for_bit(tmp_path):
    rows1, _ = run_main(str(tmp_path), "w1", 1)
    rows2, summary = run_main(str(tmp_path), "w2", 2)
    assert summary["world"] == 2
    assert torch.equal(rows1, rows2)
