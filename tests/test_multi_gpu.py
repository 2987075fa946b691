"""Multi-process tests on REAL GPUs over NCCL, launched from ``pytest -m gpu``: each test spawns ``torchrun`` itself when the
box shows enough GPUs and is skipped otherwise (a 1-GPU box runs none of them).  They are the GPU twins of the reference's
tests/test_training/test_train.py:22-83 (2 GPUs, checkpoint / resume) and :115-206 (2 DiLoCo workers + resume) plus the
numerical checks of every NVLink outer-step transport and of the straggler policies."""
import os
import pickle
import socket
import subprocess
import sys

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
NGPU = torch.cuda.device_count() if torch.cuda.is_available() else 0


def need(n):
    return pytest.mark.skipif(NGPU < n, reason=f"needs {n} GPUs, {NGPU} visible")


def free_port() -> int:
    with socket.socket(socket.AF_INET, socket.SOCK_STREAM) as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def torchrun(nproc: int, target: list[str], timeout: int = 900, extra_env: dict | None = None) -> str:
    env = dict(os.environ, PYTHONPATH=ROOT, WANDB_MODE="disabled", OMP_NUM_THREADS="4", **(extra_env or {}))
    env.pop("CUDA_VISIBLE_DEVICES", None) if env.get("CUDA_VISIBLE_DEVICES") == "" else None
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={nproc}", "--master-addr", "127.0.0.1",
           "--master-port", str(free_port()), *target]
    res = subprocess.run(cmd, env=env, capture_output=True, text=True, timeout=timeout)
    if res.returncode != 0:
        pytest.fail(f"{' '.join(cmd)}\n{res.stdout[-4000:]}\n{res.stderr[-3000:]}")
    return res.stdout + "\n" + res.stderr


@need(2)
def test_outer_step_transports_match_the_reference_loop_nccl():
    """Every outer-step transport (flat NCCL, the three fused NVLink kernels, peer loads/stores, bf16 window, the five
    codecs) against a single-process simulation of train_diloco_torch.py:336-353; parameters AND outer momentum must be
    identical on all workers afterwards."""
    n = 8 if NGPU >= 8 else (4 if NGPU >= 4 else 2)
    out = torchrun(n, [os.path.join(ROOT, "tests", "dist_workers", "outer_equiv.py")])
    assert "FAIL" not in out and out.count(" OK") >= 12, out[-3000:]
    assert "fused=sharded" in out or "fused=pipelined" in out or "fused=sequential" in out, out[-3000:]


@need(3)
def test_wait_for_all_straggler_is_skipped_nccl():
    """WAIT_FOR_ALL + ODB_FAULT_INJECT on GPUs: N-1 workers finish the epoch without the straggler after
    ``timeout_waiting_for_peers``; it closes the epoch alone and the next full round repairs the drift."""
    n = 4 if NGPU >= 4 else 3
    out = torchrun(n, [os.path.join(ROOT, "tests", "dist_workers", "straggler_wait_for_all.py")], extra_env={"ODB_TEST_DEVICE": "cuda"})
    assert "FAIL" not in out and out.count("ALL OK") == n, out[-3000:]


@need(3)
def test_wait_for_all_straggler_with_the_fused_nvlink_kernel():
    """Same with the fused NVLink outer step for the full rounds (sharded kernel: the partial round has to wait for the
    background momentum all-gather before it reads the full outer state)."""
    n = 4 if NGPU >= 4 else 3
    out = torchrun(n, [os.path.join(ROOT, "tests", "dist_workers", "straggler_wait_for_all.py")],
                   extra_env={"ODB_TEST_DEVICE": "cuda", "ODB_TEST_FUSED": "1"})
    assert "FAIL" not in out and out.count("ALL OK") == n, out[-3000:]


@need(3)
def test_elastic_no_wait_rounds_nccl():
    """NO_WAIT elastic rounds (partial round, solo round, drift repair, state download by a lagging peer, post-resync full
    round) on 3 GPUs over NCCL point-to-point transfers."""
    out = torchrun(3, [os.path.join(ROOT, "tests", "dist_workers", "elastic_rounds.py")], extra_env={"ODB_TEST_DEVICE": "cuda"})
    assert "FAIL" not in out and out.count("ALL OK") == 3, out[-3000:]


BASE = ["--path_model", "14m", "--fake_data", "--no-torch_compile", "--lr", "1e-2", "--per_device_train_batch_size", "4",
        "--total_batch_size", "16", "--seq_length", "128", "--metric_logger_type", "dummy", "--precision", "bf16-mixed",
        "--warmup_steps", "4", "--total_steps", "100"]


def _load(path):
    with open(path, "rb") as f:
        return {d["step"]: (d["Loss"], d["lr"]) for d in pickle.load(f)}


@need(2)
@pytest.mark.parametrize("sharding", ["NO_SHARD", "SHARD_GRAD_OP", "FULL_SHARD"])
def test_ckpt_resume_data_parallel_2gpu(tmp_path, sharding):
    """reference tests/test_training/test_train.py:22-83 on our kernels: 2 GPUs, save at 4/8, resume from 8, atol 1e-3."""
    ckpt = f"{tmp_path}/ckpt"
    log1, log2 = f"{tmp_path}/log1.pkl", f"{tmp_path}/log2.pkl"
    common = ["-m", "opendiloco_b200.train_fsdp"] + BASE + ["--max_steps", "12", "--sharding_strategy", sharding]
    torchrun(2, common + ["--ckpt.path", ckpt, "--ckpt.interval", "4", "--project", log1])
    torchrun(2, common + ["--ckpt.path", ckpt, "--ckpt.resume", f"{ckpt}/model_step_8", "--project", log2])
    a, b = _load(log1), _load(log2)
    assert set(a) & set(b) == {9, 10, 11, 12}
    for s in (9, 10, 11, 12):
        assert np.allclose(a[s][0], b[s][0], atol=1e-3), f"Loss at step {s} is different: {a[s][0]} vs {b[s][0]}"
        assert a[s][1] == b[s][1], f"Lr at step {s} is different"


@need(2)
def test_ckpt_resume_two_diloco_workers_2gpu(tmp_path):
    """reference tests/test_training/test_train.py:115-206: 2 DiLoCo workers (fused NVLink outer step), resume at a step
    that is not a multiple of H, atol 1e-2."""
    ckpt = f"{tmp_path}/ckpt"
    log1, log2 = f"{tmp_path}/log1.pkl", f"{tmp_path}/log2.pkl"
    hv = ["--hv.local_steps", "5", "--hv.galaxy_size", "2", "--hv.skip_load_from_peers", "--hv.fail_rank_drop", "--hv.matchmaking_time", "1",
          "--total_batch_size", "8", "--max_steps", "16"]
    common = ["-m", "opendiloco_b200.train_fsdp"] + BASE + hv
    torchrun(2, common + ["--ckpt.path", ckpt, "--ckpt.interval", "7", "--project", log1])
    assert os.path.isdir(f"{ckpt}/model_step_7/diloco_rank_0") and os.path.isdir(f"{ckpt}/model_step_7/diloco_rank_1")
    torchrun(2, common + ["--ckpt.path", ckpt, "--ckpt.resume", f"{ckpt}/model_step_7", "--project", log2])
    a, b = _load(log1), _load(log2)
    assert set(a) & set(b) == set(range(8, 17))
    for s in range(8, 17):
        assert np.allclose(a[s][0], b[s][0], atol=1e-2), f"Loss at step {s} is different: {a[s][0]} vs {b[s][0]}"
        assert a[s][1] == b[s][1]


@need(4)
def test_hybrid_two_workers_by_two_gpus_no_wait(tmp_path):
    """2 DiLoCo workers x 2 GPUs (ZeRO-2 inside the worker), NO_WAIT: every local rank runs its own outer group; their
    board keys are namespaced per group, so both groups elect a leader and complete their rounds."""
    hv = ["--hv.local_steps", "3", "--hv.galaxy_size", "2", "--hv.skip_load_from_peers", "--hv.matchmaking_time", "1",
          "--hv.all_reduce_strategy", "NO_WAIT", "--total_batch_size", "8", "--max_steps", "9",
          "--sharding_strategy", "SHARD_GRAD_OP", "--project", f"{tmp_path}/log.pkl"]
    torchrun(4, ["-m", "opendiloco_b200.train_fsdp"] + BASE + hv)
    losses = _load(f"{tmp_path}/log.pkl")
    assert len(losses) == 9 and all(np.isfinite(v[0]) for v in losses.values())


@need(2)
def test_fused_zero_step_kernel_matches_the_nccl_path():
    """Kernel-level equivalence on identical weights / gradients (tests/dist_workers/zero_equiv.py): master slab, both Adam
    moments, clip statistics, the bf16 weights on every rank and the zeroed gradient arena, ZeRO-2 and FULL_SHARD."""
    n = 4 if NGPU >= 4 else 2
    out = torchrun(n, [os.path.join(ROOT, "tests", "dist_workers", "zero_equiv.py")])
    assert "FAIL" not in out and out.count("ALL OK") == n, out[-3000:]


@need(2)
@pytest.mark.parametrize("sharding", ["SHARD_GRAD_OP", "FULL_SHARD"])
def test_training_with_the_fused_zero_step_2gpu(tmp_path, sharding):
    """End to end through the CLI: the one-kernel ZeRO step is picked up (log line) and trains like the NCCL path (GPU
    reductions are not bit-reproducible run to run - wgrad / norm partials accumulate with atomics - hence atol 1e-2)."""
    logs = {}
    for mode in ("1", "0"):
        logs[mode] = f"{tmp_path}/zero{mode}.pkl"
        out = torchrun(2, ["-m", "opendiloco_b200.train_fsdp"] + BASE + ["--max_steps", "8", "--sharding_strategy", sharding,
                                                                         "--project", logs[mode]], extra_env={"ODB_ZERO_FUSED": mode})
        assert ("fused ZeRO step:" in out) == (mode == "1"), out[-2000:]
    a, b = _load(logs["1"]), _load(logs["0"])
    assert set(a) == set(b) == set(range(1, 9))
    for s in a:
        assert np.allclose(a[s][0], b[s][0], atol=1e-2), f"loss at step {s}: fused {a[s][0]} vs nccl {b[s][0]}"
