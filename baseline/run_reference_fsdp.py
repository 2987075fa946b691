"""Second reference arm of bench.py (``--impl reference-fsdp``): the UNMODIFIED ``baseline/_ref/open_diloco/train_fsdp.py`` -
the reference's flagship script - on its stock non-hivemind path: FSDP(NO_SHARD, MixedPrecision(bf16)) + ``torch.compile``
(``Config.torch_compile`` defaults to True, train_fsdp.py:106,246-247) + ``--fake-data``.  This is the STRONGER baseline
on this box (Inductor-fused norms / RoPE / SwiGLU / CE); the eager ``train_diloco_torch.py`` arm stays the DiLoCo
numerics oracle.  The hivemind (DiLoCo) half of train_fsdp.py cannot run here at all: ``hivemind`` is not installed.

Nothing in ``baseline/_ref`` is edited.  Provided from OUTSIDE the reference, only so that its imports resolve:
  * ``hivemind``: import-time placeholders for the names ``open_diloco.hivemind_diloco`` / ``train_fsdp`` import
    (base classes that are subclassed at import, a ``logger``); none of them is ever instantiated on the non-hv path;
  * ``pydantic_config.parse_argv``: the installed pydantic_config (0.3.0) has ``BaseConfig`` but not ``parse_argv``
    (used only under ``__main__``); the arm builds ``Config(...)`` directly and calls ``train(config)``;
  * ``AutoTokenizer.from_pretrained("mistralai/Mistral-7B-v0.1")`` needs the hub even with --fake-data
    (train_fsdp.py:218-221): a locally built tokenizer with the same pad token is returned instead;
  * ``FakeTokenizedDataset`` draws from vocab 1024 in the reference (TEST_VOCAB_SIZE); it is replaced by the same class
    drawing from the model's vocabulary so the token law equals the other arms' (uniform over [3, vocab));
  * the model directory is a random-init Llama-150M written with ``save_pretrained``.
Timing: rank 0's metric logger (the reference's DummyLogger) is called once per optimizer step right after the
``loss_batch.item()`` host sync; the timed window is the wall clock between the calls that close step W and step W+K,
bracketed by ``torch.cuda.synchronize()``.  Data-parallel training synchronises all ranks every step (gradient all-reduce),
so rank 0's window is the job's window.
"""
from __future__ import annotations

import json
import logging
import os
import sys
import tempfile
import time
import types

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)


def install_hivemind_placeholders() -> None:
    if "hivemind" in sys.modules:
        return
    layout = {
        "hivemind": ["NoCompression", "Float16Compression", "ScaledFloat16Compression", "Uniform8BitQuantization",
                     "Quantile8BitQuantization", "BlockwiseQuantization"],
        "hivemind.averaging": [], "hivemind.averaging.averager": ["DecentralizedAverager"],
        "hivemind.averaging.control": ["StepControl"],
        "hivemind.compression": [], "hivemind.compression.base": ["CompressionBase", "NoCompression"],
        "hivemind.dht": [], "hivemind.dht.dht": ["DHT"],
        "hivemind.optim": [], "hivemind.optim.optimizer": ["Optimizer"],
        "hivemind.optim.progress_tracker": ["GlobalTrainingProgress", "ProgressTracker", "TrainingProgressSchema",
                                            "LocalTrainingProgress"],
        "hivemind.optim.state_averager": ["LRSchedulerBase", "OptimizerFactory", "Parameters", "ParamGroups", "SchedulerFactory",
                                          "TorchOptimizer", "TrainingStateAverager"],
        "hivemind.utils": ["get_dht_time"], "hivemind.utils.timed_storage": ["DHTExpiration"],
        "hivemind.utils.networking": ["log_visible_maddrs"],
    }
    for name, attrs in layout.items():
        m = types.ModuleType(name)
        m.__path__ = []            # a package, so that sub-module imports resolve through sys.modules
        for a in attrs:
            # inert: ``NoCompression()`` is evaluated as a default argument when hivemind_diloco.py is imported
            setattr(m, a, type(a, (), {"__init__": lambda self, *x, **k: None, "__doc__": "placeholder: hivemind is not installed"}))
        sys.modules[name] = m
    for name in layout:
        if "." in name:
            parent, child = name.rsplit(".", 1)
            setattr(sys.modules[parent], child, sys.modules[name])
    log = logging.getLogger("hivemind-placeholder")
    if not log.handlers:
        log.addHandler(logging.StreamHandler(sys.stderr))
    log.setLevel(logging.WARNING)
    sys.modules["hivemind.optim.optimizer"].logger = log
    sys.modules["hivemind.utils"].get_dht_time = time.time
    sys.modules["hivemind.utils.networking"].log_visible_maddrs = lambda *a, **k: None


def install_shims() -> None:
    os.environ.setdefault("WANDB_MODE", "disabled")
    os.environ.setdefault("WANDB_SILENT", "true")
    os.environ.setdefault("TOKENIZERS_PARALLELISM", "false")
    os.environ.setdefault("HF_HUB_OFFLINE", "1")
    install_hivemind_placeholders()
    import pydantic_config

    if not hasattr(pydantic_config, "parse_argv"):
        pydantic_config.parse_argv = lambda *a, **k: {}
    ref_dir = os.path.join(HERE, "_ref")
    if ref_dir not in sys.path:
        sys.path.insert(0, ref_dir)


class TimedLogger:
    """Drop-in for the reference's DummyLogger (utils.py:189-204) that also timestamps every optimizer step."""

    marks: dict = {}
    losses: list = []
    warmup = steps = 0
    sampler = None

    def __init__(self, project, config, *args, **kwargs):
        self.project, self.config = project, config

    def log(self, metrics: dict):
        import torch

        step = int(metrics["step"])
        TimedLogger.losses.append((step, float(metrics["Loss"])))
        if step == TimedLogger.warmup:
            torch.cuda.synchronize()
            if TimedLogger.sampler is not None:
                TimedLogger.sampler.start()
            TimedLogger.marks["t0"] = time.perf_counter()
        elif step == TimedLogger.warmup + TimedLogger.steps:
            torch.cuda.synchronize()
            TimedLogger.marks["t1"] = time.perf_counter()
            if TimedLogger.sampler is not None:
                TimedLogger.marks["clocks"] = TimedLogger.sampler.stop()

    def finish(self):
        pass


def make_model_dir(model: str, rank: int) -> tuple[str, int]:
    import torch

    cfg_json = json.load(open(os.path.join(ROOT, "opendiloco_b200", "configs", f"config_{model}.json")))
    model_dir = os.path.join(tempfile.gettempdir(), f"odb_ref_llama_{model}")
    if rank == 0 and not os.path.exists(os.path.join(model_dir, "config.json")):
        from transformers import LlamaConfig, LlamaForCausalLM

        torch.manual_seed(0)
        hf_cfg = LlamaConfig(**{k: v for k, v in cfg_json.items() if k not in ("architectures", "model_type")})
        LlamaForCausalLM(hf_cfg).save_pretrained(model_dir)
    return model_dir, cfg_json.get("vocab_size", 32000)


def run(a, ClockSampler) -> dict:
    import torch
    import torch.distributed as dist

    install_shims()
    from run_reference import _local_tokenizer

    import open_diloco.train_fsdp as ref

    world = int(os.environ.get("WORLD_SIZE", 1))
    rank = int(os.environ.get("RANK", 0))
    local_rank = int(os.environ.get("LOCAL_RANK", 0))
    for k, v in (("LOCAL_RANK", "0"), ("RANK", "0"), ("WORLD_SIZE", "1"), ("LOCAL_WORLD_SIZE", str(world)),
                 ("MASTER_ADDR", "127.0.0.1"), ("MASTER_PORT", "29534")):
        os.environ.setdefault(k, v)
    assert world == a.gpus, f"--gpus {a.gpus} but WORLD_SIZE={world}"
    torch._dynamo.config.suppress_errors = "PRIME_INTELLECT_DEV" not in os.environ      # as the reference's __main__ does
    torch.set_float32_matmul_precision("high")
    ref.ddp_setup()
    model_dir, vocab = make_model_dir(a.model, rank)
    dist.barrier(device_ids=[local_rank])

    class _Tok:
        @staticmethod
        def from_pretrained(*args, **kwargs):
            return _local_tokenizer(vocab)

    ref.AutoTokenizer = _Tok
    stock_fake = ref.FakeTokenizedDataset
    ref.FakeTokenizedDataset = lambda seq_len, _vocab: stock_fake(seq_len, vocab)     # the reference class, model vocabulary
    TimedLogger.marks, TimedLogger.losses = {}, []
    TimedLogger.warmup, TimedLogger.steps = a.warmup, a.steps
    TimedLogger.sampler = ClockSampler(local_rank) if rank == 0 else None
    ref.DummyLogger = TimedLogger
    accum = a.batch // a.micro_batch
    config = ref.Config(path_model=model_dir, torch_compile=True, seq_length=a.seq, lr=4e-4, total_batch_size=a.batch * world,
                        per_device_train_batch_size=a.micro_batch, warmup_steps=1000, total_steps=88_000,
                        sharding_strategy="NO_SHARD", precision="bf16-mixed", project=os.path.join(tempfile.gettempdir(), "odb_ref_fsdp_log.pkl"),
                        metric_logger_type="dummy", fake_data=True, max_steps=a.warmup + a.steps,
                        ckpt=ref.CkptConfig(path=os.path.join(tempfile.gettempdir(), "odb_ref_fsdp_ckpt")))
    ref.train(config)
    out = {}
    if rank == 0:
        secs = TimedLogger.marks["t1"] - TimedLogger.marks["t0"]
        tokens_per_step = a.batch * a.seq * world
        value = tokens_per_step * a.steps / secs
        out = {
            "metric": "tokens_per_sec (data-parallel inner steps, reference train_fsdp.py non-hv path, torch.compile)",
            "value": value, "unit": "tokens/s", "n_gpus": world, "steps": a.steps, "warmup": a.warmup,
            "ms_per_step": secs * 1e3 / a.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "bf16", "data": "synthetic tokens (uniform over vocab), random-init weights", "impl": "reference-fsdp",
            "config": {"model": f"llama-{a.model}", "global_batch": a.batch * world, "per_worker_batch": a.batch,
                       "micro_batch": a.micro_batch, "grad_accum": accum, "seq_len": a.seq, "parallelism": f"dp{world}",
                       "local_steps": None, "outer_steps_in_timed_window": 0},
            "entry": "open_diloco.train_fsdp.train (unmodified, baseline/_ref): FSDP NO_SHARD + MixedPrecision(bf16) + torch.compile",
            "tokens_per_sec_per_gpu": value / world,
            "clocks": TimedLogger.marks.get("clocks"),
            "e2e": {"value": value, "unit": "tokens/s", "h2d_bytes_per_step": a.micro_batch * a.seq * 8 * 3 * accum,
                    "d2h_bytes_per_step": 4, "note": "the reference loop is inherently end-to-end (DataLoader workers, pageable H2D "
                                                     "every micro-batch, loss .item() every step); wall-clock on rank 0"},
            "last_loss": TimedLogger.losses[-1][1] if TimedLogger.losses else None,
            "gpu_launches": None,
        }
    if dist.is_initialized():
        dist.destroy_process_group()
    return out
