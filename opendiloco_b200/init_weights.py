"""Create a randomly initialised checkpoint from a config (reference: open_diloco/init_weights.py:10-25).

    python -m opendiloco_b200.init_weights --config-name-or-path 150m --save-to-disk ./llama-150m-fresh [--seed 0]

Writes ``config.json`` + ``model.safetensors`` in the HF layout (names of SURVEY.md Appendix B), loadable by both this
framework and ``transformers``.  ``--hub-model-id`` is accepted for CLI parity; pushing needs network access."""
from __future__ import annotations

from .models.config import LlamaConfig
from .models.llama import LlamaForCausalLM
from .utils.config import BaseConfig, parse_argv


class InitConfig(BaseConfig):
    config_name_or_path: str
    hub_model_id: str | None = None
    save_to_disk: str | None = None
    seed: int = 0


def main(argv: list[str] | None = None) -> None:
    cfg = InitConfig(**parse_argv(argv))
    model = LlamaForCausalLM(LlamaConfig.from_pretrained(cfg.config_name_or_path), device="cpu", precision="32-true", seed=cfg.seed)
    print(f"{cfg.config_name_or_path}: {model.num_parameters():,} parameters")
    if cfg.save_to_disk:
        model.save_pretrained(cfg.save_to_disk)
        print(f"saved to {cfg.save_to_disk}")
    if cfg.hub_model_id:
        raise SystemExit("--hub-model-id: pushing to the hub needs network access; use --save-to-disk and upload the directory")


if __name__ == "__main__":
    main()
