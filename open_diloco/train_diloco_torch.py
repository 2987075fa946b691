"""``torchrun open_diloco/train_diloco_torch.py ...`` - same CLI as the reference script."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from opendiloco_b200.train_diloco_torch import TorchDilocoConfig, cli, main  # noqa: E402,F401

if __name__ == "__main__":
    cli()
