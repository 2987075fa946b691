"""``python open_diloco/init_weights.py --config-name-or-path 150m --save-to-disk ./llama-150m-fresh``"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from opendiloco_b200.init_weights import main  # noqa: E402

if __name__ == "__main__":
    main()
