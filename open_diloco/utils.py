"""Reference ``open_diloco.utils`` surface."""
from opendiloco_b200.parallel.compression import get_compression_kwargs  # noqa: F401
from opendiloco_b200.utils.data import FakeTokenizedDataset  # noqa: F401
from opendiloco_b200.utils.logger import DummyLogger, Logger, WandbLogger  # noqa: F401
from opendiloco_b200.utils.metrics import get_grad_norm, log_activations_hook, register_metrics_hooks  # noqa: F401
from opendiloco_b200.utils.training import found_inf_grad, hash_tensor_content  # noqa: F401
from opendiloco_b200.parallel.sharding import get_sharding_strategy  # noqa: F401

register_hooks_log_activations = register_metrics_hooks
