"""``from open_diloco.hivemind_diloco import DiLoCoOptimizer, AllReduceStrategy, ...`` (reference module of the same name)."""
from opendiloco_b200.parallel.diloco import (DEFAULT_TIMEOUT_WAITING_FOR_PEERS, AllReduceStrategy, DiLoCoGradAverager,  # noqa: F401
                                             DiLoCoOptimizer, DiLoCoStateAverager, DiloCoProgressTracker)
from opendiloco_b200.parallel.swarm import DHT, StepControl, get_dht_time  # noqa: F401
