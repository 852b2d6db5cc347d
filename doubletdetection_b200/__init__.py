"""doubletdetection_b200 -- B200-native ``BoostClassifier.fit`` hot path of DoubletDetection.

``from doubletdetection_b200 import BoostClassifier`` is the drop-in for
``from doubletdetection import BoostClassifier`` (reference: doubletdetection/__init__.py:1,14).
"""

from .classifier import BoostClassifier, iteration_shard  # noqa: F401

__version__ = "0.1.0"
__all__ = ["BoostClassifier", "iteration_shard"]
