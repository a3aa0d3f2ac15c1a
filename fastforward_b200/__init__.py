"""fastforward_b200 -- B200-native backend for FastForward's quantization hot path."""
from . import ops  # noqa: F401
