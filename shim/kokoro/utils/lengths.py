"""Length regulation utilities (reference src/kokoro/utils/lengths.py:16-208)."""
from kokoro_ruslan_b200.lengths import (LengthRegulator, average_by_duration, length_regulate,  # noqa: F401
                                        vectorized_expand_tokens)
