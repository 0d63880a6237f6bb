"""``kokoro.model.model.KokoroModel`` (reference src/kokoro/model/model.py:35-845)."""
from kokoro_ruslan_b200.model import KokoroModel  # noqa: F401
