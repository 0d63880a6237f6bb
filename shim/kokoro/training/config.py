"""``kokoro.training.config.TrainingConfig`` (reference src/kokoro/training/config.py): the class checkpoints pickle as
their ``config`` entry (trainer.py:1994-2031), so this import path has to exist wherever such a checkpoint is loaded."""
from kokoro_ruslan_b200.cli import TrainingConfig  # noqa: F401
