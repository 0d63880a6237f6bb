"""``kokoro-train = kokoro.cli.training:main`` (reference setup.py:53, src/kokoro/cli/training.py)."""
from kokoro_ruslan_b200.cli import main  # noqa: F401

if __name__ == "__main__":
    import sys
    sys.exit(main())
