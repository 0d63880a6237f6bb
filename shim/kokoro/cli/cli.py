"""Argument parser and config construction (reference src/kokoro/cli/cli.py:33-290)."""
from kokoro_ruslan_b200.cli import build_parser, create_config_from_args  # noqa: F401
