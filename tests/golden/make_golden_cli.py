"""Generates tests/golden/cli_flags.json from the LIVE reference parser (run in the build container only):
    python tests/golden/make_golden_cli.py
The reference builds its parser inside parse_arguments() and parses immediately, so parse_args is intercepted."""
import argparse
import json
import os
import sys

sys.path.insert(0, "/root/reference/src")
captured = []
argparse.ArgumentParser.parse_args = lambda self, *a, **k: (captured.append(self), argparse.Namespace())[1]
from kokoro.cli.cli import parse_arguments  # noqa: E402

parse_arguments()
parser = captured[0]
rows = []
for a in parser._actions:
    if isinstance(a, argparse._HelpAction):
        continue
    rows.append({"options": sorted(a.option_strings), "dest": a.dest, "default": a.default,
                 "type": getattr(a.type, "__name__", None), "action": type(a).__name__})
out = os.path.join(os.path.dirname(os.path.abspath(__file__)), "cli_flags.json")
json.dump(sorted(rows, key=lambda r: r["dest"] + "".join(r["options"])), open(out, "w"), indent=1)
print("wrote", out, len(rows), "options")
