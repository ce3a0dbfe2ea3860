#!/usr/bin/env python
"""Regenerate docs/cli.md from the argparse parsers (the reference renders its CLI reference at build time,
crates/*/build.rs:10-28).  `--check` exits 1 when the committed file is stale (tests/test_cli_misc.py)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
os.environ["COLUMNS"] = "100"                    # argparse wraps to the terminal width: pin it


def _parser_markdown(prog: str, parser) -> str:
    out = [f"# {prog}", "", "```", parser.format_help().rstrip(), "```", ""]
    subs = parser._subparsers._group_actions[0].choices if parser._subparsers else {}
    for name, sp in subs.items():
        out += [f"## {prog} {name}", "", "```", sp.format_help().rstrip(), "```", ""]
    return "\n".join(out)


def render() -> str:
    from hypha_b200 import cli
    from hypha_b200.cli import box, certutil

    parts = ["# CLI reference (generated from the argparse parsers: `python tools/gen_cli_docs.py`)", ""]
    for role in ("gateway", "scheduler", "worker", "data"):
        parts.append(cli.cli_markdown(role))
    parts.append(_parser_markdown("hypha-certutil", certutil.build_parser()))
    parts.append(_parser_markdown("hypha-box", box.build_parser()))
    return "\n".join(parts).rstrip() + "\n"


def main() -> int:
    path = os.path.join(ROOT, "docs", "cli.md")
    text = render()
    if "--check" in sys.argv[1:]:
        if open(path).read() != text:
            print("docs/cli.md is stale: run python tools/gen_cli_docs.py", file=sys.stderr)
            return 1
        return 0
    with open(path, "w") as f:
        f.write(text)
    print(f"wrote {path} ({len(text.splitlines())} lines)")
    return 0


if __name__ == "__main__":
    sys.exit(main())
