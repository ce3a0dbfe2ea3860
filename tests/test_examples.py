"""The programs under `examples/` stay runnable: the fabric example on every link type, and the two others at least
parse (their full runs need a worker / a GPU)."""
import asyncio
import importlib.util
import os
import subprocess
import sys

import pytest

EX = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "examples")


def _load(name):
    spec = importlib.util.spec_from_file_location(f"example_{name}", os.path.join(EX, f"{name}.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


@pytest.mark.parametrize("scheme", ["memory", "tcp", "udp"])
def test_fabric_example(scheme):
    out = asyncio.run(asyncio.wait_for(_load("fabric_request_response").main(scheme), 60))
    assert out["loud"] == {"echo": "HELLO", "by": "shout"} and out["quiet"] == {"echo": "hello", "by": "plain"}
    assert out["gossip"] == ["client", "an advertisement"] and out["record"] == ['{"num_slices": 3}', "server"]
    assert out["pushed_ok"] and out["pulled"] == "slice-2-of-example" and out["bytes_out_client"] > 1 << 20


@pytest.mark.parametrize("name", ["custom_executor", "train_python_api"])
def test_other_examples_parse_their_arguments(name):
    r = subprocess.run([sys.executable, os.path.join(EX, f"{name}.py"), "--help"], capture_output=True, text=True, timeout=120)
    assert r.returncode == 0 and "usage:" in r.stdout
