# Developer entry points (everything is also reachable as plain python commands; see README.md).
PY ?= python

.PHONY: build test test-gpu bench sanitize docs sass clean

build:            ## compile _core (C++) and _C (sm_100a CUDA) in-tree
	$(PY) -m hypha_b200.build

test: build       ## CPU suite: control plane, C++ core vectors, loopback + multi-process jobs
	$(PY) -m pytest tests -q -m "not gpu"

test-gpu: build   ## kernel numerics, trainer, multi-GPU outer step (needs B200s)
	$(PY) -m pytest tests -q -m gpu

bench: build      ## Llama-3-8B DiLoCo tokens/s on one GPU (one JSON line)
	$(PY) bench.py --gpus 1 --steps 8 --warmup 3

sanitize:         ## ASAN/UBSAN over the C++ core, TSAN over the barrier model and the PS reducer
	sh tools/sanitize.sh core
	sh tools/sanitize.sh tsan

docs:             ## regenerate the CLI reference from the argparse parsers
	$(PY) tools/gen_cli_docs.py

sass: build       ## per-kernel SASS listings under profiles/sass
	$(PY) tools/dump_sass.py

clean:
	$(PY) -m hypha_b200.build clean
