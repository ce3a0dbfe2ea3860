"""hypha-data: the data node (reference: crates/data/src/bin/hypha-data.rs:153-209, tensor_data.rs:8-16).

Announces ``kad::Record{key = <dataset directory name>, value = JSON DataRecord{num_slices}}`` plus a provider entry,
then serves `/hypha-tensor-stream/pull`: header ``DataSlice{dataset, index}`` -> the bytes of slice ``index``.
Differences to the reference, on purpose: slices are indexed in SORTED file order (the reference indexes
``read_dir()`` order, which is unspecified: quirk Q5) and the bounds check is ``0 <= index < n`` (the reference
tests ``>``).  The announcement is repeated when a gateway comes back (a restarted gateway has an empty DHT) and by the
DHT upkeep timer of every role.
"""
from __future__ import annotations

import asyncio
import json
import os
from typing import Any, List, Optional

from .. import config as C
from .. import messages as M


def dataset_files(path: str) -> List[str]:
    """Slices of a dataset directory in SORTED order (the reference indexes ``read_dir()`` order,
    which is unspecified: quirk Q5)."""
    return sorted(os.path.join(path, f) for f in os.listdir(path) if os.path.isfile(os.path.join(path, f)))


async def run_data(cfg: C.DataConfig, *, peer_id: Optional[str] = None,
                   started: Optional[asyncio.Future] = None) -> None:
    from ..roles import RoleError, start_node

    C.validate(cfg)
    node = await start_node("data", cfg, peer_id=peer_id)
    name = os.path.basename(os.path.normpath(cfg.dataset_path))
    files = dataset_files(cfg.dataset_path)
    if not files:
        raise RoleError(f"dataset directory {cfg.dataset_path} is empty")

    async def serve(peer: str, header: Any):
        sl = M.DataSlice.from_wire(header)
        if sl.dataset != name:
            raise RoleError(f"unknown dataset {sl.dataset!r}")
        if not (0 <= sl.index < len(files)):            # correct bounds (reference uses `>`: Q5)
            raise RoleError(f"slice {sl.index} out of range 0..{len(files) - 1}")
        return files[sl.index]

    node.network.serve_pull(serve)

    async def announce(*_):
        await node.network.put_record(name, json.dumps(M.DataRecord(len(files)).to_wire()).encode())
        await node.network.provide(name)

    await announce()
    node.on_gateway_reconnect.append(announce)       # a restarted gateway has an empty DHT
    node.ready.set()
    node.telemetry.logger.info("data node %s announces %s (%d slices)", node.peer_id, name, len(files))
    if started is not None and not started.done():
        started.set_result(node)
    try:
        await node.stop.wait()
    finally:
        await node.shutdown()
