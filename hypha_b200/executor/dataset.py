"""Streaming dataset over safetensors slices (reference: .../accelerate_executor/dataset.py:9-41)
plus a synthetic-slice writer for offline runs."""
from __future__ import annotations

import os
from typing import Any, Callable, Dict, Iterator, List, Optional

import torch
from torch.utils.data import IterableDataset


class IterableStreamDataSet(IterableDataset):
    """Each file yielded by ``files`` is one slice: a safetensors dict of equally long tensors;
    samples are emitted row by row as ``{input_name: tensor}``."""

    def __init__(self, files: Iterator[str], model_inputs: List[str], processor_inputs: Optional[List[str]] = None,
                 preprocessor: Optional[Callable[..., Dict[str, Any]]] = None):
        super().__init__()
        self.files, self.model_inputs = files, model_inputs
        self.processor_inputs, self.preprocessor = processor_inputs or [], preprocessor

    def __iter__(self):
        from safetensors.torch import load_file

        for path in self.files:
            data = load_file(path, device="cpu")
            if self.preprocessor is not None:
                extra = self.preprocessor(**{k: data.pop(k) for k in self.processor_inputs})
                data = {**extra, **data}
            for values in zip(*(data[k] for k in self.model_inputs)):
                yield dict(zip(self.model_inputs, values))


def forever(loader) -> Iterator[Dict[str, torch.Tensor]]:
    while True:
        yield from loader


def write_synthetic_token_slices(path: str, num_slices: int, samples_per_slice: int, seq_len: int, vocab_size: int,
                                 seed: int = 0, with_labels: bool = True) -> List[str]:
    """Create ``num_slices`` safetensors files of random token ids (the offline stand-in for a
    pre-tokenised corpus; one file = one slice, docs/training.md:123-128)."""
    from safetensors.torch import save_file

    os.makedirs(path, exist_ok=True)
    g = torch.Generator().manual_seed(seed)
    out = []
    for i in range(num_slices):
        ids = torch.randint(0, vocab_size, (samples_per_slice, seq_len), generator=g, dtype=torch.int64)
        d = {"input_ids": ids}
        if with_labels:
            d["labels"] = ids.clone()
        f = os.path.join(path, f"slice-{i:05d}.safetensors")
        save_file(d, f)
        out.append(f)
    return out


def write_synthetic_image_slices(path: str, num_slices: int, samples_per_slice: int, image_size: int = 28,
                                 num_classes: int = 10, seed: int = 0, pixel_key: str = "pixel_values") -> List[str]:
    """MNIST-shaped stand-in: one fixed random prototype per class + noise, so a classifier can learn it.
    ``pixel_key="images"`` writes the raw key an image preprocessor consumes."""
    from safetensors.torch import save_file

    os.makedirs(path, exist_ok=True)
    g = torch.Generator().manual_seed(seed)
    base = torch.randn(num_classes, image_size, image_size, generator=g)       # the same classes in every slice
    out = []
    for i in range(num_slices):
        labels = torch.randint(0, num_classes, (samples_per_slice,), generator=g)
        px = base[labels] + 0.1 * torch.randn(samples_per_slice, image_size, image_size, generator=g)
        f = os.path.join(path, f"slice-{i:05d}.safetensors")
        save_file({pixel_key: px.contiguous(), "labels": labels}, f)
        out.append(f)
    return out
