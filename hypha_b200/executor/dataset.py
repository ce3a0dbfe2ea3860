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


# ---------------------------------------------------------------------------------------------------------------------
# dataset preparation: a token corpus -> a directory of slices a data node can serve (`hypha-data prepare`)
# ---------------------------------------------------------------------------------------------------------------------
def load_token_corpus(src: str, dtype: str = "uint16", tokenizer: Optional[str] = None, append_eos: bool = True):
    """A corpus as ONE flat array of token ids or, for already packed data, a ``[samples, seq_len]`` array (a numpy
    array -- memory-mapped for ``.npy`` / ``.bin``, so a corpus larger than RAM can be sliced -- or a tensor).

    * ``.npy`` -- a numpy array of integer ids, 1-D (a token stream) or 2-D (packed samples);
    * ``.bin`` -- raw little-endian ids of ``dtype`` (``uint16`` | ``uint32`` | ``int32`` | ``int64``), the layout
      of the common pre-tokenised corpora; read through a memory map;
    * ``.safetensors`` -- its ``input_ids`` tensor;
    * ``.txt`` / ``.jsonl`` -- text (one document per line; jsonl: the ``text`` field), tokenised with the Hugging
      Face tokenizer stored in the LOCAL directory ``tokenizer`` (no download), documents joined by the EOS token.
    """
    import numpy as np

    ext = os.path.splitext(src)[1].lower()
    if ext == ".npy":
        arr = np.load(src, mmap_mode="r")
    elif ext == ".bin":
        if dtype not in ("uint16", "uint32", "int32", "int64"):
            raise ValueError(f"unsupported token dtype {dtype!r}")
        arr = np.memmap(src, dtype=np.dtype(dtype).newbyteorder("<"), mode="r")
    elif ext == ".safetensors":
        from safetensors.torch import load_file

        return load_file(src)["input_ids"].to(torch.int64)
    elif ext in (".txt", ".jsonl"):
        if not tokenizer:
            raise ValueError("text input needs --tokenizer <local directory of a Hugging Face tokenizer>")
        import json

        import transformers

        tok = transformers.AutoTokenizer.from_pretrained(tokenizer, local_files_only=True)
        eos = tok.eos_token_id if append_eos else None
        ids: List[int] = []
        with open(src, encoding="utf-8") as f:
            for line in f:
                line = line.rstrip("\n")
                if ext == ".jsonl":
                    if not line.strip():
                        continue
                    line = json.loads(line).get("text", "")
                if not line:
                    continue
                ids.extend(tok(line, add_special_tokens=False)["input_ids"])
                if eos is not None:
                    ids.append(eos)
        return torch.tensor(ids, dtype=torch.int64)
    else:
        raise ValueError(f"unsupported corpus format {ext!r} (.npy | .bin | .safetensors | .txt | .jsonl)")
    if arr.dtype.kind not in "iu":
        raise ValueError(f"{src}: token ids must be integers, not {arr.dtype}")
    if arr.ndim not in (1, 2):
        raise ValueError(f"{src}: expected a 1-D token stream or 2-D packed samples, got shape {arr.shape}")
    return arr                                   # possibly a memory map: slices are converted as they are written


def write_token_slices(tokens, out_dir: str, seq_len: int, samples_per_slice: int,
                       with_labels: bool = True, vocab_size: Optional[int] = None, shuffle_seed: Optional[int] = None,
                       drop_last: bool = False) -> Dict[str, Any]:
    """Pack ``tokens`` into samples of ``seq_len`` ids (a 1-D stream is cut without overlap, the tail shorter than
    one sample is dropped; a 2-D tensor must already be ``[samples, seq_len]``) and write them as
    ``slice-00000.safetensors`` ... with ``samples_per_slice`` rows each -- the unit a worker fetches and the scheduler
    hands out (data_scheduler.rs:28-103).  ``labels`` = ``input_ids`` (the Hugging Face convention: the model shifts).
    ``shuffle_seed`` permutes the samples before slicing (a corpus that is a long run of one source after another
    ends up spread over the workers).  Returns a summary, also written to ``<out_dir>.json`` -- next to the slice
    directory, not inside it: a data node serves every file of its directory as a slice."""
    from safetensors.torch import save_file

    import numpy as np

    if seq_len < 2 or samples_per_slice < 1:
        raise ValueError("seq_len must be >= 2 and samples_per_slice >= 1")
    if isinstance(tokens, torch.Tensor):
        tokens = tokens.numpy()
    if tokens.ndim == 1:
        n = tokens.shape[0] // seq_len
        samples = tokens[: n * seq_len].reshape(n, seq_len)
    elif tokens.ndim == 2:
        if tokens.shape[1] != seq_len:
            raise ValueError(f"packed corpus has sequence length {tokens.shape[1]}, asked for {seq_len}")
        samples = tokens
    else:
        raise ValueError("tokens must be 1-D or 2-D")
    if samples.shape[0] == 0:
        raise ValueError(f"corpus of {tokens.size} tokens is shorter than one sample of {seq_len}")
    lo, hi = None, None
    for begin in range(0, samples.shape[0], 1 << 16):               # range check in pieces (the corpus may be mapped)
        part = samples[begin:begin + (1 << 16)]
        lo = int(part.min()) if lo is None else min(lo, int(part.min()))
        hi = int(part.max()) if hi is None else max(hi, int(part.max()))
    if lo < 0 or (vocab_size is not None and hi >= vocab_size):
        raise ValueError(f"token ids span {lo}..{hi}, outside the vocabulary of {vocab_size}")
    order = None
    if shuffle_seed is not None:
        order = torch.randperm(samples.shape[0], generator=torch.Generator().manual_seed(shuffle_seed)).numpy()
    os.makedirs(out_dir, exist_ok=True)
    if os.listdir(out_dir):
        raise ValueError(f"{out_dir} is not empty: a data node serves every file of its directory as a slice")
    files, written = [], 0
    for i, begin in enumerate(range(0, samples.shape[0], samples_per_slice)):
        rows = samples[begin:begin + samples_per_slice] if order is None \
            else samples[np.sort(order[begin:begin + samples_per_slice])]
        if rows.shape[0] < samples_per_slice and drop_last and files:
            break
        ids = torch.from_numpy(np.array(rows, dtype=np.int64))       # a copy: the source may be a read-only map
        d = {"input_ids": ids}
        if with_labels:
            d["labels"] = ids.clone()
        f = os.path.join(out_dir, f"slice-{i:05d}.safetensors")
        save_file(d, f)
        files.append(f)
        written += int(ids.shape[0])
    summary = {"dataset": os.path.basename(os.path.normpath(out_dir)), "num_slices": len(files), "samples": written,
               "seq_len": seq_len, "samples_per_slice": samples_per_slice, "tokens": written * seq_len,
               "max_token_id": hi, "inputs": ["input_ids"] + (["labels"] if with_labels else [])}
    import json

    with open(os.path.normpath(out_dir) + ".json", "w") as f:
        json.dump(summary, f, indent=1)
    return summary
