"""Model / preprocessor construction from fetched artifacts.

The reference resolves 38 HF auto-classes with ``trust_remote_code=True`` (model.py:47-126,
SURVEY quirk Q12).  Here architectures are LOCAL: ``config.json``'s ``model_type`` selects one of
our model families (llama and mistral -> sm_100a kernels on GPU; gpt2 / lenet -> stock torch), weights come
from ``model.safetensors`` when present and are random-init otherwise (no network).  Unknown
model types fall back to ``transformers`` auto-classes without remote code."""
from __future__ import annotations

import json
import os
from typing import Any, Optional

import torch

AUTO_CLASS = {
    "auto": "AutoModel", "pretraining": "AutoModelForPreTraining", "causal-lm": "AutoModelForCausalLM",
    "masked-lm": "AutoModelForMaskedLM", "mask-generation": "AutoModelForMaskGeneration",
    "seq2-seq-lm": "AutoModelForSeq2SeqLM", "sequence-classification": "AutoModelForSequenceClassification",
    "multiple-choice": "AutoModelForMultipleChoice", "next-sentence-prediction": "AutoModelForNextSentencePrediction",
    "token-classification": "AutoModelForTokenClassification", "question-answering": "AutoModelForQuestionAnswering",
    "text-encoding": "AutoModelForTextEncoding", "depth-estimation": "AutoModelForDepthEstimation",
    "image-classification": "AutoModelForImageClassification", "video-classification": "AutoModelForVideoClassification",
    "keypoint-detection": "AutoModelForKeypointDetection", "keypoint-matching": "AutoModelForKeypointMatching",
    "object-detection": "AutoModelForObjectDetection", "image-segmentation": "AutoModelForImageSegmentation",
    "image-to-image": "AutoModelForImageToImage", "semantic-segmentation": "AutoModelForSemanticSegmentation",
    "instance-segmentation": "AutoModelForInstanceSegmentation", "universal-segmentation": "AutoModelForUniversalSegmentation",
    "zero-shot-image-classification": "AutoModelForZeroShotImageClassification",
    "zero-shot-object-detection": "AutoModelForZeroShotObjectDetection",
    "audio-classification": "AutoModelForAudioClassification", "audio-frame-classification": "AutoModelForAudioFrameClassification",
    "ctc": "AutoModelForCTC", "speech-seq2-seq": "AutoModelForSpeechSeq2Seq", "audio-x-vector": "AutoModelForAudioXVector",
    "text-to-spectrogram": "AutoModelForTextToSpectrogram", "text-to-waveform": "AutoModelForTextToWaveform",
    "audio-tokenization": "AutoModelForAudioTokenization", "table-question-answering": "AutoModelForTableQuestionAnswering",
    "document-question-answering": "AutoModelForDocumentQuestionAnswering", "vison2-seq": "AutoModelForVision2Seq",
    "image-text-to-text": "AutoModelForImageTextToText", "time-series-prediction": "AutoModelForTimeSeriesPrediction",
}


def hf_shift_labels(labels: torch.Tensor) -> tuple:
    """Hugging Face causal-LM convention -> per-position targets.  A dataset hands ``labels`` that are ALIGNED with
    ``input_ids`` (usually a copy of them); the model scores position ``t`` against ``labels[t + 1]`` and the last
    position of every sequence against nothing.  Returns ``(targets, n_valid)``: the shifted targets with ``-100`` in
    the last position (and wherever the dataset masked with ``-100``) and the number of scored positions (counted
    here for host tensors -- what a data loader hands out; ``None`` for device tensors, the loss then counts itself)."""
    shifted = torch.roll(labels, -1, dims=-1)
    shifted[..., -1] = -100
    if shifted.is_cuda:
        return shifted, None                         # counted on the device by the loss (one host synchronisation)
    return shifted, int((shifted != -100).sum())     # host tensor (what the data loader hands out): counting is free


class _LlamaAdapter(torch.nn.Module):
    """Gives ``LlamaForCausalLM`` the HF calling convention (``labels`` == inputs, shift inside)."""

    def __init__(self, net):
        super().__init__()
        self.net = net

    def forward(self, input_ids, labels=None, targets=None, n_valid=None, **_):
        """``labels``: Hugging Face convention (aligned with the inputs, shifted here).  ``targets`` / ``n_valid``: the
        already shifted targets and their count, prepared on the host by the caller (`InnerLoop.step`) so that the
        masked loss needs no device-side count."""
        if targets is not None:
            return {"loss": self.net(input_ids, targets, ignore_index=-100, n_valid=n_valid)}
        if labels is None:
            return {"logits": self.net(input_ids)}
        # HF semantics (what the reference's `model(**batch)` computes): position t is scored against token t+1, the
        # last position has no target.
        shifted, n_valid = hf_shift_labels(labels)
        return {"loss": self.net(input_ids, shifted, ignore_index=-100, n_valid=n_valid)}


def weight_files(model_path: str) -> list:
    """The safetensors files of a model directory: ``model.safetensors``, or the shards named by
    ``model.safetensors.index.json`` (how checkpoints beyond a few GB are published)."""
    single = os.path.join(model_path, "model.safetensors")
    if os.path.exists(single):
        return [single]
    index = os.path.join(model_path, "model.safetensors.index.json")
    if os.path.exists(index):
        names = sorted(set(json.load(open(index)).get("weight_map", {}).values()))
        missing = [n for n in names if not os.path.exists(os.path.join(model_path, n))]
        if missing:
            raise FileNotFoundError(f"{index} names {missing[:3]} which are not in {model_path}: list every shard in "
                                    "the job's model `filenames`")
        return [os.path.join(model_path, n) for n in names]
    return []


def load_weights(model_path: str):
    """State dict of a model directory (all shards merged), or None when it holds no weights."""
    files = weight_files(model_path)
    if not files:
        return None
    from safetensors.torch import load_file

    sd: dict = {}
    for f in files:
        sd.update(load_file(f))
    return sd


def get_model(model_path: str, model_type: str, device: str = "cpu") -> torch.nn.Module:
    cfg_path = os.path.join(model_path, "config.json")
    if not os.path.exists(cfg_path):
        raise FileNotFoundError(f"{cfg_path} missing: fetch the model artifact first")
    meta = json.load(open(cfg_path))
    family = meta.get("model_type", "")
    weights = os.path.join(model_path, "model.safetensors")
    # without a checkpoint every worker must still start from the SAME weights (the reference gets this
    # from the shared HF download): random init is seeded from the config, not from the process
    rng_state = torch.random.get_rng_state()
    torch.manual_seed(int(meta.get("seed", 0)))
    try:
        return _build_model(model_path, model_type, device, meta, family, weights, cfg_path)
    finally:
        torch.random.set_rng_state(rng_state)


def _build_model(model_path: str, model_type: str, device: str, meta: dict, family: str, weights: str,
                 cfg_path: str) -> torch.nn.Module:
    if family == "gpt2":
        from ..models.gpt2 import GPT2Config, GPT2LMHeadModel

        net = GPT2LMHeadModel(GPT2Config.from_json(cfg_path))
    elif family == "lenet":
        from ..models.lenet import LeNetConfig, LeNetForImageClassification

        net = LeNetForImageClassification(LeNetConfig.from_json(cfg_path))
    elif family in ("llama", "mistral", "qwen2", "qwen3", "phi3"):      # models.llama.SM100_FAMILIES
        from ..models.llama import LlamaConfig, LlamaForCausalLM

        c = LlamaConfig.from_hf_dict(meta)          # llama / mistral -> the sm_100a kernels (see its docstring)
        dtype = torch.bfloat16 if device.startswith("cuda") else torch.float32
        inner = LlamaForCausalLM(c, device=device, dtype=dtype)
        sd = load_weights(model_path)
        if sd is not None:
            inner.load_hf_state_dict(sd)
        return _LlamaAdapter(inner)
    else:
        import transformers

        cls_name = AUTO_CLASS.get(model_type)
        if cls_name is None:
            raise RuntimeError(f"model type {model_type!r} not supported")
        return getattr(transformers, cls_name).from_pretrained(model_path)
    sd = load_weights(model_path)
    if sd is not None:
        net.load_state_dict(sd, strict=False)
    return net.to(device)


def get_preprocessor(cfg: Optional[dict], model_path: str) -> Optional[Any]:
    """tokenizer | feature | image | video | auto -> transformers processors (utils.py:37-53);
    the built-in ``image`` fallback just renames ``images`` -> ``pixel_values``."""
    if not cfg:
        return None
    kind = cfg["task"]
    pp_path = os.path.join(model_path, "preprocessor_config.json")
    try:
        import transformers

        cls = {"auto": "AutoProcessor", "feature": "AutoFeatureExtractor", "image": "AutoImageProcessor",
               "tokenizer": "AutoTokenizer", "video": "AutoVideoProcessor"}[kind]
        if os.path.exists(pp_path) and json.load(open(pp_path)).get("image_processor_type") != "hypha-identity":
            return getattr(transformers, cls).from_pretrained(model_path)
    except KeyError:
        raise RuntimeError(f"Pre-Processor of type {kind} not found")
    except Exception as exc:      # a broken processor config: say so, then use the built-in fallback if there is one
        print(f"preprocessor of type {kind} could not be loaded from {model_path}: {type(exc).__name__}: {exc}", flush=True)
    if kind == "image":
        return lambda images: {"pixel_values": images}
    return None
