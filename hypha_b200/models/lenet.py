"""LeNet-5 image classifier: the reference's default job (``hypha-space/lenet`` on MNIST,
crates/scheduler/src/scheduler_config.rs:52-78)."""
from __future__ import annotations

import json
from dataclasses import asdict, dataclass
from typing import Optional

import torch
import torch.nn as nn
import torch.nn.functional as F


@dataclass
class LeNetConfig:
    num_classes: int = 10
    in_channels: int = 1
    image_size: int = 28
    model_type: str = "lenet"

    @classmethod
    def from_json(cls, path: str) -> "LeNetConfig":
        d = json.load(open(path))
        return cls(**{k: v for k, v in d.items() if k in cls.__dataclass_fields__})

    def to_json(self, path: str) -> None:
        d = asdict(self)
        d["architectures"] = ["LeNetForImageClassification"]
        json.dump(d, open(path, "w"), indent=1)


class LeNetForImageClassification(nn.Module):
    def __init__(self, config: LeNetConfig = LeNetConfig()):
        super().__init__()
        self.config = config
        self.conv1 = nn.Conv2d(config.in_channels, 6, 5, padding=2)
        self.conv2 = nn.Conv2d(6, 16, 5)
        side = ((config.image_size // 2) - 4) // 2
        self.fc1 = nn.Linear(16 * side * side, 120)
        self.fc2 = nn.Linear(120, 84)
        self.fc3 = nn.Linear(84, config.num_classes)

    def forward(self, pixel_values: torch.Tensor, labels: Optional[torch.Tensor] = None, **_):
        x = pixel_values.float()
        if x.dim() == 3:
            x = x.unsqueeze(1)
        x = F.max_pool2d(F.relu(self.conv1(x)), 2)
        x = F.max_pool2d(F.relu(self.conv2(x)), 2)
        x = F.relu(self.fc1(x.flatten(1)))
        logits = self.fc3(F.relu(self.fc2(x)))
        if labels is None:
            return {"logits": logits}
        return {"loss": F.cross_entropy(logits, labels.long()), "logits": logits}
