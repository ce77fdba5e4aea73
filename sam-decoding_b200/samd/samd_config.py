"""SamdConfig and the small state holders (reference: samd/samd_config.py:9-96)."""
import json
import os
from dataclasses import dataclass, field
from enum import Enum
from typing import Any, Dict, List, Literal, Optional

import torch


def load_token_recycle(tree_path: Optional[str] = None) -> List[List[int]]:
    """Children lists of the static draft tree (config/token_recycle.json, 61 nodes)."""
    name = tree_path or "token_recycle.json"
    with open(os.path.join(os.path.dirname(__file__), "config", name)) as f:
        adj = json.load(f)["tree_adj"]
    return [adj[str(i)] for i in range(len(adj))]


def load_eagle(tree_model_path: str, tree_path: Optional[str] = None):
    name = tree_path or "eagle.json"
    with open(os.path.join(os.path.dirname(__file__), "config", name)) as f:
        tree = json.load(f)["tree_choices"]
    with open(os.path.join(tree_model_path, "config.json")) as f:
        return tree, json.load(f)


def load_eagle2(tree_model_path: str):
    with open(os.path.join(tree_model_path, "config.json")) as f:
        return json.load(f)


@dataclass
class SamdConfig:
    n_predicts: int = field(default=40)
    max_predicts: int = field(default=70)
    len_threshold: int = field(default=5)
    len_bias: int = field(default=5)
    cache_type: Literal["dynamic", "static"] = field(default="static")
    use_last_hidden_states: bool = field(default=False)
    tree_method: Literal["token_recycle", "eagle", "eagle2"] = field(default="token_recycle")
    tree_model_path: Optional[str] = field(default=None)
    tree_path: Optional[str] = field(default=None)
    tree: Optional[List[List[int]]] = field(default=None)
    tree_config: Optional[Dict[str, Any]] = field(default=None)

    def __post_init__(self):
        if self.tree is not None:
            return
        if self.tree_method == "token_recycle":
            self.tree = load_token_recycle(self.tree_path)
        elif self.tree_method == "eagle":
            self.tree, self.tree_config = load_eagle(self.tree_model_path, self.tree_path)
            self.use_last_hidden_states = True
        elif self.tree_method == "eagle2":
            self.tree_config = load_eagle2(self.tree_model_path)
            self.use_last_hidden_states = True
        else:
            raise ValueError


class ForwardType(str, Enum):
    prefill = "prefill"
    seq_decode = "seq_decode"
    tree_decode = "tree_decode"


class ForwardState:
    def __init__(self, forward_type: Optional[ForwardType]) -> None:
        self.forward_type = forward_type


class MaskState:
    def __init__(self, mask: Optional[torch.Tensor]) -> None:
        self.mask = mask

    def set_state(self, mask: Optional[torch.Tensor]) -> None:
        self.mask = mask
