"""Event-weight lookup feeding the pooling weights (reference: src/utils/config.py:27-50,
configs/config.yaml:29-32).  Host-side, pure Python: it only turns event names into the floats the
pooling kernels consume."""
from __future__ import annotations

from typing import Any, Dict

DEFAULT_EVENT_WEIGHTS = {"view": 1, "add_to_cart": 5, "purchase": 10}   # configs/config.yaml:29-32

_ALIASES = {"view": "view", "addtocart": "add_to_cart", "add_to_cart": "add_to_cart",
            "purchase": "purchase", "buy": "purchase"}


def get_event_weight(event_name: str, config: Dict[str, Any]) -> int:
    """Case-insensitive event name -> weight; unknown events weigh 1 (config.py:39-50)."""
    weights = config.get("event_weights", {})
    name = event_name.lower()
    return weights.get(_ALIASES.get(name, name), 1)
