"""`/retrieve` hot path kept on the device.

The reference serves a request as `encoder.encode_buyer(interactions)` -> `.cpu().numpy()` ->
`vector_db.retrieve(embedding, k)` (src/api/server.py:241-244): every history item is re-encoded through
the text tower on every request (src/inference/encoder.py:276-292) and the pooled embedding takes a host
round trip before the search.  The catalog rows ARE those item embeddings (the same ItemTower output,
encoder.py:235-240), so here a request is: history product ids -> catalog row indices (host dict) ->
fused gather + pooling + L2 norm out of the device-resident fp32 table (tt_pool_*_gather) -> exact top-K
(tt_flat_search) with the pooled embedding never leaving the device.

History semantics follow encode_buyer: sort by timestamp when every interaction has one, keep the last
`max_interaction_history` (configs/config.yaml:14), event weights through get_event_weight.  One deliberate
difference: a product id that is not in the catalog pools as an all-zero row here (and is counted in the
`unknown` result of encode_histories), whereas the reference would encode the empty text for it.
"""
from __future__ import annotations

from typing import Any, Dict, List, Optional, Sequence, Tuple

import numpy as np
import torch

from .buyer_tower import BuyerTower
from .config import DEFAULT_EVENT_WEIGHTS, get_event_weight
from .vector_db import PendingSearch, VectorDatabase


class RetrievalPipeline:
    def __init__(self, buyer_tower: BuyerTower, vector_db: VectorDatabase, config: Optional[Dict[str, Any]] = None,
                 max_interaction_history: int = 100):
        if vector_db.index is None:
            raise ValueError("Index not built. Call build_index() or load_index() first.")
        self.tower = buyer_tower
        self.db = vector_db
        self.index = vector_db.index
        self.config = config if config is not None else {"event_weights": dict(DEFAULT_EVENT_WEIGHTS)}
        self.max_history = int(max_interaction_history)
        self.table = self.index.xn                              # [N, D] fp32, device resident
        self.item_logits = buyer_tower.precompute_item_logits(self.table)   # None in weighted_avg mode
        self._pins: Dict[Tuple[int, int], Tuple[torch.Tensor, torch.Tensor]] = {}

    # -- host side: interactions -> (row indices, weights), padded with (-1, 0) ---------------------
    def encode_histories(self, batch: Sequence[Sequence[Dict[str, Any]]]) -> Tuple[np.ndarray, np.ndarray, int]:
        lens, rows_all, w_all, unknown = [], [], [], 0
        for interactions in batch:
            if all(it.get("timestamp") is not None for it in interactions):
                interactions = sorted(interactions, key=lambda x: x["timestamp"])       # encoder.py:263-264
            interactions = interactions[-self.max_history:]                             # encoder.py:267-268
            rows = [self.db.id_to_index.get(it["product_id"], -1) for it in interactions]
            unknown += sum(1 for r in rows if r < 0)
            rows_all.append(rows)
            w_all.append([float(get_event_weight(it["event_type"], self.config)) for it in interactions])
            lens.append(len(rows))
        S = max(max(lens, default=0), 1)
        idx = np.full((len(batch), S), -1, np.int64)
        w = np.zeros((len(batch), S), np.float32)
        for b, (rows, ws) in enumerate(zip(rows_all, w_all)):
            idx[b, :len(rows)] = rows
            w[b, :len(ws)] = ws
        return idx, w, unknown

    # -- device side ------------------------------------------------------------------------------
    def encode_device(self, indices: torch.Tensor, weights: torch.Tensor) -> torch.Tensor:
        """history rows i64 [B,S] + weights f32 [B,S] (device) -> buyer embeddings [B,D] (device)."""
        return self.tower.forward_gather(self.table, indices, weights, self.item_logits)

    def retrieve_device_async(self, indices: torch.Tensor, weights: torch.Tensor, k: int) -> PendingSearch:
        k = min(k, self.index.ntotal)                                                  # vector_db.py:159
        return self.index.search_async(self.encode_device(indices, weights), k)

    def retrieve_arrays(self, batch: Sequence[Sequence[Dict[str, Any]]], k: int = 10):
        """-> (scores f32 [B,k'], row indices i64 [B,k']) numpy; one H2D of the padded history, one D2H."""
        idx, w, _ = self.encode_histories(batch)
        key = idx.shape
        pins = self._pins.get(key)
        if pins is None:
            if len(self._pins) > 16:
                self._pins.clear()
            pins = self._pins[key] = (torch.empty(key, dtype=torch.int64, pin_memory=True),
                                      torch.empty(key, dtype=torch.float32, pin_memory=True))
        pins[0].copy_(torch.from_numpy(idx))
        pins[1].copy_(torch.from_numpy(w))
        dev = self.index.device
        scores, ids, _ = self.retrieve_device_async(pins[0].to(dev, non_blocking=True), pins[1].to(dev, non_blocking=True),
                                                    k).result()
        return scores.cpu().numpy(), ids.cpu().numpy()

    def retrieve(self, interactions: Sequence[Dict[str, Any]], k: int = 10) -> List[Tuple[str, float]]:
        """One request, the shape of server.py:241-244 -> [(product_id, score)] in descending score."""
        scores, ids = self.retrieve_arrays([interactions], k)
        pids = self.db.product_ids
        return [(pids[i], float(s)) for i, s in zip(ids[0].tolist(), scores[0].tolist()) if 0 <= i < len(pids)]

    def retrieve_batch(self, batch: Sequence[Sequence[Dict[str, Any]]], k: int = 10) -> List[List[Tuple[str, float]]]:
        scores, ids = self.retrieve_arrays(batch, k)
        pids = self.db.product_ids
        n = len(pids)
        return [[(pids[i], s) for i, s in zip(ri, rs) if 0 <= i < n] for rs, ri in zip(scores.tolist(), ids.tolist())]
