"""Request micro-batching in front of the retrieval path.

The reference handles one request per call (`nq = 1`, src/api/server.py:241-244) and its FastAPI handlers
are serial per worker; on this path a single-query search is bound by one pass over the catalog (≈1 ms for
10M×384 on a B200) no matter whether it carries 1 or 128 queries.  `MicroBatcher` lets concurrent callers
share that pass: requests queue up, a worker thread drains up to `max_batch` of them (waiting at most
`max_wait_ms` after the first one), runs ONE batched call and completes every caller's future.  Requests
asking for different `k` are served with the largest `k` of the batch and truncated per request.

`batch_fn(payloads, k) -> sequence of per-request results` is any batched entry point, e.g.
`RetrievalPipeline.retrieve_batch` (payload = interactions) or a wrapper over `VectorDatabase.retrieve_batch`.
Host-side only; no device code here.
"""
from __future__ import annotations

import queue
import threading
import time
from concurrent.futures import Future
from typing import Any, Callable, List, Optional, Sequence, Tuple


class ArrayRow:
    """One request's result as array views: `.ids` i64 [k], `.scores` f32 [k]; `row[:k]` truncates both, and
    `ids, scores = row` unpacks."""
    __slots__ = ("ids", "scores")

    def __init__(self, ids, scores):
        self.ids, self.scores = ids, scores

    def __getitem__(self, sl):
        return ArrayRow(self.ids[sl], self.scores[sl])

    def __len__(self):
        return len(self.ids)

    def __iter__(self):
        return iter((self.ids, self.scores))


class ArrayRows:
    """Batched results as two arrays (ids i64 [B,k], scores f32 [B,k]) that MicroBatcher hands out per request
    without building Python lists: a sequence of B `ArrayRow` views."""
    __slots__ = ("ids", "scores")

    def __init__(self, ids, scores):
        self.ids, self.scores = ids, scores

    def __len__(self):
        return len(self.ids)

    def __getitem__(self, r):
        return ArrayRow(self.ids[r], self.scores[r])

    def __iter__(self):
        return (ArrayRow(self.ids[r], self.scores[r]) for r in range(len(self.ids)))


class MicroBatcher:
    def __init__(self, batch_fn: Callable[[List[Any], int], Sequence[Any]], max_batch: int = 128,
                 max_wait_ms: float = 0.2):
        if max_batch < 1:
            raise ValueError("max_batch must be >= 1")
        self._fn = batch_fn
        self.max_batch = int(max_batch)
        self.max_wait = float(max_wait_ms) * 1e-3
        self._q: "queue.Queue[Optional[Tuple[Any, int, Future]]]" = queue.Queue()
        self._closed = False
        self.batches = 0           # statistics: batches run / requests served
        self.requests = 0
        self._worker = threading.Thread(target=self._run, name="tt-microbatcher", daemon=True)
        self._worker.start()

    # -- caller side -------------------------------------------------------------------------------
    def submit(self, payload: Any, k: int = 10) -> Future:
        """Queues one request; the future resolves to that request's result (first k entries)."""
        if self._closed:
            raise RuntimeError("MicroBatcher is closed")
        fut: Future = Future()
        self._q.put((payload, int(k), fut))
        return fut

    def __call__(self, payload: Any, k: int = 10):
        """Blocking convenience: submit + wait (what a request handler does)."""
        return self.submit(payload, k).result()

    def close(self) -> None:
        if not self._closed:
            self._closed = True
            self._q.put(None)
            self._worker.join()

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    # -- worker ------------------------------------------------------------------------------------
    def _drain(self, first) -> Tuple[List[Tuple[Any, int, Future]], bool]:
        batch, stop = [first], False
        deadline = time.monotonic() + self.max_wait
        while len(batch) < self.max_batch:
            remaining = deadline - time.monotonic()
            try:
                item = self._q.get(timeout=remaining) if remaining > 0 else self._q.get_nowait()
            except queue.Empty:
                break
            if item is None:
                stop = True
                break
            batch.append(item)
        return batch, stop

    def _run(self) -> None:
        while True:
            first = self._q.get()
            if first is None:
                return
            batch, stop = self._drain(first)
            live = [(p, k, f) for p, k, f in batch if f.set_running_or_notify_cancel()]
            if live:
                kmax = max(k for _, k, _ in live)
                try:
                    results = self._fn([p for p, _, _ in live], kmax)
                    if len(results) != len(live):
                        raise RuntimeError(f"batch_fn returned {len(results)} results for {len(live)} requests")
                    for (_, k, f), r in zip(live, results):
                        f.set_result(r[:k])
                except BaseException as e:           # every caller of the batch sees the failure
                    for _, _, f in live:
                        if not f.done():
                            f.set_exception(e)
                self.batches += 1
                self.requests += len(live)
            if stop:
                return
