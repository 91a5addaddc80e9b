"""Dynamic micro-batching in front of ``VQAInference.predict_batch``.

The reference serves one request at a time: ``POST /predict`` calls ``engine.predict`` synchronously inside the event
loop (api/main.py:159-221), so concurrent requests queue behind each other at batch 1, where the fused forward is
latency-bound (0.7 ms for 1 pair, 1.6 ms for 256).  ``MicroBatcher`` collects requests that arrive within
``max_wait_ms`` of the first one (or until ``max_batch``) and answers them with ONE ``predict_batch`` call; every
caller gets the same result dict ``predict`` would have returned (response schema of api/main.py:46-53).

Host logic only (a worker thread and a queue); all arithmetic stays in ``engine.predict_batch``.
"""
from __future__ import annotations

import queue
import threading
import time
from concurrent.futures import Future, InvalidStateError
from typing import Dict, List, Optional, Tuple


def _powers_of_two(limit: int) -> List[int]:
    sizes, s = [], 1
    while s < limit:
        sizes.append(s)
        s *= 2
    return sizes + [limit]


def _resolve(f: Future, result=None, error: Optional[BaseException] = None):
    """Complete a client's Future; a Future the client cancelled (or that is already resolved) is left alone instead of
    raising InvalidStateError inside the worker thread."""
    try:
        if error is not None:
            f.set_exception(error)
        else:
            f.set_result(result)
    except InvalidStateError:
        pass


class MicroBatcher:
    def __init__(self, engine, max_batch: int = 64, max_wait_ms: float = 2.0, pad_to: Optional[List[int]] = "pow2"):
        """``pad_to``: allowed batch sizes, ascending: a batch is padded with copies of its last request up to the next
        allowed size, so that the engine replays a handful of captured CUDA graphs (and keeps a handful of plan
        workspaces) instead of one per batch size.  Default: the powers of two up to ``max_batch``.  None = run
        whatever size was collected (the engine's plan / graph caches are LRUs, so this stays bounded too)."""
        if max_batch < 1:
            raise ValueError("max_batch must be >= 1")
        self.engine = engine
        self.max_batch = int(max_batch)
        self.max_wait = float(max_wait_ms) / 1e3
        if isinstance(pad_to, str):
            if pad_to != "pow2":
                raise ValueError("pad_to must be a list of batch sizes, None or 'pow2'")
            pad_to = _powers_of_two(self.max_batch)
        self.pad_to = sorted(int(s) for s in pad_to) if pad_to else None
        self._q: "queue.Queue[Optional[Tuple[object, str, int, Future]]]" = queue.Queue()
        self._closed = False
        self.batches = 0          # predict_batch calls made
        self.requests = 0         # requests answered
        self._thread = threading.Thread(target=self._loop, name="vqa-microbatcher", daemon=True)
        self._thread.start()

    # ------------------------------------------------------------------ client side
    def submit(self, image, question: str, top_k: int = 5) -> Future:
        if self._closed:
            raise RuntimeError("MicroBatcher is closed")
        f: Future = Future()
        self._q.put((image, question, int(top_k), f))
        return f

    def predict(self, image, question: str, top_k: int = 5, timeout: Optional[float] = None) -> Dict:
        """Blocking form with the signature of ``VQAInference.predict``."""
        return self.submit(image, question, top_k).result(timeout)

    def close(self):
        if not self._closed:
            self._closed = True
            self._q.put(None)
            self._thread.join()

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    # ------------------------------------------------------------------ worker
    def _collect(self):
        first = self._q.get()
        if first is None:
            return None
        batch = [first]
        deadline = time.monotonic() + self.max_wait
        while len(batch) < self.max_batch:
            left = deadline - time.monotonic()
            try:
                item = self._q.get(timeout=left) if left > 0 else self._q.get_nowait()
            except queue.Empty:
                break
            if item is None:
                self._q.put(None)      # close() arrived behind live requests: answer them, then stop
                break
            batch.append(item)
        return batch

    def _run(self, items, top_k: int):
        images = [it[0] for it in items]
        questions = [it[1] for it in items]
        n = len(items)
        if self.pad_to:
            target = next((s for s in self.pad_to if s >= n), n)
            images += [images[-1]] * (target - n)
            questions += [questions[-1]] * (target - n)
        try:
            results = self.engine.predict_batch(images, questions, top_k=top_k)
            self.batches += 1
            self.requests += n
        except Exception as e:             # one bad request fails its batch: retry singly so the others still get answers
            if n == 1:
                _resolve(items[0][3], error=e)
                return
            for it in items:
                try:
                    r = self.engine.predict_batch([it[0]], [it[1]], top_k=top_k)[0]
                except Exception as e1:
                    _resolve(it[3], error=e1)
                    continue
                self.batches += 1
                self.requests += 1
                _resolve(it[3], r)
            return
        for it, r in zip(items, results[:n]):
            _resolve(it[3], r)

    def _loop(self):
        while True:
            batch = self._collect()
            if batch is None:
                break
            by_k: Dict[int, list] = {}
            for it in batch:
                if it[3].set_running_or_notify_cancel():     # skip requests whose client cancelled while they were queued
                    by_k.setdefault(it[2], []).append(it)
            for k, items in by_k.items():
                self._run(items, k)
        # fail whatever is still queued after close()
        while True:
            try:
                it = self._q.get_nowait()
            except queue.Empty:
                break
            if it is not None and it[3].set_running_or_notify_cancel():
                _resolve(it[3], error=RuntimeError("MicroBatcher is closed"))
