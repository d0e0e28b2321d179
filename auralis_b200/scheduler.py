"""TwoPhaseScheduler — same contract as the reference's
(`/root/reference/src/auralis/common/scheduling/two_phase_scheduler.py:10-459`):

* ``run(inputs, request_id, first_phase_fn, second_phase_fn)`` is an async generator;
* phase 1 (``first_phase_fn(inputs) -> {'parallel_inputs': [...], 'request': ...}``) runs once per request;
* phase 2 runs one async generator per parallel input, at most ``second_phase_concurrency`` at a time;
* outputs are yielded strictly in sequence-index order, each generator's items in their own order
  (two_phase_scheduler.py:308-388, App. B.13);
* ``request_timeout`` / ``generator_timeout`` raise ``TimeoutError``; a failing phase fails the request and
  the exception propagates to the caller (two_phase_scheduler.py:279-291,439-440).

What changed: the reference polls its buffers every 10 ms (:350) and funnels everything through a queue
and worker tasks because vLLM did the batching; here batching happens inside the native engine, so this
class only orders, bounds and times out — it is event-driven (asyncio.Condition), no polling.
"""
from __future__ import annotations

import asyncio
import time
from typing import Any, AsyncGenerator, Awaitable, Callable, Dict, List


class TwoPhaseScheduler:
    def __init__(self, second_phase_concurrency: int = 10, request_timeout: float = None,
                 generator_timeout: float = None):
        self.second_phase_concurrency = second_phase_concurrency
        self.request_timeout = request_timeout
        self.generator_timeout = generator_timeout
        self._sem: asyncio.Semaphore | None = None
        self._sem_loop = None
        self.is_running = False
        self.active = 0

    def _semaphore(self) -> asyncio.Semaphore:
        loop = asyncio.get_running_loop()
        if self._sem is None or self._sem_loop is not loop:
            self._sem = asyncio.Semaphore(self.second_phase_concurrency)
            self._sem_loop = loop
        return self._sem

    async def start(self):
        self.is_running = True

    async def run(self, inputs: Any, request_id: str,
                  first_phase_fn: Callable[[Any], Awaitable[Dict]],
                  second_phase_fn: Callable[[Dict], AsyncGenerator]) -> AsyncGenerator[Any, None]:
        if not self.is_running:
            await self.start()
        t_start = time.time()

        def remaining():
            if self.request_timeout is None:
                return None
            return max(0.0, self.request_timeout - (time.time() - t_start))

        # ---- phase 1
        try:
            ctx = await asyncio.wait_for(first_phase_fn(inputs), timeout=remaining())
        except asyncio.TimeoutError:
            raise TimeoutError(f"first phase timed out for request {request_id}")
        parallel_inputs: List[Dict] = ctx["parallel_inputs"]
        n = len(parallel_inputs)
        buffers: List[List[Any]] = [[] for _ in range(n)]
        done = [False] * n
        errors: List[BaseException] = []
        cond = asyncio.Condition()
        sem = self._semaphore()

        async def pump(idx: int, gen_input: Dict):
            try:
                async with sem:
                    self.active += 1
                    try:
                        agen = second_phase_fn(gen_input)
                        while True:
                            try:
                                item = await asyncio.wait_for(agen.__anext__(), timeout=self.generator_timeout)
                            except StopAsyncIteration:
                                break
                            except asyncio.TimeoutError:
                                raise TimeoutError(f"generator {idx} of request {request_id} timed out")
                            async with cond:
                                buffers[idx].append(item)
                                cond.notify_all()
                    finally:
                        self.active -= 1
            except BaseException as e:      # noqa: BLE001 — surfaced to the caller below
                errors.append(e)
            finally:
                async with cond:
                    done[idx] = True
                    cond.notify_all()

        tasks = [asyncio.ensure_future(pump(i, gi)) for i, gi in enumerate(parallel_inputs)]
        # ---- ordered drain
        try:
            for idx in range(n):
                pos = 0
                while True:
                    async with cond:
                        while pos >= len(buffers[idx]) and not done[idx] and not errors:
                            try:
                                await asyncio.wait_for(cond.wait(), timeout=remaining())
                            except asyncio.TimeoutError:
                                raise TimeoutError(f"request {request_id} timed out")
                        if errors:
                            raise errors[0]
                        items = buffers[idx][pos:]
                        finished = done[idx]
                    for it in items:
                        yield it
                    pos += len(items)
                    if finished and pos >= len(buffers[idx]):
                        break
            if errors:
                raise errors[0]
        finally:
            for t in tasks:
                if not t.done():
                    t.cancel()
            await asyncio.gather(*tasks, return_exceptions=True)

    async def shutdown(self):
        self.is_running = False
