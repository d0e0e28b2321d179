"""Per-generation metrics (`/root/reference/src/auralis/common/metrics/performance.py:12-151`): a sliding window of requests,
tokens and audio-seconds, logged every `log_interval` seconds as the reference's line

    Generation metrics | Throughput: <req/s> req/s | <tokens/s> tokens/s | Latency: <ms> ms per second of audio generated

— same quantities, same update rule (one `update_metrics` per chunk that carries a `start_time`, window reset after each
log).  When the engine exposes native counters (`XTTSv2Engine.stats()`), the log line also carries what the reference cannot
see from Python: decode steps, kernel launches and device milliseconds of the GPT and the vocoder since the last line."""
from __future__ import annotations

import logging
import time
from dataclasses import dataclass, field
from typing import Optional

logger = logging.getLogger("auralis_b200.metrics")


@dataclass
class TTSMetricsTracker:
    window_start: float = field(default_factory=time.time)
    last_log_time: float = field(default_factory=time.time)
    log_interval: float = 5.0
    window_tokens: int = 0
    window_audio_seconds: float = 0.0
    window_requests: int = 0
    _native_prev: Optional[dict] = None

    @property
    def requests_per_second(self) -> float:
        elapsed = time.time() - self.window_start
        return self.window_requests / elapsed if elapsed > 0 else 0

    @property
    def tokens_per_second(self) -> float:
        elapsed = time.time() - self.window_start
        return self.window_tokens / elapsed if elapsed > 0 else 0

    @property
    def ms_per_second_of_audio(self) -> float:
        elapsed = (time.time() - self.window_start) * 1000
        return elapsed / self.window_audio_seconds if self.window_audio_seconds > 0 else 0

    def reset_window(self) -> None:
        now = time.time()
        self.last_log_time = now
        self.window_start = now
        self.window_tokens = 0
        self.window_audio_seconds = 0.0
        self.window_requests = 0

    def update_metrics(self, tokens: int, audio_seconds: float) -> bool:
        self.window_tokens += tokens
        self.window_audio_seconds += audio_seconds
        self.window_requests += 1
        return time.time() - self.last_log_time >= self.log_interval

    def line(self, engine=None) -> str:
        msg = (f"Generation metrics | Throughput: {self.requests_per_second:.2f} req/s | {self.tokens_per_second:.1f} tokens/s | "
               f"Latency: {self.ms_per_second_of_audio:.0f}ms per second of audio generated")
        stats = getattr(engine, "stats", None)
        if callable(stats):
            try:
                now = stats()
                prev = self._native_prev or {k: 0 for k in now}
                d = {k: now[k] - prev.get(k, 0) for k in now}
                self._native_prev = now
                msg += (f" | native: {int(d.get('decode_steps', 0))} decode steps, {int(d.get('kernel_launches', 0))} kernels, "
                        f"gpt {d.get('gpt_ms', 0):.0f} ms, vocoder {d.get('vocoder_ms', 0):.0f} ms on the device")
            except Exception:      # noqa: BLE001 — metrics never break generation
                pass
        return msg


metrics = TTSMetricsTracker()


def track(output, engine=None, tracker: TTSMetricsTracker = metrics) -> None:
    """performance.py:track_generation, per yielded chunk"""
    if getattr(output, "start_time", None):
        if tracker.update_metrics(output.token_length or 0, output.array.shape[0] / output.sample_rate):
            logger.info(tracker.line(engine))
            tracker.reset_window()
