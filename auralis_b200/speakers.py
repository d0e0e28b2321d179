"""Speaker-slot bookkeeping for the native engine's conditioning cache (pure Python, no GPU).

The native library keeps `max_speakers` slots of (GPT conditioning latents, d-vector, vocoder bias vectors); a chunk
refers to its speaker by slot number.  This table maps a content key (hash of the reference audio + conditioning
parameters, or of the latents themselves) to a slot — the per-speaker cache the reference builds with
`prepare_for_streaming_generation` (`core/tts.py:91-105`) — and guarantees three things the slot
numbers alone cannot:

  * a slot whose contents are still being computed is never handed out as a hit (waiters block on the same future);
  * a slot referenced by a queued or running chunk is pinned and never evicted;
  * a (key, slot) pair handed to a caller earlier can be re-validated later (`holds`), so a stale slot number is
    detected instead of silently selecting another speaker's voice.
"""
from __future__ import annotations

import threading
from concurrent.futures import Future
from typing import Dict, List, Optional, Tuple


class SpeakerSlotsFull(RuntimeError):
    """Every slot is pinned by a chunk in flight or still being computed."""


class SpeakerSlots:
    def __init__(self, n_slots: int):
        if n_slots < 1:
            raise ValueError("n_slots must be >= 1")
        self.n = n_slots
        self._lock = threading.Lock()
        self._slot_of: Dict[str, int] = {}
        self._key_of: Dict[int, str] = {}
        self._lru: List[str] = []                     # least recently used first
        self._pins: Dict[int, int] = {}               # slot -> chunks in flight
        self._ready: Dict[str, Future] = {}           # key -> resolves when the slot's contents are valid

    # ---- lookup / allocation ----------------------------------------------------------------
    def acquire(self, key: str) -> Tuple[int, Optional[Future], bool]:
        """-> (slot, future, owner).  owner=True: the caller must fill the slot and then call `ready(key)` or
        `failed(key, exc)`; owner=False: wait on `future` (None if the slot is already valid)."""
        with self._lock:
            if key in self._slot_of:
                self._touch(key)
                f = self._ready.get(key)
                return self._slot_of[key], (f if f is not None and not f.done() else None), False
            slot = self._free_slot_locked()
            self._slot_of[key] = slot
            self._key_of[slot] = key
            self._lru.append(key)
            f = Future()
            self._ready[key] = f
            return slot, f, True

    def _touch(self, key: str) -> None:
        self._lru.remove(key)
        self._lru.append(key)

    def _free_slot_locked(self) -> int:
        if len(self._slot_of) < self.n:
            used = set(self._key_of)
            return next(s for s in range(self.n) if s not in used)
        for victim in self._lru:                      # oldest unpinned, fully computed entry
            s = self._slot_of[victim]
            f = self._ready.get(victim)
            if self._pins.get(s, 0) == 0 and (f is None or f.done()):
                self._drop_locked(victim)
                return s
        raise SpeakerSlotsFull(f"all {self.n} speaker slots are in use by chunks in flight (raise max_speakers)")

    def _drop_locked(self, key: str) -> None:
        s = self._slot_of.pop(key)
        self._key_of.pop(s, None)
        self._lru.remove(key)
        self._ready.pop(key, None)

    def ready(self, key: str) -> None:
        with self._lock:
            f = self._ready.get(key)
        if f is not None and not f.done():
            f.set_result(True)

    def failed(self, key: str, exc: BaseException) -> None:
        """The owner could not fill the slot: forget the entry and wake the waiters with the error."""
        with self._lock:
            f = self._ready.get(key)
            if key in self._slot_of and self._pins.get(self._slot_of[key], 0) == 0:
                self._drop_locked(key)
        if f is not None and not f.done():
            f.set_exception(exc)

    # ---- validation / pinning ---------------------------------------------------------------
    def holds(self, key: str, slot: int) -> bool:
        with self._lock:
            f = self._ready.get(key)
            return self._slot_of.get(key) == slot and (f is None or (f.done() and f.exception() is None))

    def pin(self, key: str, slot: int) -> bool:
        """Pins `slot` for one chunk if it still holds `key`; False means the caller has to re-register the speaker."""
        with self._lock:
            f = self._ready.get(key)
            if self._slot_of.get(key) != slot or (f is not None and not (f.done() and f.exception() is None)):
                return False
            self._pins[slot] = self._pins.get(slot, 0) + 1
            self._touch(key)
            return True

    def unpin(self, slot: int) -> None:
        with self._lock:
            n = self._pins.get(slot, 0)
            if n <= 1:
                self._pins.pop(slot, None)
            else:
                self._pins[slot] = n - 1

    def pinned(self, slot: int) -> int:
        with self._lock:
            return self._pins.get(slot, 0)

    def __len__(self) -> int:
        with self._lock:
            return len(self._slot_of)
