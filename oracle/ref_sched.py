"""TEST INFRASTRUCTURE — container-only loader for the reference's own TwoPhaseScheduler
(`/root/reference/src/auralis/common/scheduling/two_phase_scheduler.py`, imported unmodified).  Its only import that is
missing here is the coloured logger (`colorama`); that one module is replaced by a stub returning a standard logger.
Used by tests/test_scheduler_vs_reference.py to compare `auralis_b200.scheduler.TwoPhaseScheduler` with it scenario by
scenario (ordering, concurrency bound, error and timeout behaviour)."""
from __future__ import annotations

import importlib
import logging
import os
import sys
import types

from . import ref_import


def load():
    if not ref_import.available():
        raise RuntimeError("reference tree not mounted")
    ref_import.load()
    base = os.path.join(ref_import.REF_SRC, "auralis", "common")
    ref_import._stub("auralis.common.definitions", os.path.join(base, "definitions"))
    ref_import._stub("auralis.common.scheduling", os.path.join(base, "scheduling"))
    ref_import._stub("auralis.common.logging", os.path.join(base, "logging"))
    if "auralis.common.logging.logger" not in sys.modules:
        m = types.ModuleType("auralis.common.logging.logger")
        m.setup_logger = lambda name=None, *a, **k: logging.getLogger("auralis.ref")
        sys.modules["auralis.common.logging.logger"] = m
    mod = importlib.import_module("auralis.common.scheduling.two_phase_scheduler")
    return mod.TwoPhaseScheduler
