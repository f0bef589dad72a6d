"""NVTX ranges around the phases of the two hot paths (the reference has no tracing at all).  Off by default —
set GSD_NVTX=1 to emit them (they show up in Nsight Systems / `ncu --nvtx`); a disabled range costs one dict lookup."""
import contextlib
import os

import torch

ENABLED = os.environ.get("GSD_NVTX", "0") == "1"


@contextlib.contextmanager
def range(name):
    if ENABLED:
        torch.cuda.nvtx.range_push(name)
        try:
            yield
        finally:
            torch.cuda.nvtx.range_pop()
    else:
        yield
