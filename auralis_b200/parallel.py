"""Data-parallel sharding of the hot path across the GPUs of one node.

One process per GPU (torchrun), every rank holds a full replica of the weights.  Text chunks are
independent units, so there is no collective inside the GPT or the vocoder; the only exchange is the
variable-length **all-gather of output waveforms** (lengths first, then max-padded payloads) over NCCL /
NVSwitch, after which rank 0 holds every request's audio in request order.

`torch.distributed` is plumbing here: tensors are containers for the gather buffers only.
"""
from __future__ import annotations

import os
from typing import Callable, Dict, List, Sequence, Tuple

import numpy as np
import torch
import torch.distributed as dist


def init_from_env(backend: str | None = None) -> Tuple[int, int, int]:
    """(rank, world, local_rank) from torchrun's environment; initialises the default group when world > 1."""
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1 and not dist.is_initialized():
        if backend is None:
            backend = "nccl" if torch.cuda.is_available() else "gloo"
        if backend == "nccl":
            torch.cuda.set_device(local)
            dist.init_process_group(backend=backend, rank=rank, world_size=world, device_id=torch.device("cuda", local))
        else:
            dist.init_process_group(backend=backend, rank=rank, world_size=world)
    return rank, world, local


def lpt_partition(costs: Sequence[float], world: int) -> List[List[int]]:
    """Longest-processing-time-first assignment of work items to ranks (cost ∝ characters ≈ tokens).
    Deterministic: ties break on the lower item index, then the lower rank."""
    order = sorted(range(len(costs)), key=lambda i: (-costs[i], i))
    load = [0.0] * world
    out: List[List[int]] = [[] for _ in range(world)]
    for i in order:
        r = min(range(world), key=lambda k: (load[k], k))
        out[r].append(i)
        load[r] += costs[i]
    for lst in out:
        lst.sort()
    return out


_HOST_BUFFERS: Dict[str, torch.Tensor] = {}


def _host_buffer(tag: str, n: int, pinned: bool) -> torch.Tensor:
    """Cached host staging buffer (pinned when a GPU is in play) — the gather runs every step with the same sizes."""
    b = _HOST_BUFFERS.get(tag)
    if b is None or b.numel() < n or b.is_pinned() != pinned:
        b = torch.empty(n, dtype=torch.float32, pin_memory=pinned)
        _HOST_BUFFERS[tag] = b
    return b[:n]


def gather_waveforms(local: Dict[int, np.ndarray], n_items: int, device: torch.device | None = None,
                     group=None, dst: int | None = None, tag: str = "") -> List[np.ndarray] | None:
    """All-gather {item index -> waveform} from every rank; returns the list in item order on every rank.

    Wire format: int64 [2*k] (index, length) table, then one flat fp32 payload per rank, both max-padded so a
    single `all_gather_into_tensor` moves each.  With NCCL the payload lives in HBM (`device`), so the
    transfer is GPU->NVSwitch->GPU; with gloo (CPU tests) it stays on the host.
    `dst`: only that rank copies the gathered payload back to the host and returns the list (others return None).
    `tag`: names the cached staging buffers — callers that overlap consecutive gathers alternate two tags."""
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    if world == 1:
        return [local[i] for i in range(n_items)]
    dev = device if device is not None else torch.device("cpu")
    idx = sorted(local.keys())
    table = torch.tensor([[i, int(local[i].shape[0])] for i in idx], dtype=torch.int64).reshape(-1)
    meta = torch.tensor([len(idx), int(sum(local[i].shape[0] for i in idx))], dtype=torch.int64, device=dev)
    metas = torch.empty(world * 2, dtype=torch.int64, device=dev)
    dist.all_gather_into_tensor(metas, meta, group=group)
    metas = metas.cpu().view(world, 2)
    max_k, max_n = int(metas[:, 0].max()), int(metas[:, 1].max())
    tpad = torch.zeros(max(1, max_k) * 2, dtype=torch.int64, device=dev)
    tpad[: table.numel()] = table.to(dev)
    tables = torch.empty(world * tpad.numel(), dtype=torch.int64, device=dev)
    dist.all_gather_into_tensor(tables, tpad, group=group)
    on_gpu = dev.type == "cuda"
    # stage this rank's audio in one (cached, pinned) host buffer -> a single H2D copy
    stage = _host_buffer("send" + tag, max(1, max_n), on_gpu)
    off = 0
    for i in idx:
        n = int(local[i].shape[0])
        stage[off: off + n] = torch.from_numpy(np.ascontiguousarray(local[i], dtype=np.float32))
        off += n
    payload = stage.to(dev, non_blocking=True) if on_gpu else stage
    allp = torch.empty(world * payload.numel(), dtype=torch.float32, device=dev)
    dist.all_gather_into_tensor(allp, payload, group=group)
    if dst is not None and dist.get_rank(group) != dst:
        return None
    tables = tables.cpu().view(world, -1)
    if on_gpu:                                    # one D2H into a cached pinned buffer; the results are views into it
        recv = _host_buffer("recv" + tag, allp.numel(), True)
        recv.copy_(allp, non_blocking=False)
        allp_h = recv.view(world, -1).numpy()
    else:
        allp_h = allp.view(world, -1).numpy()
    out: List[np.ndarray | None] = [None] * n_items
    for r in range(world):
        off = 0
        for j in range(int(metas[r, 0])):
            i, n = int(tables[r, 2 * j]), int(tables[r, 2 * j + 1])
            out[i] = allp_h[r, off: off + n] if on_gpu else allp_h[r, off: off + n].copy()
            off += n
    missing = [i for i, w in enumerate(out) if w is None]
    if missing:
        raise RuntimeError(f"gather_waveforms: items {missing} were produced by no rank")
    return out  # type: ignore[return-value]


def run_sharded(items: Sequence, costs: Sequence[float], synth: Callable[[List[int]], Dict[int, np.ndarray]],
                device: torch.device | None = None, group=None, dst: int | None = None) -> List[np.ndarray] | None:
    """Shard `items` by LPT, run `synth(indices)` on this rank's share, all-gather the audio
    (`dst`: only that rank materialises the gathered list on the host; the others return None)."""
    rank = dist.get_rank(group) if dist.is_initialized() else 0
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    mine = lpt_partition(costs, world)[rank]
    local = synth(mine) if mine else {}
    return gather_waveforms(local, len(items), device, group, dst)
