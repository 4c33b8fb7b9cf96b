"""Fork/join of independent GPU work onto side streams.

At batch 1 the Styl3R encoder is a chain of ~3000 kernels that each fill a fraction of the 148 SMs (M = 257/514-row
GEMMs, 16x16...64x64 pyramid levels), so the forward is bound by the *length of the dependency chain*, not by FLOPs.
The model has wide independent branches - content ViT vs style ViT, backbone decoder vs token stylizer decoder,
`dec_blocks` (view 0) vs `dec_blocks2` (views >= 1) inside every decoder layer, and 5 DPT pyramids per view
(encoder_noposplat_multi_token_style.py:144-176 runs all of them sequentially on one stream).  `fork_join` puts such
branches on side streams between two events; inside a CUDA-graph capture (GraphedEncoder) the events become graph
edges, so the replayed graph runs the branches concurrently with no host involvement.

Memory discipline (torch caching allocator, no record_stream needed): every side stream starts by waiting on an event
recorded on the parent stream *after* everything enqueued so far, and the parent waits for every branch at the join.
A block allocated on a side stream and freed after the join can therefore only be reused by later work of that side
stream, which again starts behind a newer parent event.  Callers must keep the *inputs* of a branch alive until
`fork_join` returns (they do: inputs are locals of the calling function).
"""
from __future__ import annotations

import threading
from typing import Callable, List, Sequence

import torch

_pools: dict = {}
_tls = threading.local()
enabled = True  # set False to run the branches sequentially on the current stream (debugging / A-B timing)


def _side_streams(device, parent: torch.cuda.Stream, depth: int, n: int) -> List[torch.cuda.Stream]:
    # children are private to their (parent stream, nesting depth): a block allocated on a child is only ever handed to
    # work that is ordered behind the parent (see the module docstring).  The depth matters: branch 0 of a fork runs on
    # the parent stream itself, so a nested fork inside it has the same parent - without the depth in the key it would
    # reuse the streams on which the outer fork's sibling branches are still running and queue behind them
    # (measured: backbone decoder 2.26 ms + stylizer decoder 1.43 ms ran in 3.72 ms "concurrently").
    key = (torch.device(device).index, parent.cuda_stream, depth)
    pool = _pools.setdefault(key, [])
    while len(pool) < n:
        pool.append(torch.cuda.Stream(device=device))
    return pool[:n]


def fork_join(fns: Sequence[Callable[[], object]], device=None, max_streams: int = 8, parallel: bool = True) -> list:
    """Run the callables as concurrent branches: fns[0] on the current stream, the others round-robin on side streams;
    returns their results in order after joining everything back into the current stream.  Re-entrant (a branch may
    fork again: every (stream, nesting depth) owns its private child streams)."""
    fns = list(fns)
    if len(fns) <= 1 or not enabled or not parallel:
        return [fn() for fn in fns]
    device = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
    n_side = min(len(fns) - 1, max_streams)
    cur = torch.cuda.current_stream(device)
    depth = getattr(_tls, "depth", 0)
    side = _side_streams(device, cur, depth, n_side)
    start = torch.cuda.Event()
    start.record(cur)
    for s in side:
        s.wait_event(start)
    results: list = [None] * len(fns)
    _tls.depth = depth + 1
    try:
        # side branches first so that their kernels are enqueued before the (usually longest) main branch
        for i in range(1, len(fns)):
            with torch.cuda.stream(side[(i - 1) % n_side]):
                results[i] = fns[i]()
        results[0] = fns[0]()
    finally:
        _tls.depth = depth
    for s in side:
        done = torch.cuda.Event()
        done.record(s)
        cur.wait_event(done)
    return results
