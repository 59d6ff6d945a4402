"""Overlapped serving loop around `G2VLMFast.recon` (the caller side of the hot path: reference
`inference_recon.py` feeds scenes one after another through `G2VLM.recon`, g2vlm.py:1240-1303).

One B200 computes scene i on the compute stream while
  * the raw views of scene i+1 cross PCIe on an upload stream into a device staging slot, and
  * the point maps of scene i-1 return to pinned host memory on a download stream,
so the host<->device copies (52 MB up, 155 MB down for 16 views of 518 px) leave the critical path except for
the first upload and the last download.  Results are bit-identical to calling `recon` directly: the same
kernels run on the same data, only the copies move to side streams.  Nothing here touches the oracle."""
from __future__ import annotations

from typing import Dict, List, Optional, Sequence

import torch

RESULT_KEYS = ("points", "local_points", "global_points", "camera_poses")


class ReconServer:
    def __init__(self, model, tokenizer, new_token_ids: Dict[str, int], keys: Sequence[str] = RESULT_KEYS,
                 slots: int = 2):
        assert slots >= 2
        self.model, self.tok, self.ids = model, tokenizer, dict(new_token_ids)
        self.keys, self.slots = tuple(keys), slots
        self.up, self.down = torch.cuda.Stream(), torch.cuda.Stream()
        self.dev_in: List[Optional[torch.Tensor]] = [None] * slots
        self.dev_out: List[Dict[str, torch.Tensor]] = [dict() for _ in range(slots)]
        self.host_out: List[Dict[str, torch.Tensor]] = [dict() for _ in range(slots)]
        ev = lambda: [torch.cuda.Event() for _ in range(slots)]
        self.up_done, self.compute_done, self.down_done = ev(), ev(), ev()
        self.submitted = 0

    @staticmethod
    def _like(store: dict, key: str, ref: torch.Tensor, **kw) -> torch.Tensor:
        t = store.get(key)
        if t is None or t.shape != ref.shape or t.dtype != ref.dtype:
            t = store[key] = torch.empty(ref.shape, dtype=ref.dtype, **kw)
        return t

    def submit(self, views_host: torch.Tensor) -> int:
        """Enqueue one scene ((N,3,H,W) fp32 in [0,1]; pinned memory makes the upload asynchronous).
        Returns the ticket to pass to `result`.  At most `slots` tickets may be outstanding."""
        s = self.submitted % self.slots
        cur = torch.cuda.current_stream()
        if self.dev_in[s] is None or self.dev_in[s].shape != views_host.shape:
            self.dev_in[s] = torch.empty(views_host.shape, dtype=views_host.dtype, device=self.model.device)
            # a fresh block may be memory that kernels already enqueued on the compute stream still use (the caching
            # allocator hands freed blocks back in stream order of the ALLOCATING stream): the first upload into it
            # must not overtake them
            self.up.wait_stream(cur)
        self.up.wait_event(self.compute_done[s])            # the slot's previous scene has consumed its views
        with torch.cuda.stream(self.up):
            self.dev_in[s].copy_(views_host, non_blocking=True)
            self.up_done[s].record(self.up)
        cur.wait_event(self.up_done[s])
        pred = self.model.recon(self.tok, self.ids, None, self.dev_in[s])
        cur.wait_event(self.down_done[s])                   # the slot's previous download has left dev_out[s]
        for k in self.keys:
            self._like(self.dev_out[s], k, pred[k], device=pred[k].device).copy_(pred[k])
        self.compute_done[s].record(cur)
        with torch.cuda.stream(self.down):
            self.down.wait_event(self.compute_done[s])
            for k in self.keys:
                self._like(self.host_out[s], k, pred[k], pin_memory=True).copy_(self.dev_out[s][k], non_blocking=True)
            self.down_done[s].record(self.down)
        self.submitted += 1
        return self.submitted - 1

    def result(self, ticket: int) -> Dict[str, torch.Tensor]:
        """Block until the scene's results are in pinned host memory and return them (views into the slot:
        copy them out before `slots` further scenes are submitted)."""
        assert self.submitted - self.slots <= ticket < self.submitted, "ticket expired or not submitted"
        s = ticket % self.slots
        self.down_done[s].synchronize()
        return self.host_out[s]

    def drain(self) -> None:
        """Make the current stream wait for every outstanding download (so an event recorded after this call
        covers the whole pipeline), then block the host."""
        cur = torch.cuda.current_stream()
        for e in self.down_done:
            cur.wait_event(e)
        cur.synchronize()
