"""2-GPU parity (NCCL over NVLink): view-sharded recon of ONE scene == single-GPU recon of the same scene.
Skipped on boxes with < 2 GPUs (run with `gpurun --gpus 2`)."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from g2vlm_b200 import schema

pytestmark = pytest.mark.gpu


class Tok:
    def encode(self, p):
        return [11, 12, 13, 14, 15, 16]


IDS = dict(bos_token_id=1, eos_token_id=2, start_of_image=3, end_of_image=4)


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, ret):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    from g2vlm_b200.model import G2VLMFast
    cfg = schema.TINY
    sd = schema.init_synthetic(cfg, seed=0)
    model = G2VLMFast(cfg, sd, device=f"cuda:{rank}")
    v = schema.synthetic_views(4, 42, 518, seed=6)
    pred = model.recon_view_sharded(Tok(), dict(IDS), v)
    v0, v1 = pred["view_range"]
    full = model.recon(Tok(), dict(IDS), None, v)          # every rank also runs the whole scene alone
    torch.cuda.synchronize()
    errs = {}
    for k in ("points", "local_points", "global_points", "camera_poses"):
        a, b = pred[k].float().cpu(), full[k][:, v0:v1].float().cpu()
        errs[k] = ((a - b).abs().max() / full[k].float().abs().max().cpu()).item()
    errs["poses_all"] = ((pred["camera_poses_all"] - full["camera_poses"]).abs().max() / full["camera_poses"].abs().max()).item()
    ret[rank] = (v0, v1, errs)
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs")
def test_view_sharded_recon_matches_single_gpu():
    world = 2
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_worker, args=(world, _free_port(), ret), nprocs=world, join=True)
    assert sorted((ret[r][0], ret[r][1]) for r in range(world)) == [(0, 2), (2, 4)]
    for r in range(world):
        for k, e in ret[r][2].items():
            assert e < 5e-3, (r, k, e)   # same kernels; only the key order inside attention differs
