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


def _worker(rank, world, port, ret, n_views=4):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    from g2vlm_b200.model import G2VLMFast
    cfg = schema.TINY
    sd = schema.init_synthetic(cfg, seed=0)
    model = G2VLMFast(cfg, sd, device=f"cuda:{rank}")
    v = schema.synthetic_views(n_views, 42, 518, seed=6)
    full = model.recon(Tok(), dict(IDS), None, v)          # every rank also runs the whole scene alone
    full = {k: full[k].clone() for k in ("points", "local_points", "global_points", "camera_poses")}
    errs = {}
    modes = ["peer", "overlap"] + (["allgather"] if n_views % world == 0 else [])   # the all-gather needs equal shards
    for mode in modes:
        model.sp_mode = mode
        pred = model.recon_view_sharded(Tok(), dict(IDS), v)
        if mode == "peer":
            errs["peer.symmetric_memory_available"] = 0.0 if any(x is not None for x in model._sp_sym.values()) else 1.0
        v0, v1 = pred["view_range"]
        torch.cuda.synchronize()
        for k in ("points", "local_points", "global_points", "camera_poses"):
            a, b = pred[k].float().cpu(), full[k][:, v0:v1].float().cpu()
            errs[f"{mode}.{k}"] = ((a - b).abs().max() / full[k].float().abs().max().cpu()).item()
        errs[f"{mode}.poses_all"] = ((pred["camera_poses_all"] - full["camera_poses"]).abs().max()
                                     / full["camera_poses"].abs().max()).item()
    ret[rank] = (v0, v1, errs)
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs")
@pytest.mark.parametrize("n_views,ranges", [(4, [(0, 2), (2, 4)]), (5, [(0, 3), (3, 5)])])
def test_view_sharded_recon_matches_single_gpu(n_views, ranges):
    world = 2
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_worker, args=(world, _free_port(), ret, n_views), nprocs=world, join=True)
    assert sorted((ret[r][0], ret[r][1]) for r in range(world)) == ranges
    for r in range(world):
        for k, e in ret[r][2].items():
            # same kernels; the key order inside attention differs (all-gather) / the two partials are rounded to
            # bf16 before the log-sum-exp merge (overlap)
            assert e < (5e-3 if k.startswith("allgather") else 1e-2), (r, k, e)


def _sp_exchange_worker(rank, world, port, ret):
    """g2vlm_sp_kv_exchange with a RAW ncclComm_t (created here through the NCCL C API, as a non-Python host would)."""
    import ctypes
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    torch.cuda.set_device(rank)
    dist.init_process_group("gloo", rank=rank, world_size=world)       # only to hand the unique id around
    from g2vlm_b200 import _lib
    lib = _lib.load()
    nccl = ctypes.CDLL("libnccl.so.2")

    class UniqueId(ctypes.Structure):
        _fields_ = [("internal", ctypes.c_char * 128)]

    uid = UniqueId()
    if rank == 0:
        assert nccl.ncclGetUniqueId(ctypes.byref(uid)) == 0
    t = torch.tensor(list(bytes(uid)), dtype=torch.uint8)
    dist.broadcast(t, 0)
    uid = UniqueId.from_buffer_copy(bytes(t.tolist()))
    comm = ctypes.c_void_p()
    nccl.ncclCommInitRank.argtypes = [ctypes.POINTER(ctypes.c_void_p), ctypes.c_int, UniqueId, ctypes.c_int]
    assert nccl.ncclCommInitRank(ctypes.byref(comm), world, uid, rank) == 0
    rows, kvw = [3, 5], 512                                              # uneven shards
    send = torch.full((rows[rank], kvw), float(rank + 1), dtype=torch.bfloat16, device="cuda")
    send += torch.arange(kvw, device="cuda").to(torch.bfloat16)[None, :] / 1024
    remote = torch.zeros(sum(rows) - rows[rank], kvw, dtype=torch.bfloat16, device="cuda")
    rr = (ctypes.c_int64 * world)(*rows)
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    rc = lib.g2vlm_sp_kv_exchange(comm, ctypes.c_int32(rank), ctypes.c_int32(world), rr, ctypes.c_int64(kvw),
                                  ctypes.c_void_p(send.data_ptr()), ctypes.c_void_p(remote.data_ptr()),
                                  ctypes.c_void_p(side.cuda_stream))
    side.synchronize()
    peer = 1 - rank
    want = torch.full((rows[peer], kvw), float(peer + 1), dtype=torch.bfloat16, device="cuda")
    want += torch.arange(kvw, device="cuda").to(torch.bfloat16)[None, :] / 1024
    ret[rank] = (rc, bool(torch.equal(remote, want)), lib.g2vlm_last_error().decode())
    nccl.ncclCommDestroy.argtypes = [ctypes.c_void_p]
    nccl.ncclCommDestroy(comm)
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs")
def test_sp_kv_exchange_with_a_raw_nccl_communicator():
    world = 2
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_sp_exchange_worker, args=(world, _free_port(), ret), nprocs=world, join=True)
    for r in range(world):
        rc, same, err = ret[r]
        assert rc == 0 and same, (r, rc, err)
