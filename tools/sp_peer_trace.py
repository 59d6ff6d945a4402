"""Where does the time of the view-sharded peer exchange go?  Per layer, relative to the moment the rank's K|V rows are
written: barrier passed, pulls done, local-key attention done.  torchrun --nproc-per-node N tools/sp_peer_trace.py [views]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist
from g2vlm_b200 import schema
from g2vlm_b200.model import G2VLMFast

rank, world, lr = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(lr)
dist.init_process_group("nccl", device_id=torch.device("cuda", lr))
views = int(sys.argv[1]) if len(sys.argv) > 1 else 64

class Tok:
    def encode(self, p): return [11, 12, 13, 14, 15, 16]
IDS = dict(bos_token_id=1, eos_token_id=2, start_of_image=3, end_of_image=4)
cfg = schema.FULL
m = G2VLMFast(cfg, schema.init_synthetic(cfg, seed=0, embed_rows=32, device="cuda"))
v = (schema.synthetic_views(views, 518, 518, seed=1) * 255).round() / 255.0
for margin in (2, 8):
    m.sp_peer_margin = margin
    m.recon_view_sharded(Tok(), dict(IDS), v)
    m.sp_trace = []
    m.recon_view_sharded(Tok(), dict(IDS), v)
    torch.cuda.synchronize()
    tr, m.sp_trace = m.sp_trace, None
    mean = lambda a, b: sum(t[a].elapsed_time(t[b]) for t in tr) / len(tr)
    print(f"rank {rank} margin {margin}: per layer, ms after the K|V rows were written: barrier passed {mean('written', 'barrier'):.3f}, "
          f"pulls done {mean('written', 'done'):.3f}, local attention done {mean('written', 'local_end'):.3f}", flush=True)
dist.barrier()
dist.destroy_process_group()
