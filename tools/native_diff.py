"""Builder tool: where do the native stage drivers and the per-op path differ (if anywhere)?"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from dataclasses import replace
import torch
from g2vlm_b200 import schema
from g2vlm_b200.model import G2VLMFast, NaiveCache

class Tok:
    def encode(self, p): return [11, 12, 13, 14, 15, 16]
IDS = dict(bos_token_id=1, eos_token_id=2, start_of_image=3, end_of_image=4)
cfg = replace(schema.FULL, num_layers=1, dino_layers=1, dec_depth=1)
m = G2VLMFast(cfg, schema.init_synthetic(cfg, seed=3, embed_rows=32, device="cuda"))
for (n, h, w) in [(2, 70, 518), (8, 294, 518), (2, 518, 518)]:
    v = (schema.synthetic_views(n, h, w, seed=4) * 255).round() / 255.0
    gi_text, nl, nr = m.prepare_prompts_addbos([0], [0], ["x"], Tok(), IDS)
    gi, _, _ = m.prepare_dino_images_pi3(nl, nr, v, None, IDS)
    res = {}
    for native in (True, False, True, False):
        m.native = native
        past, last = m.forward_cache_update_dino(NaiveCache(1), prompt=gi_text, update_past_key_values=False, **gi)
        last = last.clone()
        out = m.reconstruct(past_key_values=past, selected_hidden_states=last, **gi)
        res.setdefault(native, []).append((last, {k: out[k].clone() for k in ("points", "local_points", "global_points", "camera_poses")}))
    torch.cuda.synchronize()
    print((n, h, w), "run-to-run native last equal:", torch.equal(res[True][0][0], res[True][1][0]),
          "| per-op:", torch.equal(res[False][0][0], res[False][1][0]),
          "| native vs per-op last:", torch.equal(res[True][0][0], res[False][0][0]),
          "max abs", (res[True][0][0] - res[False][0][0]).abs().max().item())
    # heads on the SAME last hidden
    m.native = True
    a = m.reconstruct(selected_hidden_states=res[False][0][0], **gi)
    a = {k: a[k].clone() for k in ("points", "local_points", "global_points", "camera_poses")}
    m.native = False
    b = m.reconstruct(selected_hidden_states=res[False][0][0], **gi)
    print("   heads on identical input:", {k: (torch.equal(a[k], b[k]), (a[k] - b[k]).abs().max().item()) for k in a})
    # stage by stage inside the heads
    ctx = m._native_ctx()
    ws = m._nws[m._nplan[:4]]
    m.native = True
    m.reconstruct(selected_hidden_states=res[False][0][0], **gi)
    nat = {k: ctx.region(ws, k, dt, c).clone() for k, dt, c in (("rec.point_hidden", torch.bfloat16, cfg.point_dim),
           ("rec.camera_hidden", torch.float32, cfg.camera_dim), ("rec.global_hidden", torch.bfloat16, cfg.point_dim),
           ("rec.hidden", torch.float32, cfg.hidden_size))}
    m.native = False
    c = {}
    m.reconstruct(selected_hidden_states=res[False][0][0], collect=c, **gi)
    for k, kk in (("rec.point_hidden", "point_hidden"), ("rec.camera_hidden", "camera_hidden"), ("rec.global_hidden", "global_hidden")):
        a, b = nat[k].float(), c[kk].reshape(-1, c[kk].shape[-1]).float()
        print("   ", k, torch.equal(a, b), (a - b).abs().max().item())
    hid = [t for (nm, sh, dt), t in m.buf._b.items() if nm == "rec.hidden"]
    print("    rec.hidden", [torch.equal(nat["rec.hidden"], t) for t in hid if t.shape == nat["rec.hidden"].shape])
    m.native = False
    o1 = m.reconstruct(selected_hidden_states=res[False][0][0], **gi); o1 = {k: o1[k].clone() for k in ("points", "camera_poses")}
    o2 = m.reconstruct(selected_hidden_states=res[False][0][0], **gi)
    print("    per-op heads run-to-run:", {k: torch.equal(o1[k], o2[k]) for k in o1})
    break
