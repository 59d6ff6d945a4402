"""CPU checks of the boundary: the C-ABI library builds/loads and exports every symbol declared in
include/g2vlm_b200.h; argument validation returns error codes (no compute without a GPU); the host-side
index construction matches the reference golden bit for bit."""
import ctypes
import os

import pytest
import torch

from g2vlm_b200 import _lib, host_prep, ops, schema

GOLDEN = os.path.join(os.path.dirname(__file__), "golden")


@pytest.fixture(scope="module")
def lib():
    _lib.build()
    return _lib.load()


def test_library_exports_every_declared_symbol(lib):
    names = _lib.declared_symbols()
    assert len(names) >= 17 and "g2vlm_gemm_bf16" in names and "g2vlm_attention" in names
    for n in names:
        assert hasattr(lib, n), n
    assert lib.g2vlm_abi_version() == _lib._header_abi_version() == 4


def test_invalid_arguments_are_rejected_without_launching(lib):
    assert lib.g2vlm_gemm_bf16(None, None) == 1
    assert b"null args" in lib.g2vlm_last_error()
    a = ops.GemmArgs()
    a.A, a.B, a.out = 16, 16, 16
    a.N, a.K, a.n_groups, a.lda, a.ldb, a.ldo = 256, 60, 1, 64, 64, 256   # K not a multiple of 8
    assert lib.g2vlm_gemm_bf16(ctypes.byref(a), None) == 1
    assert b"multiples of 8" in lib.g2vlm_last_error()
    at = ops.AttnArgs()
    at.q = at.k = at.v = at.out = 16
    at.head_dim, at.num_q_heads, at.num_kv_heads = 96, 16, 16              # 96 must be padded to 128
    assert lib.g2vlm_attention(ctypes.byref(at), None) == 1
    assert b"head_dim" in lib.g2vlm_last_error()
    assert lib.g2vlm_gather_rows(ctypes.c_void_p(16), ctypes.c_int64(10), ctypes.c_void_p(16), ctypes.c_int64(16),
                                 None, ctypes.c_int64(1), ctypes.c_int64(16), ctypes.c_int32(0), None) == 1


def test_attention_trace_build_compiles(lib):
    """The -DG2_ATTN_TRACE debug build (tools/attn_trace.py: clock-stamped timeline of the attention kernel) must keep
    compiling next to the product build, and only IT carries the trace export."""
    import subprocess
    import sys
    tool = os.path.join(os.path.dirname(os.path.dirname(__file__)), "tools", "attn_trace.py")
    subprocess.run([sys.executable, tool, "--build"], check=True, capture_output=True, timeout=600)
    trace = ctypes.CDLL(str(_lib.PKG_DIR / "libg2vlm_b200_trace.so"))
    assert hasattr(trace, "g2vlm_debug_attn_trace") and not hasattr(lib, "g2vlm_debug_attn_trace")
    assert all(hasattr(trace, n) for n in _lib.declared_symbols())


def test_ops_refuse_cpu_tensors():
    x = torch.zeros(4, 64, dtype=torch.bfloat16)
    with pytest.raises(ops.G2Error, match="CUDA tensor"):
        ops.gemm(x, x, x, epilogue=ops.EPI_STORE_BF16)


def test_model_refuses_to_run_without_gpu():
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from g2vlm_b200.model import G2VLMFast
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        G2VLMFast(schema.TINY, {})


class Tok:
    def encode(self, p):
        return [11, 12, 13, 14, 15, 16]


IDS = dict(bos_token_id=1, eos_token_id=2, start_of_image=3, end_of_image=4)


@pytest.mark.parametrize("name", ["a", "b"])
def test_host_prep_matches_reference_golden_exactly(name):
    g = torch.load(os.path.join(GOLDEN, f"recon_tiny_{name}.pt"))
    c = g["case"]
    gi, newlens, new_rope = host_prep.prepare_prompts_addbos([0], [0], ["Reconstruct the 3D scene."], Tok(), IDS)
    for k in ("packed_text_ids", "packed_text_position_ids", "packed_text_indexes", "text_token_lens"):
        assert gi[k].dtype == g["text." + k].dtype and torch.equal(gi[k], g["text." + k]), k
    v = schema.synthetic_views(c["n"], c["h"], c["w"], seed=c["seed"])
    gd, nl, nr = host_prep.prepare_dino_images_pi3(newlens, new_rope, v, IDS)
    for k in ("packed_text_ids", "packed_text_indexes", "dino_token_seqlens", "packed_dino_token_indexes",
              "packed_position_ids", "packed_seqlens", "packed_indexes", "packed_key_value_indexes", "key_values_lens"):
        assert gd[k].dtype == g["dino." + k].dtype and torch.equal(gd[k], g["dino." + k]), k
    P = (c["h"] // 14) * (c["w"] // 14)
    assert nl == [7 + c["n"] * (P + 2)]


def test_host_prep_edge_cases():
    # a single view, and a tall image (gh > gw): rope step uses max(gh, gw)
    for n, h, w in ((1, 14, 14), (2, 70, 28)):
        v = torch.rand(n, 3, h, w)
        gd, nl, nr = host_prep.prepare_dino_images_pi3([7], [7], v, IDS)
        gh, gw = h // 14, w // 14
        assert nr == [7 + n * (max(gh, gw) + 2)]
        assert gd["packed_position_ids"].shape == (3, n * (gh * gw + 2))
        both = torch.cat([gd["packed_text_indexes"], gd["packed_dino_token_indexes"]]).sort().values
        assert torch.equal(both, torch.arange(n * (gh * gw + 2)))


def test_load_and_resize14_matches_reference_loader():
    from PIL import Image
    u8 = (schema.synthetic_views(2, 100, 180, seed=9) * 255).round().to(torch.uint8)
    pil = [Image.fromarray(u8[i].permute(1, 2, 0).numpy()) for i in range(2)]
    x = host_prep.load_and_resize14(pil, 518)
    assert x.shape == (2, 3, round(100 * 518 / 180 / 14) * 14, 518)
    assert 0.0 <= float(x.min()) and float(x.max()) <= 1.0 + 1e-6
    from oracle import ref_harness
    if ref_harness.available():
        ref_harness.install_shims()
        from data.transforms_vggt import load_and_resize14 as ref_loader
        assert torch.equal(ref_loader(pil, 518), x)


def test_attention_work_table():
    w = ops.attention_work_table([0, 300, 300, 813], [0, 1000, 1000, 1513])
    assert w.dtype == torch.int32 and w.shape == (5, 8)
    assert w[:, 0].tolist() == [0, 256, 300, 556, 812]
    assert w[1].tolist() == [256, 0, 300, 0, 1000, 0, 0, 0]
