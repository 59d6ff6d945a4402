"""g2vlm_b200 — B200-native (sm_100a) implementation of G2VLM's reconstruction forward pass.

Only what the hot path needs lives here: ``csrc/`` (CUDA kernels + the C ABI declared in
``include/g2vlm_b200.h``), ``ops`` (torch-tensor wrappers over the C ABI) and the host-side mirror
of the reference's model interface (``model.G2VLMFast``: ``recon``, ``forward_cache_update_*``,
``reconstruct`` with the reference's ``state_dict`` key schema).
"""
__version__ = "0.1.0"
