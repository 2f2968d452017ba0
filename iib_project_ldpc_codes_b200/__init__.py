"""B200-native (sm_100a) implementation of the BEC message-passing Monte-Carlo hot path of
roryhighnam/iib_project_ldpc_codes: bit-sliced flooding erasure decoding on random (dv,dc)-regular
Tanner graphs behind the reference's own C ABI (message_passing.c:7) plus batched / fused entry points.

GPU-only: the CUDA library must be built (``__graft_entry__.build()``); there is no CPU fallback.
"""
from .api import (Graph, McResult, counters_len, draw_patterns, generate_graph, generate_graph_indexed, mc_run_ensemble,  # noqa: F401
                  message_passing)
from ._lib import LdpcB200Error, LIB_PATH  # noqa: F401

__all__ = ["Graph", "McResult", "counters_len", "draw_patterns", "generate_graph", "generate_graph_indexed", "mc_run_ensemble", "message_passing",
           "LdpcB200Error", "LIB_PATH"]
