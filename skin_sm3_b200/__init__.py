"""skin_sm3_b200 -- B200-native (sm_100a) implementation of SM3's contrastive hot path.

Public surface (all CUDA-only, backed by libsm3_b200.so; no CPU fallback):
    cal_logits, fused_infonce, fused_infonce_multi, cluster_memory, spherical_kmeans, l2_normalize, multihead_ce, bce_with_logits, sim_topk, proto_heads, HostInfoNCE, HostInfoNCEPipeline, GraphedInfoNCE
    dropin.install()  -- put the shadow ``src.models.simclr`` in front of the reference's on sys.path
"""
from ._lib import ALGO_AUTO, ALGO_SIMT, ALGO_TC, LIB_PATH, build, lib  # noqa: F401
from .functional import (GraphedInfoNCE, cluster_memory, spherical_kmeans, HostInfoNCE, HostInfoNCEPipeline, NUM_CLASSES, bce_with_logits, cal_logits, core, fused_infonce, fused_infonce_multi,  # noqa: F401
                         gather_global_order, knn_predict, l2_normalize, multihead_ce, pick_precision, reload_env, sim_topk, TailSpec, tail_cal_logits, tail_supported,
                         proto_heads, proto_heads_supported, mlc_model_forward)

__version__ = "0.1.0"
