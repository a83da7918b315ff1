"""fastkv_b200 — B200-native TSP + KV-cache-compression hot path behind FastKV's plugin surface.

Public surface (same names as the reference, baselines/monkeypatch.py and baselines/*/utils.py):
``replace_llama``, ``replace_mistral``, ``set_model``, ``compress_fastkv``, ``init_fastkv``,
``FastKVCluster``, ``SnapKVCluster``, ``find_context`` / ``standard_dis_index``.
"""
from .cluster import FastKVCluster, SnapKVCluster, find_context, standard_dis_index  # noqa: F401

__all__ = ["FastKVCluster", "SnapKVCluster", "find_context", "standard_dis_index"]
