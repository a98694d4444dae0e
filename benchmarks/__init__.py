"""Benchmark harnesses driven by bench.py (not part of the product package): the layer workloads on one GPU and on
destination-row shards, their parity check against a single-GPU run, and the epoch workloads of the reference's
training scripts."""
