"""Benchmark baselines (never imported by the product package gs_dynamics_b200)."""
