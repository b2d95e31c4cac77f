"""Synthetic workloads (scenes, detections, network weights) shared by bench.py, tools/ and tests/."""
