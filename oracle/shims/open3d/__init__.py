"""Empty open3d stub: only imported, never called, on the hot path (mv3d/eval/metricfunctions.py:2)."""
