"""obia_b200: B200-native SLIC segmentation + per-segment zonal statistics behind
obia's Python API (drop-in for that hot path only).

Like the reference there are no package-level re-exports: import the full
module paths, e.g. `from obia_b200.segmentation.segment import segment`.
"""
__version__ = "0.1.0"
