"""caesar_yolo_b200 — B200-native (sm_100a) implementation of caesar-yolo's tiled source-finding hot path."""
__version__ = "0.1.0"
