"""ambersim_b200: B200-native batched rigid-body rollout engine behind ambersim's Python APIs."""
from pathlib import Path

ROOT = str(Path(__file__).resolve().parent)  # reference: ambersim/__init__.py:8
