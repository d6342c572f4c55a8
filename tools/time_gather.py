"""Stand-alone K1 timing at the cfg3 item-tower shape (bench.gather_roofline): python tools/time_gather.py
The 8 launches of one pass over the batch pool are replayed as a CUDA graph (kernel time, not host launch cost)."""
import json
import sys
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import bench  # noqa: E402
import two_tower_b200 as tt  # noqa: E402

peaks = json.loads((ROOT / "MEASURED_PEAKS.json").read_text()) if (ROOT / "MEASURED_PEAKS.json").exists() else {}
tt.set_precision("bf16")
r = bench.gather_roofline(tt, torch, torch.device("cuda", 0), peaks)
print(json.dumps({k: r[k] for k in ("us_per_launch", "achieved", "frac", "algorithmic_bytes_per_launch")}))
