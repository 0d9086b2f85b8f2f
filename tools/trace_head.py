"""Per-role timeline (tools/trace_conv.py, -DY11_TRACE build) of the Detect head's last 1x1 layers on the P3 map."""
import sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent))
sys.argv = sys.argv[:1] + ["noop"]
import trace_conv as T  # noqa: E402

T.run(64, 80, 80, 128, 80, 1, 1, label="P3 cls 1x1 128->80 (bf16 out)")
T.run(64, 80, 80, 128, 128, 1, 1, label="P3 cls tower 1x1 128->128")
T.run(64, 80, 80, 64, 64, 1, 1, label="P3 box 1x1 64->64")
