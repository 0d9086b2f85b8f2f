"""Debug (needs a -DY11_TRACE build): per-role timelines of selected YOLO11s layers at batch 64."""
import sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent))
from trace_conv import run
run(64, 80, 80, 256, 256, 3, 2, label="11s model.5 (256->256 s2)")
run(64, 160, 160, 64, 128, 3, 2, label="11s model.3 (64->128 s2)")
run(64, 20, 20, 64, 64, 3, 1, label="11s model.8.m.0.m.0.cv1 (64->64 20x20)")
run(64, 80, 80, 128, 64, 3, 1, label="11s model.23.cv2.0.0 (128->64 80x80)")
run(64, 40, 40, 256, 256, 1, 1, label="11s 1x1 256->256 40x40")
