"""Debug (needs a -DY11_TRACE build): per-role timelines of the narrow-channel 3x3 layers of YOLO11s at batch 64."""
import sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent))
from trace_conv import run
run(64, 160, 160, 32, 16, 3, 1, in_ct=96, in_off=32, label="11s model.2.m.0.cv1 (halo, 32->16)")
run(64, 80, 80, 64, 32, 3, 1, in_ct=192, in_off=64, label="11s model.4.m.0.cv1 (halo, 64->32)")
run(64, 160, 160, 32, 32, 1, 1, label="11n model.2.cv1 (1x1 lsu 32->32)")
run(64, 160, 160, 96, 128, 1, 1, label="11s model.2.cv2 (1x1 tma 96->128)")
run(64, 80, 80, 128, 128, 1, 1, label="11s P3 tower 1x1 (tma 128->128)")
