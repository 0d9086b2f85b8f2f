"""yolo_infer_b200: B200-native (sm_100a) YOLO11 detection inference path behind the reference's
``YOLO11Model(...).predict`` surface.  See DESIGN.md / INTEGRATION.md.

(The task names the package ``yolo-infer_b200``; a hyphen cannot be imported, so the directory is
``yolo_infer_b200``.)"""
from .model import YOLO11Factory, YOLO11Model  # noqa: F401
from .engine import YOLO  # noqa: F401
from .results import Boxes, Results  # noqa: F401

__all__ = ["YOLO11Model", "YOLO11Factory", "YOLO", "Results", "Boxes"]
