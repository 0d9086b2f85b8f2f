"""Drop-in for the reference's model wrapper (``core.model.YOLO11Model`` / ``YOLO11Factory``,
/root/reference/core/model.py:29-323) on the detect path: same constructor, attributes, ``predict`` forwarding,
``get_model_info`` keys and ``benchmark`` keys - with ``self.model`` being the B200 engine
(yolo_infer_b200.engine.YOLO) instead of ``ultralytics.YOLO``.

Behavioural notes for a maintainer switching over (INTEGRATION.md has the two-line patch):
  * only task='detect' is implemented; other tasks raise NotImplementedError at construction
    (the reference validates the same way, then fails later when ultralytics downloads weights);
  * device: CUDA (sm_100) only; 'cpu' / 'mps' raise - there is no CPU fallback;
  * weights: ``yolo11{n,s,m,l,x}.yaml`` builds a random-init model; ``*.pt`` must be a plain state_dict
    (ultralytics key names) saved by ``YOLO11Model.save``; ``model_path=None`` resolves ``yolo11{size}.pt`` in
    the working directory like the reference (core/model.py:105-107) and falls back to random init offline.
"""
from __future__ import annotations

import logging
import time
from pathlib import Path
from typing import Any, Dict, List, Optional, Union

import torch

from .engine import YOLO

logger = logging.getLogger(__name__)


class YOLO11Model:
    SUPPORTED_TASKS = {"detect": "yolo11n.pt", "segment": "yolo11n-seg.pt", "classify": "yolo11n-cls.pt",
                       "pose": "yolo11n-pose.pt", "obb": "yolo11n-obb.pt"}
    SUPPORTED_SIZES = ["n", "s", "m", "l", "x"]

    def __init__(self, model_path: Optional[Union[str, Path]] = None, task: str = "detect", size: str = "n",
                 device: Optional[str] = None, verbose: bool = True):
        self.task, self.size, self.verbose, self.model_path = task, size, verbose, model_path
        self.device = device or self._get_default_device()
        self.optimization_history: List[Dict[str, Any]] = []
        self._validate_inputs()
        self.model = self._load_model()
        self.original_model = None
        if verbose:
            logger.info("YOLO11 (B200 path) ready: task=%s size=%s device=%s", task, self.model.scale, self.device)

    @staticmethod
    def _get_default_device() -> str:
        if not torch.cuda.is_available():
            raise RuntimeError("yolo_infer_b200 needs an sm_100 CUDA device; there is no CPU fallback")
        return "cuda"

    def _validate_inputs(self) -> None:
        if self.task not in self.SUPPORTED_TASKS:
            raise ValueError(f"Unsupported task: {self.task}. Supported: {list(self.SUPPORTED_TASKS)}")
        if self.size not in self.SUPPORTED_SIZES:
            raise ValueError(f"Unsupported size: {self.size}. Supported: {self.SUPPORTED_SIZES}")
        if self.task != "detect":
            raise NotImplementedError(f"task={self.task!r}: only the detect path is built for B200 (SURVEY.md section 8)")

    def _resolve_path(self) -> str:
        if self.model_path:
            return str(self.model_path)
        stem = Path(self.SUPPORTED_TASKS[self.task]).stem  # 'yolo11n'
        return f"{stem[:-1]}{self.size}.pt"

    def _load_model(self) -> YOLO:
        engine = YOLO(self._resolve_path(), task=self.task)
        if self.device:
            engine.to(self.device)
        self.size = engine.scale if self.model_path else self.size
        return engine

    def predict(self, source, **kwargs):
        return self.model.predict(source, **kwargs)

    def train(self, *a, **k):
        return self.model.train(*a, **k)

    def val(self, data=None, **kwargs):
        return self.model.val(data=data, **kwargs)

    def export(self, *a, **k):
        return self.model.export(*a, **k)

    def save(self, path: Union[str, Path]) -> None:
        self.model.save(path)
        logger.info("Model saved to: %s", path)

    def load(self, path: Union[str, Path]) -> None:
        self.model = YOLO(str(path), task=self.task)
        if self.device:
            self.model.to(self.device)
        logger.info("Model loaded from: %s", path)

    def get_model_info(self) -> Dict[str, Any]:
        info: Dict[str, Any] = {"task": self.task, "size": self.size, "device": self.device, "model_path": self.model_path,
                                "optimization_history": self.optimization_history}
        params = list(self.model.model.parameters())
        total = sum(p.numel() for p in params)
        info.update(total_parameters=total, trainable_parameters=sum(p.numel() for p in params if p.requires_grad),
                    model_size_mb=total * 4 / (1024 * 1024))
        return info

    def benchmark(self, data_source, num_runs: int = 100, warmup_runs: int = 10) -> Dict[str, float]:
        for _ in range(warmup_runs):
            self.predict(data_source, verbose=False)
        times = []
        for _ in range(num_runs):
            t0 = time.time()
            self.predict(data_source, verbose=False)
            times.append(time.time() - t0)
        mean = sum(times) / len(times)
        return {"avg_inference_time": mean, "min_inference_time": min(times), "max_inference_time": max(times), "fps": 1.0 / mean}

    def __repr__(self) -> str:
        return (f"YOLO11Model(task={self.task}, size={self.size}, device={self.device}, "
                f"optimized={len(self.optimization_history) > 0})")


class YOLO11Factory:
    @staticmethod
    def create_detector(size: str = "n", **kwargs) -> YOLO11Model:
        return YOLO11Model(task="detect", size=size, **kwargs)

    @staticmethod
    def _unsupported(task: str):
        raise NotImplementedError(f"{task}: only the detect path is built for B200 (SURVEY.md section 8)")

    @staticmethod
    def create_segmenter(size: str = "n", **kwargs):
        YOLO11Factory._unsupported("segment")

    @staticmethod
    def create_classifier(size: str = "n", **kwargs):
        YOLO11Factory._unsupported("classify")

    @staticmethod
    def create_pose_estimator(size: str = "n", **kwargs):
        YOLO11Factory._unsupported("pose")

    @staticmethod
    def create_obb_detector(size: str = "n", **kwargs):
        YOLO11Factory._unsupported("obb")
