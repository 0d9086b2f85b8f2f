"""The reference's speed-benchmark harness over the B200 path.

Mirror of `SpeedBenchmark` (/root/reference/benchmarks/speed_benchmark.py:29-422): same constructor, same methods, same result
keys - `_benchmark_inference` (:307-350: avg/min/max/std_inference_time, fps, throughput), `benchmark_model_sizes` (:61-122),
`benchmark_throughput` (:211-305: total_inferences, duration_seconds, images_per_second, resource_usage, ...), summary,
JSON / text reports - so that `main.py benchmark` keeps producing the files it produced.  The reference's harness imports
`get_device_info` and `ResourceMonitor` from utils/helpers.py (:21-66, :715-834), which need GPUtil (not installable
here): the equivalents below read the same quantities through NVML (`pynvml`, already a dependency of torch's CUDA build)
and keep the keys the reports print.

Input: exactly what the reference feeds - `torch.randn(B, 3, S, S)[.cuda()]` straight into `model.predict(x, verbose=False)`
(:100-102, :238-240, :326-333).  On the B200 path such a tensor runs as ONE CUDA-graph replay (device-side max -> /255 rule,
forward, decode, NMS) plus one result D2H, which is also the call's synchronisation point, so the wall-clock timing the
reference takes around `predict` measures completed work.

`sizes` / `model_size` entries may be a scale letter ('n', ..., as in the reference: resolves yolo11{size}.pt, which must
exist - there is no download) or a path to a `.pt` state_dict / `yolo11{scale}.yaml`.
"""
from __future__ import annotations

import json
import logging
import platform
import statistics
import threading
import time
from pathlib import Path
from typing import Any, Dict, List, Optional, Union

import torch

from .model import YOLO11Model

logger = logging.getLogger(__name__)


def _nvml():
    try:
        import pynvml
        pynvml.nvmlInit()
        return pynvml
    except Exception:
        return None


def _gpu_rows(nv) -> List[Dict[str, Any]]:
    rows = []
    if nv is None:
        return rows
    for i in range(nv.nvmlDeviceGetCount()):
        h = nv.nvmlDeviceGetHandleByIndex(i)
        mem = nv.nvmlDeviceGetMemoryInfo(h)
        try:
            util = nv.nvmlDeviceGetUtilizationRates(h).gpu
        except Exception:
            util = 0
        try:
            temp = nv.nvmlDeviceGetTemperature(h, nv.NVML_TEMPERATURE_GPU)
        except Exception:
            temp = None
        name = nv.nvmlDeviceGetName(h)
        rows.append({"id": i, "name": name.decode() if isinstance(name, bytes) else name, "memory_total_mb": mem.total / 2 ** 20,
                     "memory_used_mb": mem.used / 2 ** 20, "memory_free_mb": mem.free / 2 ** 20, "temperature": temp,
                     "load": util / 100.0})
    return rows


def get_device_info() -> Dict[str, Any]:
    """Same keys as the reference's utils/helpers.py:21-66 (GPUtil replaced by NVML)."""
    import psutil
    vm = psutil.virtual_memory()
    info = {"platform": platform.platform(), "processor": platform.processor(), "architecture": platform.architecture()[0],
            "python_version": platform.python_version(), "cpu_count": psutil.cpu_count(),
            "memory_total_gb": round(vm.total / 1024 ** 3, 2), "memory_available_gb": round(vm.available / 1024 ** 3, 2),
            "memory_total": vm.total,
            "torch_version": torch.__version__, "cuda_available": torch.cuda.is_available(), "mps_available": False}
    if torch.cuda.is_available():
        info["cuda_version"] = torch.version.cuda
        info["cudnn_version"] = torch.backends.cudnn.version()
        info["cuda_device_count"] = torch.cuda.device_count()
        try:
            info["gpus"] = _gpu_rows(_nvml())
        except Exception as e:   # same behaviour as the reference: report, keep going
            logger.warning("Could not get GPU information: %s", e)
            info["gpus"] = []
    return info


class ResourceMonitor:
    """utils/helpers.py:715-834 with NVML instead of GPUtil: a daemon thread samples CPU / memory / GPU every `interval` s."""

    def __init__(self, interval: float = 1.0):
        self.interval, self.monitoring, self.history = interval, False, []

    def start_monitoring(self):
        import psutil
        self.monitoring, self.history = True, []
        nv = _nvml()

        def loop():
            while self.monitoring:
                try:
                    cpu = psutil.cpu_percent(interval=0.1)
                    mem = psutil.virtual_memory()
                    gpu = [{"id": g["id"], "load": g["load"] * 100, "memory_used": g["memory_used_mb"], "memory_total": g["memory_total_mb"],
                            "temperature": g["temperature"]} for g in _gpu_rows(nv)]
                    self.history.append({"timestamp": time.time(), "cpu_percent": cpu, "memory_percent": mem.percent,
                                         "memory_used": mem.used, "memory_total": mem.total, "gpu_usage": gpu})
                    if len(self.history) > 1000:
                        self.history.pop(0)
                    time.sleep(self.interval)
                except Exception as e:
                    logger.error("Error in resource monitoring: %s", e)
                    break

        self.monitor_thread = threading.Thread(target=loop, daemon=True)
        self.monitor_thread.start()

    def stop_monitoring(self):
        self.monitoring = False
        if hasattr(self, "monitor_thread"):
            self.monitor_thread.join(timeout=2)

    def get_current_usage(self) -> Dict[str, Any]:
        return self.history[-1] if self.history else {}

    def get_average_usage(self, last_n: Optional[int] = None) -> Dict[str, float]:
        data = (self.history[-last_n:] if last_n else self.history) if self.history else []
        if not data:
            return {}
        out = {"avg_cpu_percent": sum(d["cpu_percent"] for d in data) / len(data),
               "avg_memory_percent": sum(d["memory_percent"] for d in data) / len(data)}
        for i, _ in enumerate(data[0].get("gpu_usage") or []):
            loads = [d["gpu_usage"][i]["load"] for d in data if i < len(d["gpu_usage"])]
            if loads:
                out[f"avg_gpu_{i}_load"] = sum(loads) / len(loads)
        return out

    def save_history(self, file_path: Union[str, Path]):
        p = Path(file_path)
        p.parent.mkdir(parents=True, exist_ok=True)
        p.write_text(json.dumps(self.history, indent=2))


class SpeedBenchmark:
    def __init__(self, output_dir: str = "benchmark_results", warmup_runs: int = 10, benchmark_runs: int = 100):
        self.output_dir = Path(output_dir)
        self.output_dir.mkdir(parents=True, exist_ok=True)
        self.warmup_runs, self.benchmark_runs = warmup_runs, benchmark_runs
        self.system_info = get_device_info()

    @staticmethod
    def _model(task: str, size: str) -> YOLO11Model:
        if str(size).endswith((".pt", ".pth", ".yaml", ".yml")):
            return YOLO11Model(model_path=size, task=task, verbose=False)
        return YOLO11Model(task=task, size=size, verbose=False)

    @staticmethod
    def _input(batch_size: int, image_size: int) -> torch.Tensor:
        x = torch.randn(batch_size, 3, image_size, image_size)
        return x.cuda() if torch.cuda.is_available() else x

    def benchmark_model_sizes(self, task: str = "detect", sizes: List[str] = ("n", "s", "m", "l", "x"),
                              image_sizes: List[int] = (320, 640, 1280), batch_sizes: List[int] = (1, 4, 8, 16)) -> Dict[str, Any]:
        results: Dict[str, Any] = {"task": task, "system_info": self.system_info, "configurations": [], "summary": {}}
        for size in sizes:
            model = self._model(task, size)
            for img_size in image_sizes:
                for batch_size in batch_sizes:
                    metrics = self._benchmark_inference(model, self._input(batch_size, img_size))
                    results["configurations"].append({"model_size": size, "image_size": img_size, "batch_size": batch_size, **metrics})
        results["summary"] = self._calculate_summary(results["configurations"])
        self._save_results(results, "model_sizes_benchmark.json")
        return results

    def benchmark_throughput(self, model_size: str = "n", task: str = "detect", duration_seconds: int = 60, image_size: int = 640,
                             batch_size: int = 1) -> Dict[str, Any]:
        model = self._model(task, model_size)
        test_input = self._input(batch_size, image_size)
        monitor = ResourceMonitor(interval=1.0)
        monitor.start_monitoring()
        for _ in range(self.warmup_runs):
            model.predict(test_input, verbose=False)
        start = time.time()
        count, times = 0, []
        while time.time() - start < duration_seconds:
            t0 = time.time()
            model.predict(test_input, verbose=False)
            times.append(time.time() - t0)
            count += 1
        total = time.time() - start
        monitor.stop_monitoring()
        fps = count / total
        results = {"model_size": model_size, "task": task, "image_size": image_size, "batch_size": batch_size,
                   "duration_seconds": total, "total_inferences": count, "avg_inference_time": statistics.mean(times), "fps": fps,
                   "throughput": fps, "images_per_second": fps * batch_size, "resource_usage": monitor.get_average_usage(),
                   "system_info": self.system_info}
        monitor.save_history(self.output_dir / "resource_history.json")
        self._save_results(results, "throughput_benchmark.json")
        return results

    def _benchmark_inference(self, model: YOLO11Model, test_input: torch.Tensor) -> Dict[str, float]:
        model.model.eval()
        with torch.no_grad():
            for _ in range(self.warmup_runs):
                model.predict(test_input, verbose=False)
        times = []
        with torch.no_grad():
            for _ in range(self.benchmark_runs):
                t0 = time.time()
                model.predict(test_input, verbose=False)
                times.append(time.time() - t0)
        avg = statistics.mean(times)
        return {"avg_inference_time": avg, "min_inference_time": min(times), "max_inference_time": max(times),
                "std_inference_time": statistics.stdev(times) if len(times) > 1 else 0.0, "fps": 1.0 / avg,
                "throughput": test_input.shape[0] / avg}

    @staticmethod
    def _calculate_summary(configurations: List[Dict]) -> Dict[str, Any]:
        if not configurations:
            return {}
        fps = [c["fps"] for c in configurations]
        thr = [c["throughput"] for c in configurations]
        return {"best_fps": max(fps), "worst_fps": min(fps), "avg_fps": statistics.mean(fps), "best_throughput": max(thr),
                "worst_throughput": min(thr), "avg_throughput": statistics.mean(thr), "total_configurations": len(configurations)}

    def _save_results(self, results: Dict[str, Any], filename: str):
        (self.output_dir / filename).write_text(json.dumps(results, indent=2, default=str))

    def generate_report(self) -> str:
        path = self.output_dir / "benchmark_report.txt"
        with open(path, "w") as f:
            f.write("YOLO11 Speed Benchmark Report\n" + "=" * 50 + "\n\nSystem Information:\n" + "-" * 20 + "\n")
            for k, v in self.system_info.items():
                f.write(f"{k}: {v}\n")
            f.write("\n")
            for rf in sorted(self.output_dir.glob("*.json")):
                if rf.name == "resource_history.json":
                    continue
                try:
                    data = json.loads(rf.read_text())
                except Exception as e:
                    logger.warning("Could not process %s: %s", rf, e)
                    continue
                f.write(f"Results from {rf.name}:\n" + "-" * 30 + "\n")
                for k, v in (data.get("summary") or {}).items():
                    f.write(f"{k}: {v}\n")
                f.write("\n")
        return str(path)
