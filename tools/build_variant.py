"""Build a debug variant of liby11_b200.so next to the product library: `python tools/build_variant.py NAME -DFLAG ...`
-> yolo_infer_b200/_lib/liby11_NAME.so (use with Y11_LIB=...).  Only the translation units that see the flags are recompiled
(conv_tc.cu by default; `--src a.cu,b.cu` to choose); the other objects come from the product build."""
import subprocess
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from yolo_infer_b200 import build as B  # noqa: E402

name = sys.argv[1]
args = sys.argv[2:]
srcs = ["conv_tc.cu"]
if args and args[0] == "--src":
    srcs = args[1].split(",")
    args = args[2:]
B.build()
objdir = B.LIBDIR / f"obj_{name}"
objdir.mkdir(exist_ok=True)
objs = []
for s in B.SOURCES:
    if s in srcs:
        o = objdir / s.replace(".cu", ".o")
        subprocess.run([B._nvcc(), *B.NVCC_FLAGS, *args, "-c", str(B.CSRC / s), "-o", str(o)], check=True)
    else:
        o = B.LIBDIR / "obj" / s.replace(".cu", ".o")
    objs.append(str(o))
out = B.LIBDIR / f"liby11_{name}.so"
subprocess.run([B._nvcc(), "-shared", "-o", str(out), *objs, "-gencode", "arch=compute_100a,code=sm_100a", "-ldl"], check=True)
print(out)
