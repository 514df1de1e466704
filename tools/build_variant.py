"""Build an A/B variant of the library: recompile the named sources with extra -D flags and link
them with the other objects of the regular build.

    python tools/build_variant.py NAME file.cu[,file2.cu] -DX=1 [-DY=2 ...]
-> wave_mamba_b200/_variants/lib_NAME.so ; select it with WM_B200_LIB=<that path>."""
import os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from wave_mamba_b200 import build as b
name, files, flags = sys.argv[1], sys.argv[2].split(","), sys.argv[3:]
b.build()
vdir = os.path.join(b.HERE, "_variants"); os.makedirs(vdir, exist_ok=True)
env = dict(os.environ); env.pop("CC", None); env.pop("CXX", None)
objs = []
for src in b.SOURCES:
    o = os.path.join(b.OBJ, src.replace(".cu", ".o"))
    if src in files:
        o = os.path.join(vdir, f"{name}_{src.replace('.cu', '.o')}")
        subprocess.run([b._nvcc(), "-c", os.path.join(b.CSRC, src), "-o", o] + b.ARCH + b.COMMON + b.EXTRA.get(src, []) + flags, check=True, env=env)
    objs.append(o)
out = os.path.join(vdir, f"lib_{name}.so")
subprocess.run([b._nvcc(), "-shared", "-o", out] + objs + b.ARCH + ["-Xcompiler", "-fPIC"], check=True, env=env)
print(out)
