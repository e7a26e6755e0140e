"""Developer helper (not product, not bench): build experimental variants of libdpi_b200.so with extra nvcc defines
and time the bf16 integral kernel with each.  Usage on the build box:  python tools/variants.py build NAME=-DFLAG ...
On the GPU box:  python tools/variants.py run [points] [M]"""
import os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "deeppicarditeration_b200")
VDIR = os.path.join(PKG, "variants")
sys.path.insert(0, ROOT)
if sys.argv[1] == "build":
    from deeppicarditeration_b200 import build as B
    os.makedirs(VDIR, exist_ok=True)
    for spec in sys.argv[2:]:
        name, _, flags = spec.partition("=")
        objs = []
        for src in B.SOURCES:
            obj = os.path.join(VDIR, f"{name}_{src.replace('.cu', '.o')}")
            if src != "dpi_mlp_tc.cu":
                obj = os.path.join(PKG, "build", src.replace(".cu", ".o"))  # unchanged translation units
            else:
                subprocess.check_call(["/usr/local/cuda/bin/nvcc", *[f for f in B.NVCC_FLAGS if f not in ("-Xptxas", "-v")],
                                       *flags.split(), "-c", os.path.join(B.CSRC, src), "-o", obj])
            objs.append(obj)
        subprocess.check_call(["/usr/local/cuda/bin/nvcc", "-shared", "-o", os.path.join(VDIR, f"lib_{name}.so"), *objs, "-lcudart"])
        print("built", name)
else:
    n = sys.argv[2] if len(sys.argv) > 2 else "2048"
    M = sys.argv[3] if len(sys.argv) > 3 else "4096"
    libs = sorted(f for f in os.listdir(VDIR) if f.endswith(".so"))
    for lib in libs:
        env = dict(os.environ, DPI_B200_LIB=os.path.join(VDIR, lib))
        r = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "quick_bench.py"), "f16", n, M], env=env,
                           capture_output=True, text=True, timeout=300)
        print(lib, (r.stdout.strip().splitlines() or [r.stderr[-300:]])[-1])
        r = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "acc16_check.py")], env=env, capture_output=True,
                           text=True, timeout=300)
        print("   ", " | ".join(r.stdout.strip().splitlines()[-2:]))
