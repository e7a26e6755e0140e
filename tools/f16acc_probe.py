"""Developer probe: tcgen05.mma with fp16 accumulators (c_format = 0) and tcgen05.ld .pack::16b, through the UMMA self-test."""
import os, sys, subprocess, numpy as np
if len(sys.argv) > 1:
    import torch
    sys.path.insert(0, '.')
    from deeppicarditeration_b200.engine import LabelEngine
    e = LabelEngine("cuda:0")
    torch.manual_seed(0)
    A = torch.randn(128, 128) * 0.5
    B = torch.randn(128, 128) * 0.5
    D = e.selftest_umma(A, B).cpu()
    ref = (A.half().double() @ B.half().double().t())
    mode = int(os.environ.get("DPI_SELFTEST_MODE", "0"))
    if mode == 1:
        raw = D.view(torch.int32)
        lo = (raw & 0xFFFF).to(torch.int16).view(torch.float16).double()
        hi = ((raw >> 16) & 0xFFFF).to(torch.int16).view(torch.float16).double()
        print("mode 1: rel err low half vs ref", float((lo - ref).norm() / ref.norm()), " high half norm", float(hi.norm()),
              " first cells", [hex(int(x) & 0xffffffff) for x in raw[0, :4]])
    else:
        print("mode", mode, ": rel err vs ref", float((D.double() - ref).norm() / ref.norm()), " max abs", float((D.double() - ref).abs().max()))
else:
    for m in ("0", "1", "2"):
        subprocess.run([sys.executable, __file__, "run"], env=dict(os.environ, DPI_SELFTEST_MODE=m))
