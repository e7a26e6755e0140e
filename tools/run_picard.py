"""Developer helper: Burgers-100d Deep Picard Iteration end to end on the GPU stages (BASELINE config 2 shapes)."""
import sys, json
sys.path.insert(0, '.')
from deeppicarditeration_b200 import equations as eqs
from deeppicarditeration_b200.picard import run_picard
n_iter = int(sys.argv[1]) if len(sys.argv) > 1 else 20
prec = sys.argv[2] if len(sys.argv) > 2 else "f16"
eq = eqs.Cha(nx=100, alpha=1.0, k=5.0, T=1.0)
out = run_picard(eq, [128] * 4, n_iterations=n_iter, data_size=4096, n_mc=4096, batch_size=512, n_epochs=16,
                 fixed_weight=1.0, precision=prec, eval_points=10000, eval_freq=max(1, n_iter // 10))
print(json.dumps({"label_s": out["label_s"], "train_s": out["train_s"], "final": out["history"][-1]}))
