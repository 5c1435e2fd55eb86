"""Small driver for compute-sanitizer (memcheck / racecheck / synccheck): a few sweeps of the f32 fast kernel (TPP 1, 2, 4) and of the
generic kernel (f64, Cross family) on small problems.   compute-sanitizer --tool racecheck python tools/sanitize_run.py"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np  # noqa: E402
import erirt_b200 as E  # noqa: E402
from helpers import make_problem, run_engine  # noqa: E402

cases = [("RtIrtLatentQr", 300, 13, 2, "f32"), ("RtIrtLatentQr", 200, 100, 3, "f32"), ("RtIrt", 150, 150, 1, "f32"), ("MlIrt", 130, 21, 0, "f32"),
         ("RtIrtLatentQr", 200, 21, 2, "f64"), ("RtIrtCrossQr", 150, 13, 0, "f32")]
if len(sys.argv) > 1:
    cases = cases[: int(sys.argv[1])]
for model, N, J, F, dt in cases:
    pb = make_problem(model, N, J, F, seed=3)
    eng = run_engine(E, pb, 3, dtype=dt)
    a = eng.get_trace("ra")[:3, N:, 0]
    assert np.all(np.isfinite(a)), (model, dt)
    eng.close()
    print("ok", model, N, J, F, dt, flush=True)
