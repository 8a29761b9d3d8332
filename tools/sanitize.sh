#!/bin/bash
# compute-sanitizer memcheck + racecheck of a small fit + predict + append + latency-path predict (run under gpurun, 1 GPU).
OUT=gpurun_out
cat > /tmp/san.py <<'PY'
import sys
sys.path.insert(0, '.')
import numpy as np
from thesis_code_b200 import GPModel, synthetic
c = synthetic.make_config("C3", N=int(sys.argv[1]), Ns=200)
m = GPModel.from_config(c["model"]["kwargs"])
n0 = c["N"] - 40
m.init(c["X"][:n0], c["Y"][:n0])
mean, var = m.get_statistics(c["Xs"], full_cov=False)
m.add_observations(c["X"][n0:], c["Y"][n0:])
mean1, var1 = m.get_statistics(c["Xs"][:3], full_cov=False)
mu = m.get_mean(c["Xs"][:1])
chk = m._get_handle().selfcheck()
print("ok", float(mean.sum()), float(var.min()), chk["factor_residual"], chk["solve_residual"])
PY
compute-sanitizer --tool memcheck --error-exitcode 9 python /tmp/san.py 700 > $OUT/r02_memcheck.log 2>&1; echo "memcheck rc=$?"; tail -4 $OUT/r02_memcheck.log
compute-sanitizer --tool racecheck --error-exitcode 9 python /tmp/san.py 400 > $OUT/r02_racecheck.log 2>&1; echo "racecheck rc=$?"; tail -4 $OUT/r02_racecheck.log
compute-sanitizer --tool synccheck --error-exitcode 9 python /tmp/san.py 400 > $OUT/r02_synccheck.log 2>&1; echo "synccheck rc=$?"; tail -3 $OUT/r02_synccheck.log
