import sys, numpy as np
sys.path.insert(0, "."); sys.path.insert(0, "tests")
import test_gpu_fuzz as F
blob = open("scratch/fuzz_lossless_609.jpg", "rb").read()
want, werr = F.run_oracle(blob); got, gerr = F.run_gpu(blob)
print("oracle", werr, "gpu", gerr, "equal", (want is not None and got is not None and np.array_equal(got, want.planes)))
