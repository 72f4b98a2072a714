"""Probe of the tcgen05 contraction variants (RB_GEMM_MODE): error vs float64 and wall time for one large product."""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from relion_b200.estep import MlDeviceBundle

def main():
    dev = MlDeviceBundle(0)
    for (M, N, K) in [(128, 256, 32), (256, 256, 64), (200, 300, 840), (640, 1000, 3030), (4096, 8192, 4096)]:
        rng = np.random.default_rng(M + N + K)
        A = rng.standard_normal((M, K)).astype(np.float32)
        B = rng.standard_normal((N, K)).astype(np.float32)
        try:
            t0 = time.time()
            got = dev.gemm_tf32x3(A, B)
            dt = time.time() - t0
        except Exception as e:  # noqa: BLE001
            print("FAILED", (M, N, K), repr(e)[:600])
            return 1
        m = min(M, 512); n = min(N, 512)
        want = A[:m].astype(np.float64) @ B[:n].astype(np.float64).T
        bound = np.abs(A[:m]).astype(np.float64) @ np.abs(B[:n]).astype(np.float64).T
        err = np.abs(got[:m, :n] - want) / bound
        print((M, N, K), "max err %.3g mean err %.3g wall %.3f s" % (err.max(), err.mean(), dt))
    return 0

if __name__ == "__main__":
    sys.exit(main())
