"""Isolated timings of the token-major (weight-gradient) tcgen05 GEMM at the bench shapes: dW[N_out, K_in] += dY^T X with
K = 73,728 token rows (CUDA events)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import gemm_probe as G

if __name__ == "__main__":
    M = 73728
    G.run(1024, 256, M, transA=True, accumulate=True, out16=False)   # FFN1 / FFN2 weight gradients
    G.run(768, 256, M, transA=True, accumulate=True, out16=False)    # QKV
    G.run(256, 256, M, transA=True, accumulate=True, out16=False)    # out-proj
