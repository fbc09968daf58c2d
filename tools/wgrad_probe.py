import sys; sys.path.insert(0, "/root/repo"); sys.path.insert(0, "/root/repo/tools")
import gemm_probe as G
M = 73728
G.run(1024, 256, M, transA=True, accumulate=True, out16=False)
G.run(768, 256, M, transA=True, accumulate=True, out16=False)
G.run(256, 256, M, transA=True, accumulate=True, out16=False)
