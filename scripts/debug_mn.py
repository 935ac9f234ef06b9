"""Decode what the tensor core actually reads for MN-major operand tiles (debug aid)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import vulkpy_b200 as vk
gpu = vk.GPU(0)
F = np.float32
TC = 2

def gemm(a, b, ta, tb, M, N, K):
    A, B = vk.Array(gpu, data=a), vk.Array(gpu, data=b)
    C = vk.Array(gpu, shape=(M, N))
    C.job = gpu.gpu.gemm(ta, tb, M, N, K, A.buffer, B.buffer, C.buffer, None, TC)
    return np.asarray(C).copy()

for (M, N, K) in ((128, 128, 32), (128, 256, 64)):
    print("=== NN: B [K,N] MN-major, A selector", M, N, K)
    a = np.zeros((M, K), F); a[np.arange(K), np.arange(K)] = 1          # C[m, n] = B[m, n] for m < K
    b = (np.arange(K)[:, None] * 1000 + np.arange(N)[None, :]).astype(F)
    c = gemm(a, b, False, False, M, N, K)
    k_src, n_src = (c[:K] // 1000).astype(int), (c[:K] % 1000).astype(int)
    ok = (c[:K] == b)
    print("match fraction", ok.mean(), "nonzero rows beyond K:", int((c[K:] != 0).sum()))
    for m in (0, 1, 2, 7, 8, 9, 16, 31):
        if m < K:
            print(f" row m={m}: n=0..7 -> (k',n') =", [(int(k_src[m, n]), int(n_src[m, n])) for n in range(8)], " n=32,33,64,127:",
                  [(int(k_src[m, n]), int(n_src[m, n])) for n in (32, 33, 64, 127)])
    print("=== TN: A [K,M] MN-major, B selector", M, N, K)
    a = (np.arange(K)[:, None] * 1000 + np.arange(M)[None, :]).astype(F)  # stored [K, M]
    b = np.zeros((N, K), F); b[np.arange(K), np.arange(K)] = 1          # C[m, n] = A[k=n, m] for n < K
    c = gemm(a, b, True, True, M, N, K)
    want = a.T[:, :K]
    print("match fraction", (c[:, :K] == want).mean(), "nonzero cols beyond K:", int((c[:, K:] != 0).sum()))
    k_src, m_src = (c[:, :K] // 1000).astype(int), (c[:, :K] % 1000).astype(int)
    for n in (0, 1, 8, 9):
        print(f" col n(k)={n}: m=0..7 -> (k',m') =", [(int(k_src[m, n]), int(m_src[m, n])) for m in range(8)], " m=32,33,64,127:",
              [(int(k_src[m, n]), int(m_src[m, n])) for m in (32, 33, 64, 127)])
