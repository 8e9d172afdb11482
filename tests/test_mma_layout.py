"""CPU: the fragment index arithmetic of the tensor-core kernels (csrc/glrm_dense_mma.cuh: mm_gemm1, mm_gemm2), emulated lane by
lane with the PTX layout of mma.sync.m8n8k4.f64 (lane = 4 g + t holds A[g][t], B[t][g], C[g][2t], C[g][2t+1]):
  * U = own * other' lands in the accumulator fragments as the kernel assumes;
  * feeding those fragments back as the A operand with the column sets {0,2,5,7} / {1,3,4,6} and the matching rows of `other`
    computes G += R * other exactly;
  * every shared-memory load pattern is bank-conflict free for the pitch 8 NT + 4 (64-bit accesses, 16 lanes per wavefront)."""
import numpy as np
import pytest


def dmma(c, a, b):
    """D = A(8x4) * B(4x8) + C on 32-lane fragments: a[lane], b[lane] scalars, c[lane] = (c0, c1)."""
    A = np.zeros((8, 4)); B = np.zeros((4, 8)); C = np.zeros((8, 8))
    for lane in range(32):
        g, t = lane >> 2, lane & 3
        A[g, t] = a[lane]; B[t, g] = b[lane]; C[g, 2 * t], C[g, 2 * t + 1] = c[lane]
    D = A @ B + C
    return [(D[lane >> 2, 2 * (lane & 3)], D[lane >> 2, 2 * (lane & 3) + 1]) for lane in range(32)]


@pytest.mark.parametrize("NT,k", [(1, 5), (3, 20), (13, 100)])
def test_fragment_arithmetic_matches_matmul(NT, k):
    P = 8 * NT + 4
    rng = np.random.default_rng(NT)
    own = np.zeros((16, P)); oth = np.zeros((32, P))
    own[:, :k] = rng.standard_normal((16, k)); oth[:, :k] = rng.standard_normal((32, k))     # zero past k, as the factors are stored
    ks = (k + 3) // 4
    # ---- mm_gemm1: acc[mt][nt] over k-steps
    acc = [[[(0.0, 0.0)] * 32 for _ in range(4)] for _ in range(2)]
    for kk in range(ks):
        for mt in range(2):
            a = [own[8 * mt + (lane >> 2), 4 * kk + (lane & 3)] for lane in range(32)]        # own[(row0 + g) * P + t + mt*8*P + 4*kk]
            for nt in range(4):
                b = [oth[8 * nt + (lane >> 2), 4 * kk + (lane & 3)] for lane in range(32)]    # oth[g * P + t + nt*8*P + 4*kk]
                acc[mt][nt] = dmma(acc[mt][nt], a, b)
    U = own[:, :k] @ oth[:, :k].T
    for mt in range(2):
        for nt in range(4):
            for lane in range(32):
                g, t = lane >> 2, lane & 3
                assert np.allclose(acc[mt][nt][lane], (U[8 * mt + g, 8 * nt + 2 * t], U[8 * mt + g, 8 * nt + 2 * t + 1]), rtol=1e-13, atol=1e-13)
    # ---- mm_gemm2: G[mt][n] += R * oth with R = the accumulator fragments
    G = [[[(0.0, 0.0)] * 32 for _ in range(NT)] for _ in range(2)]
    for sb in range(4):
        for n in range(NT):
            b1 = [oth[8 * sb + 2 * (lane & 3) + ((lane & 3) >> 1), 8 * n + (lane >> 2)] for lane in range(32)]       # r1[8 * nt]
            b2 = [oth[8 * sb + 2 * (lane & 3) + 1 - ((lane & 3) >> 1), 8 * n + (lane >> 2)] for lane in range(32)]   # r2[8 * nt]
            for mt in range(2):
                a1 = [acc[mt][sb][lane][1] if (lane & 3) >> 1 else acc[mt][sb][lane][0] for lane in range(32)]
                a2 = [acc[mt][sb][lane][0] if (lane & 3) >> 1 else acc[mt][sb][lane][1] for lane in range(32)]
                G[mt][n] = dmma(dmma(G[mt][n], a1, b1), a2, b2)
    want = U @ oth[:, :8 * NT]
    for mt in range(2):
        for n in range(NT):
            for lane in range(32):
                g, t = lane >> 2, lane & 3
                assert np.allclose(G[mt][n][lane], (want[8 * mt + g, 8 * n + 2 * t], want[8 * mt + g, 8 * n + 2 * t + 1]), rtol=1e-12, atol=1e-12)


@pytest.mark.parametrize("NT", [1, 2, 3, 4, 6, 8, 10, 12, 13])
def test_shared_memory_loads_are_conflict_free(NT):
    P = 8 * NT + 4

    def conflict_free(addr_of_lane):           # 64-bit loads: a wavefront serves 16 lanes, 16 bank pairs of 8 bytes
        for half in (range(0, 16), range(16, 32)):
            banks = [addr_of_lane(lane) % 16 for lane in half]
            if len(set(banks)) != 16:
                return False
        return True

    for kk in range(2):
        assert conflict_free(lambda lane: (lane >> 2) * P + (lane & 3) + 4 * kk)                                   # mm_gemm1 A and B fragments
    for sb in range(4):
        assert conflict_free(lambda lane: (8 * sb + 2 * (lane & 3) + ((lane & 3) >> 1)) * P + (lane >> 2))          # mm_gemm2 r1
        assert conflict_free(lambda lane: (8 * sb + 2 * (lane & 3) + 1 - ((lane & 3) >> 1)) * P + (lane >> 2))      # mm_gemm2 r2
    # the naive pairing (columns 2t and 2t+1 as they sit in the lane) would conflict two ways
    assert not conflict_free(lambda lane: (2 * (lane & 3)) * P + (lane >> 2))
