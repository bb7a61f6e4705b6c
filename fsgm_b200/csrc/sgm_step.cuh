// One 1-D path step on 2*NREG consecutive labels per lane, u16x2 lanes, no-wrap domain (see aggregate.cu header):
//   L(d) = C(d) + min( Lpre(d), Lpre(d-1)+P1, Lpre(d+1)+P1, M+P2 ) - M        M = min_k Lpre(k)
// (reference calc_cost_sgm.cpp:33-66).  Shared by the generic one-warp-per-scanline kernel (aggregate.cu) and the
// row-synchronous cluster kernel (vsweep.cu).
#pragma once
#include <stdint.h>
#include <cuda_fp16.h>

namespace fsgm {

constexpr uint32_t STEP_BIG2 = 0x3F003F00u;      // "infinite" u16 pair: never wins a min, never overflows when P1/P2 are added

// unpack 2*NREG cost bytes (NREG/2 32-bit words, or one 16-bit half for NREG == 1) to u16x2
template <int NREG>
__device__ __forceinline__ void unpack_cost(const uint32_t* w, uint32_t (&c)[NREG])
{
#pragma unroll
    for (int i = 0; i < NREG; ++i) c[i] = __byte_perm(w[i >> 1], 0, (i & 1) ? 0x4342 : 0x4140);
}

// pack u16x2 values (<= 255) back to bytes
template <int NREG>
__device__ __forceinline__ void pack_cost(const uint32_t (&L)[NREG], uint32_t* w)
{
#pragma unroll
    for (int k = 0; k < (NREG + 1) / 2; ++k)
        w[k] = (2 * k + 1 < NREG) ? __byte_perm(L[2 * k], L[2 * k + 1], 0x6420) : __byte_perm(L[2 * k], 0, 0x6420);
}

// lo_mask / hi_mask: STEP_BIG2-style masks that are non-zero only in lane 0 (low half) / lane 31 (high half).
// cP2 = c + P2 (packed).  The far term is applied on the output side:
//   C + min(b, M+P2) - M  ==  min(C + (b - M), C + P2),        b = min(Lpre(d), min(Lpre(d-1), Lpre(d+1)) + P1) >= M
// so a step costs per register two VIADDMNMX + one VIMNMX on the ALU pipe and one plain subtraction (b - M, free to
// go to the FMA pipe as IMAD.IADD).  ALU-pipe instructions are what bounds the aggregation (ncu: alu pipe ~78 %).
template <int NREG>
__device__ __forceinline__ uint32_t sgm_step_u16(const uint32_t (&c)[NREG], const uint32_t (&cP2)[NREG], const uint32_t (&Lpre)[NREG],
                                                 uint32_t M, uint32_t P1P1, uint32_t lo_mask, uint32_t hi_mask, uint32_t (&L)[NREG])
{
    uint32_t q[NREG + 1];
    const uint32_t up = __shfl_up_sync(0xffffffffu, Lpre[NREG - 1], 1);
    const uint32_t dn = __shfl_down_sync(0xffffffffu, Lpre[0], 1);
    q[0] = __byte_perm(up, Lpre[0], 0x5432) | lo_mask;               // (label 2i-1, label 2i)
#pragma unroll
    for (int i = 1; i < NREG; ++i) q[i] = __byte_perm(Lpre[i - 1], Lpre[i], 0x5432);
    q[NREG] = __byte_perm(Lpre[NREG - 1], dn, 0x5432) | hi_mask;
    const uint32_t MM = M * 0x10001u;
    uint32_t m = 0xFFFFFFFFu;
#pragma unroll
    for (int i = 0; i < NREG; ++i) {
        const uint32_t nb = __vminu2(q[i], q[i + 1]);                 // min(Lpre(d-1), Lpre(d+1))
        const uint32_t b = __viaddmin_u16x2(nb, P1P1, Lpre[i]);       // min(nb + P1, Lpre(d))
        L[i] = __viaddmin_u16x2(c[i], b - MM, cP2[i]);                // min(C + b - M, C + P2); b >= M in both halves: no borrow
        m = __vminu2(m, L[i]);
    }
    m = min(m & 0xFFFFu, m >> 16);
    return __reduce_min_sync(0xffffffffu, m);
}

// ---------------------------------------------------------------------------------------------------------------------------
// The same step with the additions on the FMA pipe.  ncu shows every aggregation kernel bound by the integer ALU pipe
// (VIMNMX / VIADDMNMX / PRMT / LOP3 issue there, one warp instruction per two cycles per sub-partition) while the FMA pipe,
// which has the same issue rate, idles.  Values are therefore carried as fp16x2 numbers with a bias of 1024:
//     v  <->  half(1024 + v),  bit pattern 0x6400 | v   (exact for 0 <= v < 1024: the ulp in [1024, 2048) is 1)
//   * min() still runs on the ALU pipe, on the BIT PATTERNS (positive halves order like unsigned integers): VIMNMX.U16x2;
//   * +P1 and +P2 are HADD2 on the FMA pipe (adding a small integer keeps the bias and stays below 2048);
//   * the far term  C + min(b - M, P2)  =  (C + P2) - relu((P2 + M) - b)  is HFMA2.RELU + HADD2: two FMA-pipe instructions and
//     no ALU-pipe instruction (the integer form spends a VIADDMNMX on it);
//   * bytes <-> biased halves cost nothing extra: unpacking is the same PRMT with 0x64 as the high byte, packing takes the
//     low bytes as before.
// Every intermediate is an integer below 2048 in magnitude, so the fp16 arithmetic is exact and the results are bit-identical to the integer step.
// ---------------------------------------------------------------------------------------------------------------------------
constexpr uint32_t H2_BIAS2 = 0x64006400u;       // half2(1024, 1024)
constexpr uint32_t H2_NEG1 = 0xBC00BC00u;        // half2(-1, -1)

__device__ __forceinline__ uint32_t h2_add(uint32_t a, uint32_t b) { uint32_t r; asm("add.rn.f16x2 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b)); return r; }
__device__ __forceinline__ uint32_t h2_sub(uint32_t a, uint32_t b) { uint32_t r; asm("sub.rn.f16x2 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b)); return r; }
// relu(k - b)
__device__ __forceinline__ uint32_t h2_relu_diff(uint32_t k, uint32_t b)
{
    uint32_t r;
    asm("fma.rn.relu.f16x2 %0, %1, %2, %3;" : "=r"(r) : "r"(b), "r"(H2_NEG1), "r"(k));
    return r;
}
// small non-negative integer (< 1024) -> half2(v, v), unbiased
__device__ __forceinline__ uint32_t h2_const(int v) { return (uint32_t)__half_as_ushort(__int2half_rn(v)) * 0x10001u; }

// bytes -> biased halves
template <int NREG>
__device__ __forceinline__ void unpack_cost_h2(const uint32_t* w, uint32_t (&c)[NREG])
{
#pragma unroll
    for (int i = 0; i < NREG; ++i) c[i] = __byte_perm(w[i >> 1], 0x64646464u, (i & 1) ? 0x4342 : 0x4140);
}
// biased halves (values <= 255) -> bytes: pack_cost<NREG>() as it is (it keeps the low byte of every half)

// cP2 = C + P2 and Lpre biased; MM = biased minimum of Lpre in BOTH halves; P1h/P2h = h2_const(P1/P2).
// sel_lo / sel_hi = h2_edge_sel_lo/hi(lane): PRMT selectors of the two words that cross the lane boundary.  Label 0 has no
// neighbour below and label D-1 none above: lane 0 / lane 31 select the label's OWN value there instead of the shuffled one,
// which never wins (min(l + P1, l) = l), so no "infinity" has to be OR-ed in.
// Instruction diet (the aggregation kernels are bound by instruction issue): P1 is added to the whole previous row once, the
// neighbour words are built from that row, and min(Lpre(d-1)+P1, Lpre(d+1)+P1, Lpre(d)) is ONE 3-input VIMNMX3.U16x2; the
// minimum over the lane's labels is folded to both halves (PRMT swap + VIMNMX) so that the 32-bit CREDUX.MIN of the bit
// patterns returns it duplicated, ready to be the next step's MM.  Per u16x2 register: 1 ALU-pipe + 3 FMA-pipe instructions.
// Returns the biased minimum of L in both halves.
__device__ __forceinline__ uint32_t h2_edge_sel_lo(int lane) { return lane == 0 ? 0x5454u : 0x5432u; }
__device__ __forceinline__ uint32_t h2_edge_sel_hi(int lane) { return lane == 31 ? 0x3232u : 0x5432u; }

template <int NREG>
__device__ __forceinline__ uint32_t sgm_step_h2(const uint32_t (&cP2)[NREG], const uint32_t (&Lpre)[NREG], uint32_t MM, uint32_t P1h,
                                                uint32_t P2h, uint32_t sel_lo, uint32_t sel_hi, uint32_t (&L)[NREG])
{
    uint32_t Lp[NREG], q[NREG + 1];
#pragma unroll
    for (int i = 0; i < NREG; ++i) Lp[i] = h2_add(Lpre[i], P1h);      // Lpre + P1
    const uint32_t up = __shfl_up_sync(0xffffffffu, Lp[NREG - 1], 1);
    const uint32_t dn = __shfl_down_sync(0xffffffffu, Lp[0], 1);
    q[0] = __byte_perm(up, Lp[0], sel_lo);                            // (label 2i-1, label 2i) + P1
#pragma unroll
    for (int i = 1; i < NREG; ++i) q[i] = __byte_perm(Lp[i - 1], Lp[i], 0x5432);
    q[NREG] = __byte_perm(Lp[NREG - 1], dn, sel_hi);
    const uint32_t K = h2_add(MM, P2h);                               // 1024 + M + P2
    uint32_t m = 0xFFFFFFFFu;
#pragma unroll
    for (int i = 0; i < NREG; ++i) {
        const uint32_t b = __vminu2(__vminu2(q[i], q[i + 1]), Lpre[i]);   // min(Lpre(d-1) + P1, Lpre(d+1) + P1, Lpre(d))
        L[i] = h2_sub(cP2[i], h2_relu_diff(K, b));                    // (C + P2) - relu(P2 + M - b)
        m = __vminu2(m, L[i]);
    }
    m = __vminu2(m, __byte_perm(m, m, 0x1032));
    return __reduce_min_sync(0xffffffffu, m);
}

}  // namespace fsgm
