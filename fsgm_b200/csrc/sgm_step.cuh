// One 1-D path step on 2*NREG consecutive labels per lane, u16x2 lanes, no-wrap domain (see aggregate.cu header):
//   L(d) = C(d) + min( Lpre(d), Lpre(d-1)+P1, Lpre(d+1)+P1, M+P2 ) - M        M = min_k Lpre(k)
// (reference calc_cost_sgm.cpp:33-66).  Shared by the generic one-warp-per-scanline kernel (aggregate.cu) and the
// row-synchronous cluster kernel (vsweep.cu).
#pragma once
#include <stdint.h>

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

}  // namespace fsgm
