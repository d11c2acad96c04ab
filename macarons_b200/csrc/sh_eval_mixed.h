// P + u_x Q evaluator of sh_horner_gen.h with a selectable instruction mix: the Horner chains in c_t of the
// columns b <= NPACK run as packed fma.rn.f32x2 (P and Q column together), the others as two scalar FFMA chains;
// the u_z steps are packed when ACC_PACKED.  Every product/sum is the same IEEE fma in the same order as in
// mac_sh_eval2_pq, so all mixes are bitwise identical; only the issue-slot / FMA-pipe balance differs.
// Measured on B200 (cfg5, 512 cameras; profiles/r02n_covgain_variants.txt): all scalar <0,false> 335 us, all packed
// 345 us, the mixes in between 345-365 us: an FFMA2 occupies the issue port for two cycles, so it saves nothing over
// two FFMAs, and the scalar stream lets ptxas keep the multipliers in the operand-reuse cache more often.
#pragma once
#include "sh_horner_gen.h"

#ifdef __CUDACC__
template <int NPACK, bool ACC_PACKED>
__device__ __forceinline__ void mac_sh_eval2_pq_mixed(const float (&g0)[8], const unsigned long long (&gp)[28],
                                                      const float ux0, const float ct0, const float uz0,
                                                      const float ux1, const float ct1, const float uz1, float &z0,
                                                      float &z1)
{
    constexpr int off[8] = {0, MAC_PQ_1, MAC_PQ_2, MAC_PQ_3, MAC_PQ_4, MAC_PQ_5, MAC_PQ_6, MAC_PQ_7};
    const unsigned long long cc0 = mac_pack2(ct0, ct0), cc1 = mac_pack2(ct1, ct1);
    const unsigned long long zz0 = mac_pack2(uz0, uz0), zz1 = mac_pack2(uz1, uz1);
    float ap0, aq0, ap1, aq1;
    mac_unpack2(gp[MAC_PQ_7], ap0, aq0);
    ap1 = ap0;
    aq1 = aq0;
#pragma unroll
    for (int b = 6; b >= 1; --b) {
        const int deg = 7 - b;
        float vp0, vq0, vp1, vq1;
        if (b <= NPACK) {
            unsigned long long v0 = gp[off[b] + deg], v1 = v0;
#pragma unroll
            for (int a = deg - 1; a >= 0; --a) {
                v0 = mac_fma2(v0, cc0, gp[off[b] + a]);
                v1 = mac_fma2(v1, cc1, gp[off[b] + a]);
            }
            mac_unpack2(v0, vp0, vq0);
            mac_unpack2(v1, vp1, vq1);
        } else {
            mac_unpack2(gp[off[b] + deg], vp0, vq0);
            vp1 = vp0;
            vq1 = vq0;
#pragma unroll
            for (int a = deg - 1; a >= 0; --a) {
                float cp, cq;
                mac_unpack2(gp[off[b] + a], cp, cq);
                vp0 = fmaf(vp0, ct0, cp);
                vq0 = fmaf(vq0, ct0, cq);
                vp1 = fmaf(vp1, ct1, cp);
                vq1 = fmaf(vq1, ct1, cq);
            }
        }
        if (ACC_PACKED) {
            unsigned long long a0 = mac_fma2(mac_pack2(ap0, aq0), zz0, mac_pack2(vp0, vq0));
            unsigned long long a1 = mac_fma2(mac_pack2(ap1, aq1), zz1, mac_pack2(vp1, vq1));
            mac_unpack2(a0, ap0, aq0);
            mac_unpack2(a1, ap1, aq1);
        } else {
            ap0 = fmaf(ap0, uz0, vp0);
            aq0 = fmaf(aq0, uz0, vq0);
            ap1 = fmaf(ap1, uz1, vp1);
            aq1 = fmaf(aq1, uz1, vq1);
        }
    }
    float c0 = g0[7], c1 = c0;
#pragma unroll
    for (int a = 6; a >= 0; --a) {
        c0 = fmaf(c0, ct0, g0[a]);
        c1 = fmaf(c1, ct1, g0[a]);
    }
    const float p0 = fmaf(ap0, uz0, c0), p1 = fmaf(ap1, uz1, c1);
    z0 = fmaf(aq0, ux0, p0);
    z1 = fmaf(aq1, ux1, p1);
}
#endif
