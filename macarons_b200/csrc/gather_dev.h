// Device side of the score exchange: wait until every rank's columns have landed in the local score board, then
// take the replicated NBV argmax (reference testers/shapenet.py:172, testers/scene.py:454: first maximum wins).
// Used by the stand-alone kernel of gather.cu and by the finishing CTA of the scoring kernel (covgain.cu).
#pragma once
#include "mac_common.h"

namespace mac {

// total order used for the argmax: NaN beats everything (torch.argmax semantics), then larger value, then lower index
__device__ __forceinline__ bool score_better(float av, int ai, float bv, int bi)
{
    const bool an = av != av, bn = bv != bv;
    if (an != bn) return an;
    if (!an && av != bv) return av > bv;
    return ai < bi;
}

// Called by all NT threads of ONE CTA: wait (bounded, ~2 s) until the `world` arrival flags have reached `epoch`.
// Returns false (and raises the sticky *status) if a peer never arrived.
template <int NT>
__device__ bool wait_for_ranks(const unsigned int *flags, int world, unsigned int epoch, int *status)
{
    __shared__ int timed_out;
    unsigned long long t_begin = 0;
    if (threadIdx.x == 0) {
        timed_out = 0;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t_begin));
    }
    __syncthreads();
    if (threadIdx.x < world) {
        unsigned long long t0, t;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
        // flags carry epochs that only grow; compare as a signed distance so that wrap-around is harmless
        while (static_cast<int>(ld_acquire_sys(flags + threadIdx.x) - epoch) < 0) {
            asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
            if (t - t0 > 2000000000ull) {  // 2 s: a peer died; do not hang the device
                timed_out = 1;
                break;
            }
            __nanosleep(100);
        }
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        // instrumentation: how long this rank waited for the slowest peer (status[1] = last wait in ns, status[2] += ns,
        // status[3] += 1); the host zeroes status[1..3] when it starts a measurement
        unsigned long long t1;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t1));
        const unsigned long long dt = t1 - t_begin;
        status[1] = static_cast<int>(dt > 0x7fffffffull ? 0x7fffffffull : dt);
        status[2] += status[1];
        status[3] += 1;
    }
    if (timed_out) {
        // status[0] is sticky: a timeout of an earlier step stays visible until the host clears it (PeerScoreBoard.check)
        if (threadIdx.x == 0) status[0] = 1;
        return false;
    }
    return true;
}

// best[b] = first maximum of scores[b, :] (all NT threads of one CTA).  One warp per row, shuffle reduction under the
// total order of score_better: no block barrier per row (the first version reduced every row through shared memory with
// 8 barriers; at 32 clouds -- BASELINE config 4 -- that loop alone took ~40 us of the finishing CTA).
template <int NT>
__device__ void argmax_rows(const float *scores, int B, int C, long long *best)
{
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int b = warp; b < B; b += NT / 32) {
        float v = 0.f;
        int idx = 0x7fffffff;
        for (int c = lane; c < C; c += 32) {
            const float x = __ldcg(scores + static_cast<size_t>(b) * C + c);  // written by peers: skip L1
            if (idx == 0x7fffffff || score_better(x, c, v, idx)) {
                v = x;
                idx = c;
            }
        }
#pragma unroll
        for (int d = 16; d >= 1; d >>= 1) {
            const float ov = __shfl_xor_sync(0xffffffffu, v, d);
            const int oi = __shfl_xor_sync(0xffffffffu, idx, d);
            if (oi != 0x7fffffff && (idx == 0x7fffffff || score_better(ov, oi, v, idx))) {
                v = ov;
                idx = oi;
            }
        }
        if (lane == 0) best[b] = idx;
    }
}

template <int NT>
__device__ void wait_scores_and_argmax(const float *scores, const unsigned int *flags, int world, unsigned int epoch, int B,
                                       int C, long long *best, int *status)
{
    if (!wait_for_ranks<NT>(flags, world, epoch, status)) return;
    argmax_rows<NT>(scores, B, C, best);
}

}  // namespace mac
