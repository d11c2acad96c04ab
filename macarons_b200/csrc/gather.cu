// Consumer side of the fused score all-gather (see include/macarons_b200.h): wait until every rank's columns
// have landed in the local score board, then take the replicated NBV argmax (reference testers/shapenet.py:172,
// testers/scene.py:454: first maximum wins).
#include "gather_dev.h"

namespace mac {
namespace {

constexpr int kGatherThreads = 256;

__global__ void __launch_bounds__(kGatherThreads) gather_wait_argmax_kernel(const float *scores, const unsigned int *flags,
                                                                            int world, unsigned int epoch, int B, int C,
                                                                            long long *best, int *status)
{
    wait_scores_and_argmax<kGatherThreads>(scores, flags, world, epoch, B, C, best, status);
}

}  // namespace
}  // namespace mac

extern "C" int mac_gather_wait_argmax(const float *scores, const unsigned int *flags, int world, unsigned int epoch,
                                      int B, int C, long long *best, int *status, void *stream)
{
    using namespace mac;
    MAC_REQUIRE(scores && flags && best && status, "null pointer");
    MAC_REQUIRE(world >= 1 && world <= MAC_MAX_PEERS, "world must be in [1, %d] (got %d)", MAC_MAX_PEERS, world);
    MAC_REQUIRE(B > 0 && C > 0, "B and C must be positive (got %d, %d)", B, C);
    gather_wait_argmax_kernel<<<1, kGatherThreads, 0, static_cast<cudaStream_t>(stream)>>>(scores, flags, world, epoch, B, C,
                                                                                            best, status);
    MAC_CUDA(cudaGetLastError());
    count_launch();
    return MAC_OK;
}
