/* Plain-C caller of the C ABI (include/macarons_b200.h): argument validation and error reporting work without a GPU.
 * Built and run by tests/test_abi_cpu.py::test_c_caller_sees_error_codes. */
#include <stdio.h>
#include <string.h>

#include "macarons_b200.h"

#define CHECK(cond)                                                     \
    do {                                                                \
        if (!(cond)) {                                                  \
            printf("FAILED line %d: %s (last error: %s)\n", __LINE__, #cond, mac_last_error()); \
            return 1;                                                   \
        }                                                               \
    } while (0)

int main(void)
{
    float pts[4 * 8] = {0}, harm[64 * 8] = {0}, cams[3 * 2] = {0}, out[2] = {0};
    unsigned char ws[256] = {0};
    CHECK(mac_version() >= 100);
    CHECK(mac_built_for_sm() == 100);
    CHECK(mac_covgain_workspace_bytes(1, 2) >= 2 * 12);
    CHECK(mac_covgain_workspace_bytes(0, 2) == 0);
    /* null pointers, bad shapes, bad camera ranges, bad activation: rejected before any CUDA call */
    CHECK(mac_covgain_f32(0, 4, harm, cams, out, 1, 8, 2, 0, 2, MAC_ACT_SIGMOID, ws, sizeof ws, 0) == MAC_ERR_INVALID_ARGUMENT);
    CHECK(strstr(mac_last_error(), "null") != 0);
    CHECK(mac_covgain_f32(pts, 4, harm, cams, out, 1, 0, 2, 0, 2, MAC_ACT_SIGMOID, ws, sizeof ws, 0) == MAC_ERR_INVALID_ARGUMENT);
    CHECK(mac_covgain_f32(pts, 2, harm, cams, out, 1, 8, 2, 0, 2, MAC_ACT_SIGMOID, ws, sizeof ws, 0) == MAC_ERR_INVALID_ARGUMENT);
    CHECK(mac_covgain_f32(pts, 4, harm, cams, out, 1, 8, 2, 1, 3, MAC_ACT_SIGMOID, ws, sizeof ws, 0) == MAC_ERR_INVALID_ARGUMENT);
    CHECK(strstr(mac_last_error(), "camera range") != 0);
    CHECK(mac_covgain_f32(pts, 4, harm, cams, out, 1, 8, 2, 0, 2, 7, ws, sizeof ws, 0) == MAC_ERR_INVALID_ARGUMENT);
    CHECK(mac_covgain_f32(pts, 4, harm, cams, out, 1, 8, 2, 0, 2, MAC_ACT_SIGMOID, ws, 8, 0) == MAC_ERR_WORKSPACE);
    CHECK(mac_visibility_f32(pts, 4, harm, 0, out, 1, 8, 2, 0, 2, MAC_ACT_RELU, 0) == MAC_ERR_INVALID_ARGUMENT);
    CHECK(mac_covgain_push_f32(pts, 4, harm, cams, 1, 8, 2, 0, 2, MAC_ACT_SIGMOID, ws, sizeof ws, 0, 0) == MAC_ERR_INVALID_ARGUMENT);
    CHECK(mac_gather_wait_argmax(out, 0, 1, 1, 1, 2, 0, 0, 0) == MAC_ERR_INVALID_ARGUMENT);
    CHECK(mac_knn16_f32(0, 0, 0, 0, 1, 1, 16, 0) != MAC_OK);
    CHECK(mac_unproject_depth_f32(pts, cams, out, 1, 1, 1, 0) == MAC_ERR_INVALID_ARGUMENT);
    CHECK(mac_signed_distance_f32(pts, pts, 0, cams, out, 1, 1, 2, 2, 1.0f, 0) == MAC_ERR_INVALID_ARGUMENT);
    CHECK(mac_sample_proxy_points_f32(pts, pts, harm, pts, 8, 5000, 0.1f, out, harm, 0, 0, ws, sizeof ws, 0) == MAC_ERR_INVALID_ARGUMENT);
    printf("abi harness ok, launches so far: %llu\n", mac_launch_count());
    return 0;
}
