/* Plain C99 consumer of include/ifb200.h: what a JNI / cgo stub would compile against.
 * Builds a 3-node forest, asks for its info and (when a GPU is present) scores three rows from host buffers. */
#include <stdio.h>
#include <string.h>

#include "ifb200.h"

int main(void) {
    if (ifb_abi_version() != IFB_ABI_VERSION) return 10;
    if (ifb_avg_path_length(2) <= 0.15f || ifb_avg_path_length(2) >= 0.16f) return 11;
    int32_t ndev = 0;
    int rc = ifb_device_count(&ndev);
    if (rc != IFB_OK || ndev == 0) {
        /* CPU-only box: creating a forest must fail loudly with IFB_ENOGPU, never fall back */
        const int32_t node_off[2] = {0, 1}, left[1] = {-1}, right[1] = {-1}, feature[1] = {-1};
        const double thr[1] = {0.0};
        const int64_t ninst[1] = {256};
        ifb_forest *f = NULL;
        rc = ifb_forest_create_standard(0, 1, node_off, left, right, feature, thr, ninst, 256, 1, &f);
        if (rc != IFB_ENOGPU || f != NULL || strlen(ifb_last_error()) == 0) return 12;
        printf("abi_smoke ok (no GPU: IFB_ENOGPU, \"%s\")\n", ifb_last_error());
        return 0;
    }
    /* IFT/IsolationTreeTest.scala:27-42: root splits feature 0 at 1.5, leaves of 10 and 20 instances */
    const int32_t node_off[2] = {0, 3}, left[3] = {1, -1, -1}, right[3] = {2, -1, -1}, feature[3] = {0, -1, -1};
    const double thr[3] = {1.5, 0.0, 0.0};
    const int64_t ninst[3] = {-1, 10, 20};
    ifb_forest *f = NULL;
    rc = ifb_forest_create_standard(0, 1, node_off, left, right, feature, thr, ninst, 256, 1, &f);
    if (rc != IFB_OK) { fprintf(stderr, "%s\n", ifb_last_error()); return 13; }
    ifb_forest_info info;
    if (ifb_forest_get_info(f, &info) != IFB_OK || info.num_nodes != 3 || info.max_depth != 1) return 14;
    const float x[3] = {1.0f, 2.0f, 1.5f};
    double scores[3];
    float psum[3];
    rc = ifb_score_host(f, x, 3, 1, 3, IFB_COL_MAJOR, scores, NULL, psum);
    if (rc != IFB_OK) { fprintf(stderr, "%s\n", ifb_last_error()); return 15; }
    if (psum[0] != 4.7488804f || psum[1] != 6.143309f || psum[2] != 6.143309f) return 16;  /* exact f32 KATs */
    if (ifb_score_host(f, x, 3, 2, 3, IFB_COL_MAJOR, scores, NULL, NULL) != IFB_EINVAL) return 17;
    ifb_forest_destroy(f);
    printf("abi_smoke ok (GPU: path lengths %.7f %.7f)\n", psum[0], psum[1]);
    return 0;
}
