/* consumer.c -- a plain C99 caller of include/jaicov_b200.h, the way a JNI shim or a C host would bind the library.
 * Test infrastructure (tests/test_c_consumer.py): proves that the header is valid C (no C++ leaks), that the library
 * links with nothing but its own name, and that the call sequence of INTEGRATION.md section 4 behaves as documented
 * when no sm_100 device is visible: set_* calls succeed (host copies), every computing call returns
 * JAICOV_NOT_INITIALISED with a message -- there is no CPU path.  With a device the same program runs the adjustment of a
 * tiny network and prints the state id. */
#include <stdio.h>
#include <string.h>

#include "jaicov_b200.h"

#define CHECK(cond)                                                          \
    do {                                                                     \
        if (!(cond)) {                                                       \
            fprintf(stderr, "consumer.c:%d: %s failed\n", __LINE__, #cond);  \
            return 1;                                                        \
        }                                                                    \
    } while (0)

int main(void) {
    jaicov_options o;
    jaicov_handle *h = NULL;
    jaicov_stats st;
    /* one camera (x0, y0, c unknown), two images, four points seen in both: columns in the reference's order
     * (object points, interior orientation, exterior orientations), no datum defect handling needed for the test */
    const double io_val[3] = {0.0, 0.0, 30.0}, r0[1] = {10.0};
    const int32_t io_col[3] = {12, 13, 14}, coef_ptr[2] = {0, 0};
    const int32_t cam_of_img[2] = {0, 0};
    const double eo_val[12] = {-500.0, 0.0, 3000.0, 0.0, 0.1, 0.0, 500.0, 0.0, 3000.0, 0.0, -0.1, 0.0};
    const int32_t eo_col[12] = {15, 16, 17, 18, 19, 20, 21, 22, 23, 24, 25, 26};
    const int64_t pt_ptr[3] = {0, 4, 8};
    const int32_t obj_idx[8] = {0, 1, 2, 3, 0, 1, 2, 3};
    double xy[16], var[16], xyz[12] = {-300, -200, 0, 300, -200, 50, 300, 200, 0, -300, 200, -50};
    const int32_t pt_col[12] = {0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11};
    const uint8_t is_datum[4] = {1, 1, 1, 1};
    const int32_t free_flags[7] = {0, 0, 0, 0, 0, 0, 0};
    int i, id;

    for (i = 0; i < 16; i++) { xy[i] = 0.1 * (i - 8); var[i] = 2.5e-7; }
    CHECK(jaicov_default_options(&o) == JAICOV_OK);
    CHECK(o.invert_mode == JAICOV_INVERT_FULL && o.max_iterations == 5000 && o.solver == JAICOV_SOLVER_AUTO);
    CHECK(sizeof(jaicov_options) == 48 && sizeof(jaicov_stats) == 112);
    o.sigma2apriori = 2.5e-7;
    CHECK(jaicov_create(NULL, &h) == JAICOV_ILLEGAL_ARGUMENT);
    CHECK(jaicov_create(&o, &h) == JAICOV_OK && h != NULL);
    CHECK(jaicov_set_cameras(h, 1, io_val, io_col, r0, coef_ptr, NULL, NULL, NULL, NULL) == JAICOV_OK);
    CHECK(jaicov_set_images(h, 2, cam_of_img, eo_val, eo_col, pt_ptr) == JAICOV_OK);
    CHECK(jaicov_set_image_points(h, 8, obj_idx, xy, var, NULL) == JAICOV_OK);
    CHECK(jaicov_set_object_points(h, 4, xyz, pt_col, is_datum) == JAICOV_OK);
    CHECK(jaicov_set_datum(h, free_flags, 27, 16) == JAICOV_OK);
    id = jaicov_estimate(h, NULL, NULL, NULL);
    if (jaicov_device_count() == 0) {
        double q[1];
        CHECK(id == JAICOV_NOT_INITIALISED);
        CHECK(strlen(jaicov_last_error(h)) > 0);
        CHECK(jaicov_iterate(h, 1, 0) == JAICOV_NOT_INITIALISED);
        CHECK(jaicov_get_qxx_packed(h, q) != JAICOV_OK);
        printf("no sm_100 device: estimate -> %d (%s)\n", id, jaicov_last_error(h));
    } else {
        CHECK(jaicov_get_stats(h, &st) == JAICOV_OK);
        printf("estimate -> %d after %d passes, u = %d\n", id, st.iterations, st.n_unknowns);
    }
    jaicov_destroy(h);
    jaicov_destroy(NULL);
    printf("consumer ok\n");
    return 0;
}
