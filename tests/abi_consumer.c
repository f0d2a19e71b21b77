/* Plain-C consumer of include/qxb200.h: builds the 2-qubit GHZ program of the reference docs
 * (docs/src/users_guide.md:71-90) through the per-statement calls, queries the host-side logic and,
 * when a GPU is present (argv[1] == "gpu"), computes the four amplitudes.  Exit code 0 = ok. */
#include <math.h>
#include <stdio.h>
#include <string.h>
#include "qxb200.h"

#define CHECK(x) do { int rc_ = (x); if (rc_ < 0) { fprintf(stderr, "%s -> %d: %s\n", #x, rc_, qxb_last_error()); return 1; } } while (0)

int main(int argc, char** argv) {
    qxb_graph* g = NULL;
    CHECK(qxb_graph_create(&g, QXB_C64));
    int64_t d2[] = {2}, d22[] = {2, 2}, d222[] = {2, 2, 2};
    int64_t l1[] = {1}, l2[] = {2}, l12[] = {1, 2}, l123[] = {1, 2, 3}, l23[] = {2, 3};
    CHECK(qxb_graph_load(g, "t5", "data_1", d2, 1));
    CHECK(qxb_graph_view(g, "t5_s", "t5", "v2", 1, 2));
    CHECK(qxb_graph_load(g, "t1", "data_2", d22, 2));
    CHECK(qxb_graph_view(g, "t1_s", "t1", "v1", 1, 2));
    CHECK(qxb_graph_load(g, "t4", "data_1", d2, 1));
    CHECK(qxb_graph_ncon(g, "I2", l1, 1, "t1_s", l12, 2, "t4", l2, 1));
    CHECK(qxb_graph_ncon(g, "t8", l12, 2, "t5_s", l1, 1, "I2", l2, 1));
    CHECK(qxb_graph_load(g, "t3", "data_4", d2, 1));
    CHECK(qxb_graph_view(g, "t3_s", "t3", "v1", 1, 2));
    CHECK(qxb_graph_ncon(g, "t9", l12, 2, "t8", l12, 2, "t3_s", l2, 1));
    CHECK(qxb_graph_output(g, "t6", 1, 2));
    CHECK(qxb_graph_view(g, "t6_s", "t6", "v1", 1, 2));
    CHECK(qxb_graph_ncon(g, "t10", l12, 2, "t9", l12, 2, "t6_s", l2, 1));
    CHECK(qxb_graph_load(g, "t2", "data_3", d222, 3));
    CHECK(qxb_graph_view(g, "t2_s", "t2", "v1", 3, 2));
    CHECK(qxb_graph_view(g, "t2_s_s", "t2_s", "v2", 2, 2));
    CHECK(qxb_graph_output(g, "t7", 2, 2));
    CHECK(qxb_graph_ncon(g, "I1", l23, 2, "t2_s_s", l123, 3, "t7", l1, 1));
    CHECK(qxb_graph_ncon(g, "t11", NULL, 0, "t10", l12, 2, "I1", l12, 2));
    CHECK(qxb_graph_save(g, "output", "t11"));
    const double s = 0.70710678118654752440;
    double data_1[] = {1, 0, 0, 0}, data_2[] = {s, 0, s, 0, s, 0, -s, 0}, data_4[] = {1, 0, 1, 0};
    double data_3[16] = {0};
    for (int o = 0; o < 2; ++o) for (int i = 0; i < 2; ++i) for (int c = 0; c < 2; ++c)
        if (o == (i ^ c)) data_3[2 * (o + 2 * i + 4 * c)] = 1.0;          /* column-major */
    CHECK(qxb_graph_set_data(g, "data_1", data_1, d2, 1));
    CHECK(qxb_graph_set_data(g, "data_2", data_2, d22, 2));
    CHECK(qxb_graph_set_data(g, "data_3", data_3, d222, 3));
    CHECK(qxb_graph_set_data(g, "data_4", data_4, d2, 1));
    int k = 0, n_out = 0; int64_t dims[8], ns = 0, vals[8];
    CHECK(qxb_graph_num_slice_vars(g, &k, dims));
    CHECK(qxb_graph_num_slices(g, &ns));
    CHECK(qxb_graph_num_outputs(g, &n_out));
    CHECK(qxb_slice_values(g, 2, vals));
    if (k != 2 || dims[0] != 2 || dims[1] != 2 || ns != 4 || n_out != 2 || vals[0] != 0 || vals[1] != 1) {
        fprintf(stderr, "bad bookkeeping\n"); return 1;
    }
    double bytes = 0;
    CHECK(qxb_graph_cost_bytes(g, 3, 4, &bytes));
    char buf[4096];
    if (qxb_graph_program_text(g, buf, sizeof buf) <= 0 || strncmp(buf, "# version:", 10) != 0) return 1;
    if (argc > 1 && strcmp(argv[1], "gpu") == 0) {
        CHECK(qxb_init(0));
        CHECK(qxb_graph_compile(g, NULL));
        uint8_t bits[8] = {0, 0, 1, 1, 0, 1, 1, 0};
        double out[8];
        CHECK(qxb_amplitudes(g, bits, 4, 0, 4, out));
        if (fabs(out[0] - s) > 1e-12 || fabs(out[2] - s) > 1e-12 || fabs(out[4]) > 1e-12 || fabs(out[6]) > 1e-12) {
            fprintf(stderr, "bad amplitudes %g %g %g %g\n", out[0], out[2], out[4], out[6]); return 1;
        }
        printf("gpu amplitudes ok\n");
    } else {
        /* no CPU fallback: compile must fail loudly without a device (or succeed on a GPU box) */
        int rc = qxb_graph_compile(g, NULL);
        if (rc != QXB_OK && rc != QXB_ERR_CUDA) { fprintf(stderr, "unexpected rc %d\n", rc); return 1; }
    }
    qxb_graph_destroy(g);
    printf("abi consumer ok (cost model %.0f bytes)\n", bytes);
    return 0;
}
