/* c_abi_gas_cell.c -- the C ABI of libpyrad_b200.so used from plain C (no Python, no torch): a small CO2 gas cell
 * through upload -> grid -> prepass -> line sum -> layer stream, the call sequence that replaces
 * Isotope.createCrossSection + Layer.transmittance (pyradClasses.py:361-407, 714-716).
 *
 *   gcc -O2 -I include examples/c_abi_gas_cell.c -o build/c_abi_gas_cell -L pyrad_b200 -lpyrad_b200 \
 *       -Wl,-rpath,$PWD/pyrad_b200 -lm
 *
 * Then the same cell with its lines held as TWO isotopologue lists (odd / even lines), uploaded from where they are with
 * prb_upload_line_groups: one prepass + one line-sum launch give a cross-section row per list (the per-isotopologue loop
 * of Layer.createCrossSection, :498-503, 566-576), and prb_layer_spectra_resident forms the transmittance on the device
 * from those rows -- it must agree with the single-list result.
 *
 * Prints the pair count, the largest absorption coefficient and the mean transmittance; exits 0 on success and
 * non-zero with the library's error text otherwise (on a machine without a B200 that is PRB_ERR_NODEVICE: there is
 * no CPU fallback). */
#include <math.h>
#include <stdio.h>
#include <stdlib.h>

#include "pyrad_b200.h"

#define CHECK(call)                                                                   \
    do {                                                                              \
        int rc_ = (call);                                                             \
        if (rc_ != PRB_OK) {                                                          \
            fprintf(stderr, "%s -> %d: %s\n", #call, rc_, prb_last_error());          \
            return rc_ == PRB_ERR_NODEVICE ? 77 : 1;                                  \
        }                                                                             \
    } while (0)

int main(void) {
    enum { N_LINES = 2000 };
    const double range_min = 600.0, range_max = 700.0, res = 0.01;
    const double T = 296.0, P = 1013.25, conc = 400e-6, molmass = 43.98983, q296 = 286.09, depth_cm = 10.0;
    const int64_t n_grid = (int64_t)((range_max - range_min) / res);
    const int64_t window = 500;                       /* len(np.arange(0, 5 * P / 1013.25, res)) */
    static double nu[N_LINES], sw[N_LINES], ga[N_LINES], gs[N_LINES], el[N_LINES], na[N_LINES], da[N_LINES];
    unsigned int seed = 12345u;
    for (int i = 0; i < N_LINES; ++i) {               /* ascending wavenumbers, HITRAN-like magnitudes */
        seed = seed * 1664525u + 1013904223u;
        const double u = (seed >> 8) / 16777216.0;
        nu[i] = 596.0 + (i + u) * (108.0 / N_LINES);
        sw[i] = pow(10.0, -24.0 + 4.0 * u);
        ga[i] = 0.05 + 0.05 * u; gs[i] = 0.07 + 0.05 * u; el[i] = 2000.0 * u; na[i] = 0.6 + 0.2 * u; da[i] = -0.003 * u;
    }
    prb_engine *e = NULL;
    CHECK(prb_create(0, &e));
    CHECK(prb_upload_lines(e, N_LINES, nu, sw, ga, gs, el, na, da, NULL, 1));
    CHECK(prb_set_grid(e, range_min, res, n_grid, 0, n_grid));
    CHECK(prb_layer_prepass(e, T, P, 1, &conc, &molmass, &q296, &q296, NULL, window));
    double *sigma = malloc(sizeof(double) * n_grid), *k = malloc(sizeof(double) * n_grid), *t = malloc(sizeof(double) * n_grid);
    CHECK(prb_line_sum(e, sigma));
    const double weight = conc * P / 1E4 / 1.38064852E-23 / T;                      /* absCoef, pyradClasses.py:583 */
    CHECK(prb_layer_stream(e, n_grid, 1, sigma, &weight, depth_cm, T, range_min, (range_max - range_min) / (n_grid - 1),
                           range_max, NULL, k, t, NULL));
    double kmax = 0, tmean = 0;
    for (int64_t i = 0; i < n_grid; ++i) { kmax = k[i] > kmax ? k[i] : kmax; tmean += t[i]; }
    printf("c_abi_gas_cell ok: abi %d, %lld pairs, k_max %.6e cm^-1, mean transmittance %.6f\n", prb_abi_version(),
           (long long)prb_pair_count(e), kmax, tmean / n_grid);
    /* --- the same lines as two groups, per-group rows in one pass, spectra from the device-resident rows --- */
    enum { HALF = N_LINES / 2 };
    static double g_nu[2][HALF], g_sw[2][HALF], g_ga[2][HALF], g_gs[2][HALF], g_el[2][HALF], g_na[2][HALF], g_da[2][HALF];
    for (int i = 0; i < N_LINES; ++i) {
        const int g = i & 1, j = i >> 1;
        g_nu[g][j] = nu[i]; g_sw[g][j] = sw[i]; g_ga[g][j] = ga[i]; g_gs[g][j] = gs[i];
        g_el[g][j] = el[i]; g_na[g][j] = na[i]; g_da[g][j] = da[i];
    }
    const int64_t counts[2] = {HALF, HALF};
    const double *p_nu[2] = {g_nu[0], g_nu[1]}, *p_sw[2] = {g_sw[0], g_sw[1]}, *p_ga[2] = {g_ga[0], g_ga[1]},
                 *p_gs[2] = {g_gs[0], g_gs[1]}, *p_el[2] = {g_el[0], g_el[1]}, *p_na[2] = {g_na[0], g_na[1]},
                 *p_da[2] = {g_da[0], g_da[1]};
    const double conc2[2] = {conc, conc}, mass2[2] = {molmass, molmass}, q2[2] = {q296, q296}, w2[2] = {weight, weight};
    CHECK(prb_upload_line_groups(e, 2, counts, p_nu, p_sw, p_ga, p_gs, p_el, p_na, p_da));
    CHECK(prb_set_grid(e, range_min, res, n_grid, 0, n_grid));
    CHECK(prb_layer_prepass(e, T, P, 2, conc2, mass2, q2, q2, NULL, window));
    double *rows = malloc(sizeof(double) * 2 * n_grid), *t2 = malloc(sizeof(double) * n_grid);
    CHECK(prb_line_sum_groups(e, rows));
    CHECK(prb_layer_spectra_resident(e, w2, NULL, depth_cm, T, range_max, NULL, NULL, t2, NULL));
    double dmax = 0, smax = 0;
    for (int64_t i = 0; i < n_grid; ++i) {
        const double ds = fabs(rows[i] + rows[n_grid + i] - sigma[i]), dt = fabs(t2[i] - t[i]);
        smax = ds > smax * sigma[i] ? ds / (sigma[i] > 0 ? sigma[i] : 1) : smax;
        dmax = dt > dmax ? dt : dmax;
    }
    printf("c_abi_gas_cell groups: 2 rows in one pass, max |dT| vs the single list %.2e, max rel d(sigma) %.2e\n", dmax, smax);
    free(sigma); free(k); free(t); free(rows); free(t2);
    CHECK(prb_destroy(e));
    return (kmax > 0 && tmean > 0 && dmax <= 1e-6 && smax <= 1e-5) ? 0 : 1;
}
