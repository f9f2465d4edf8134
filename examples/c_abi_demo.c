/* Plain-C consumer of libdigicamtoy_sm100.so: no Python, no torch, no CUDA headers.
 *
 *   c_abi_demo <config.bin> <seed> <n_events>
 *
 * <config.bin> is the flat dump written by digicamtoy_b200.native.export_config(): the scalar
 * fields of dct_config followed by its arrays as float64.  The program generates n_events events
 * into a malloc'ed host buffer through dct_generate_host and prints the sum and a simple checksum
 * of the cube, which tests/test_c_abi_gpu.py compares with the Python path. */
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "digicamtoy_b200.h"

static double* read_doubles(FILE* f, size_t n) {
    double* p = (double*)malloc(n * sizeof(double));
    if (!p || fread(p, sizeof(double), n, f) != n) { fprintf(stderr, "short config file\n"); exit(2); }
    return p;
}

int main(int argc, char** argv) {
    if (argc != 4) { fprintf(stderr, "usage: %s config.bin seed n_events\n", argv[0]); return 2; }
    FILE* f = fopen(argv[1], "rb");
    if (!f) { perror(argv[1]); return 2; }
    uint32_t head[10];
    double grid[3];
    if (fread(head, sizeof(uint32_t), 10, f) != 10 || fread(grid, sizeof(double), 3, f) != 3) return 2;
    dct_config c;
    memset(&c, 0, sizeof(c));
    c.struct_size = sizeof(c);
    c.n_pixels = head[0]; c.n_pulses = head[1]; c.n_bins = head[2]; c.n_drop = head[3];
    c.n_knots = head[4]; c.poisson = head[5]; c.sub_binning = head[6];
    c.adc_min = (int32_t)head[7]; c.adc_max = (int32_t)head[8]; c.kernel = head[9];
    c.t0 = grid[0]; c.t_end = grid[1]; c.dt = grid[2];
    const size_t P = c.n_pixels, K = c.n_pulses;
    c.nsb_rate = read_doubles(f, P); c.crosstalk = read_doubles(f, P); c.gain = read_doubles(f, P);
    c.sigma_1 = read_doubles(f, P); c.sigma_e = read_doubles(f, P); c.baseline = read_doubles(f, P);
    c.n_photon = read_doubles(f, P * K); c.time_signal = read_doubles(f, P * K);
    c.jitter = read_doubles(f, P * K);
    c.knot_x = read_doubles(f, c.n_knots); c.coef = read_doubles(f, ((size_t)c.n_knots - 1) * 4);
    fclose(f);

    dct_handle* h = NULL;
    if (dct_create(&c, 0, &h) != DCT_OK) { fprintf(stderr, "dct_create: %s\n", dct_last_error()); return 1; }
    const uint64_t seed = strtoull(argv[2], NULL, 10);
    const uint32_t n = (uint32_t)strtoul(argv[3], NULL, 10);
    const size_t n_el = (size_t)n * P * c.n_bins;
    int16_t* adc = (int16_t*)malloc(n_el * sizeof(int16_t));
    if (dct_generate_host(h, seed, 0, n, adc, NULL, NULL) != DCT_OK) {
        fprintf(stderr, "dct_generate_host: %s\n", dct_last_error());
        return 1;
    }
    unsigned long long sum = 0, mix = 1469598103934665603ull;
    for (size_t i = 0; i < n_el; ++i) {
        sum += (unsigned long long)(uint16_t)adc[i];
        mix = (mix ^ (uint16_t)adc[i]) * 1099511628211ull;
    }
    printf("abi=%d kernel=%d samples=%zu sum=%llu fnv=%llu\n", dct_abi_version(), dct_selected_kernel(h),
           n_el, sum, mix);
    dct_destroy(h);
    free(adc);
    return 0;
}
